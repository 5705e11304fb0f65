"""HBM bandwidth by access mix (what the HBM-bound kernels can hope for): copy, pure write, pure read, 1:4."""
import torch
dev = "cuda"
N = 1 << 28          # 268 M bf16 = 512 MiB
a = torch.randn(N, device=dev, dtype=torch.float32).bfloat16()
b = torch.empty_like(a)
small = a[: N // 4]
big4 = torch.empty(4, N // 4, device=dev, dtype=torch.bfloat16)


def timed(fn, nbytes, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:34s} {nbytes / 1e6:8.1f} MB  {best * 1e3:8.1f} us  {nbytes / best / 1e6:7.0f} GB/s", flush=True)


timed(lambda: b.copy_(a), 2 * a.numel() * 2, "copy (read N + write N)")
timed(lambda: b.zero_(), a.numel() * 2, "fill (write N)")
timed(lambda: b.fill_(1.5), a.numel() * 2, "fill value (write N)")
timed(lambda: a.float().sum() if False else torch.sum(a, dtype=torch.float32), a.numel() * 2, "sum (read N)")
timed(lambda: big4.copy_(small.unsqueeze(0).expand(4, -1)), (small.numel() + big4.numel()) * 2, "broadcast (read N/4 + write N)")
timed(lambda: torch.add(a, a, out=b), 2 * a.numel() * 2, "add a+a (read N + write N)")
for mb in (8, 21, 42, 84, 168, 336):
    n = mb * 1000 * 1000 // 4
    x = a[:n]; y = b[:n]
    timed(lambda: y.copy_(x), 4 * n, f"copy of {2 * n * 2 / 1e6:.0f} MB total")
