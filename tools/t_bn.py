"""Times the HBM-bound BatchNorm kernels on the stage-4 layer shapes (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import ops, _lib

for o in sys.argv[1:]:
    k, v = o.split("=")
    _lib.set_option(k, int(v))
dev = torch.device("cuda", 0)
B = 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
LAYERS = [("inc", 64, 256, True), ("down1", 128, 128, True), ("down2", 256, 64, True), ("down3", 512, 32, True),
          ("down4", 512, 16, False), ("up1", 256, 32, False), ("up2", 128, 64, False), ("up3", 64, 128, False),
          ("up4", 64, 256, False)]


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)


tot = [0.0, 0.0, 0.0]
for name, C, R, pool in LAYERS:
    z = torch.randn(B, R, R, C, device=dev).bfloat16()
    dA = torch.randn(B, R, R, C, device=dev).bfloat16()
    dP = torch.randn(B, R // 2, R // 2, C, device=dev).bfloat16() if pool else None
    coef = [torch.rand(C, device=dev) + 0.5 for _ in range(5)]
    small = torch.empty(3, C, device=dev)
    dz = torch.empty_like(z)
    out = torch.empty_like(z)
    n = z.numel()
    t_b = timed(lambda: ops.bn_relu_bwd(z, dA, dP, coef[0], coef[1], coef[2], coef[3], coef[4], small[:2], small[2], dz=dz))
    by_b = n * 2 * 5 + (dP.numel() * 2 * 2 if pool else 0)
    pooled = torch.empty(B, R // 2, R // 2, C, dtype=torch.bfloat16, device=dev) if pool else None
    t_a = timed(lambda: ops.bn_relu_apply(z, coef[0], coef[1], pool=pool, out=out, pooled=pooled))
    by_a = n * 2 * 2 + (pooled.numel() * 2 if pool else 0)
    tot[0] += t_b; tot[1] += t_a
    print(f"{name:6s} C{C:4d} @{R:3d} pool={int(pool)} | bwd (2 passes) {t_b:7.1f} us {by_b / t_b / 1e6:5.2f} TB/s | "
          f"apply {t_a:6.1f} us {by_a / t_a / 1e6:5.2f} TB/s", flush=True)
print(f"sum bwd {tot[0]:.1f} us, apply {tot[1]:.1f} us (conv2-type layers only)")
