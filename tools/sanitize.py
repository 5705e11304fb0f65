"""One small launch of every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
    compute-sanitizer --tool synccheck python tools/sanitize.py

Shapes are the smallest that select each kernel variant (per-tap GEMM at N = 64 / 128 / 256, CTA-pair GEMM,
halo kernels single / pair / resident-weight, both wgrad kernels, ConvTranspose2d views, the element-wise
and reduction kernels, the fused optimiser and the fused Herlev head).  Summaries are kept under
profiles/r02_sanitizer_*.txt.  A selection can be given on the command line (family names).

    python tools/sanitize.py --guards          # our own out-of-bounds-write check (guard bands), see below
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ug_pg_unet_b200 import ops

bf16 = torch.bfloat16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def act(B, H, W, C):
    return torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()


def conv_family(B, H, W, c0, c1, cout, wgrad=True):
    x0 = act(B, H, W, c0)
    x1 = act(B, H, W, c1) if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, device=dev, generator=g) / 30
    wf, wd = ops.pack_weights(w)
    st = ops.new_stats(cout, dev)
    z = ops.conv_gemm(x0, x1, wf, 9, cout, bias=torch.zeros(cout, device=dev), stats=st)
    dz = act(B, H, W, cout)
    if c1:
        ops.conv_gemm(dz, None, wd, 9, c0 + c1, split_n=c0)
    else:
        ops.conv_gemm(dz, None, wd, 9, c0)
    if wgrad:
        dwp = torch.zeros(9 * (c0 + c1) * cout, device=dev)
        ops.conv_wgrad(x0, x1, dz, 9, dwp)
        ops.unpack_wgrad(dwp, cout, c0 + c1)
    return z


FAMILIES = {}


def family(fn):
    FAMILIES[fn.__name__] = fn
    return fn


@family
def gemm64():       # conv_gemm_kernel<64>, conv_wgrad_kernel
    conv_family(1, 16, 16, 64, 0, 64)


@family
def gemm128():      # conv_gemm_kernel<128>
    conv_family(1, 32, 32, 128, 0, 128)


@family
def gemm256_pair():  # conv_pair_kernel (cta_group::2) + conv_gemm_kernel<256>
    conv_family(2, 16, 16, 64, 64, 256)
    conv_family(1, 8, 16, 64, 0, 256)          # one m tile: no pair


@family
def halo64():       # conv_halo_kernel<64> resident weights, conv_wgrad_halo_kernel
    conv_family(1, 4, 128, 64, 0, 64)


@family
def halo64_pair():  # conv_halo_pair_kernel<64> (two chunks per tile)
    conv_family(1, 4, 128, 64, 64, 64)


@family
def halo128():      # conv_halo_kernel<128> / conv_halo_pair_kernel<128>
    conv_family(1, 4, 128, 64, 0, 128)
    conv_family(1, 4, 128, 128, 0, 128)


@family
def general_size():  # masked accumulator rows, 4-D stores
    conv_family(1, 28, 28, 64, 0, 64)


@family
def tconv():
    x = act(1, 16, 16, 128)
    w = torch.randn(128, 64, 2, 2, device=dev, generator=g) / 10
    wf, wd = ops.pack_weights_tconv(w)
    y = ops.conv_transpose2x2_fwd(x, wf, 64, bias=torch.zeros(64, device=dev))
    ops.conv_transpose2x2_dgrad(y, wd, 128)
    dwp = torch.zeros(4 * 128 * 64, device=dev)
    ops.conv_transpose2x2_wgrad(x, y, dwp)
    ops.unpack_wgrad_tconv(dwp, 128, 64)
    ops.channel_sum(y)


@family
def elementwise():
    z = act(2, 16, 16, 64)
    C = 64
    coef = torch.rand(4, C, device=dev, generator=g) + 0.5
    a, p = ops.bn_relu_apply(z, coef[0], coef[1], pool=True)
    ops.bn_relu_apply(z, coef[0], coef[1])
    small = torch.empty(3, C, device=dev)
    gamma = torch.ones(C, device=dev)
    dA, dP = act(2, 16, 16, C), act(2, 8, 8, C)
    ops.bn_relu_bwd(z, dA, None, coef[0], coef[1], coef[2], coef[3], gamma, small[:2], small[2])
    ops.bn_relu_bwd(z, dA, dP, coef[0], coef[1], coef[2], coef[3], gamma, small[:2], small[2])
    ops.bn_relu_bwd(z, None, dP, coef[0], coef[1], coef[2], coef[3], gamma, small[:2], small[2])
    st = ops.new_stats(C, dev)
    rm, rv, nb = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    ops.bn_finalize(st + 1.0, 512, gamma, torch.zeros(C, device=dev), 1e-5, 0.1, rm, rv, nb, coef[0], coef[1], coef[2], coef[3])
    ops.maxpool2x2(z, return_indices=True)
    ops.maxpool2x2_bwd(z, dP)
    up = ops.upsample2x(z)
    ops.upsample2x_bwd(up)
    img = torch.randn(2, 3, 32, 32, device=dev, generator=g)
    ops.im2col_first(img)
    ops.im2col_first(img, 16, 16)
    ops.im2col_first_u8(torch.randint(0, 256, (2, 32, 32, 3), device=dev, dtype=torch.uint8, generator=g), 16, 16)
    ops.global_avg_pool(z)
    ops.global_avg_pool_bwd(torch.randn(2, C, device=dev, generator=g), 16, 16)


@family
def head_loss():
    x = act(2, 32, 32, 64)
    w = torch.randn(1, 64, device=dev, generator=g)
    b = torch.zeros(1, device=dev)
    l = ops.outconv_fwd(x, w, b)
    dx = torch.empty_like(x)
    ops.outconv_bwd(x, w, torch.randn_like(l), dx, False, torch.zeros(1, 64, device=dev), torch.zeros(1, device=dev))
    maps = [torch.randn(2, 1, 8 << i, 8 << i, device=dev, generator=g) for i in range(3)]
    lg = ops.head_combine_fwd(maps)
    for m in maps:
        ops.head_combine_bwd(lg, m.shape[-1])
    A = ops.uncertainty_margin(torch.randn(2, 1, 16, 16, device=dev, generator=g), 32, 32)
    y = torch.randint(0, 2, (2, 1, 32, 32), device=dev, generator=g).float()
    ops.bce_uncertainty_loss(lg, y, A, 1.0, 5.0, True, 1.0 / lg.numel())
    ops.bce_uncertainty_loss(lg, y.to(torch.uint8), A, 1.0, 5.0, True, 1.0 / lg.numel())
    z7 = torch.randn(4, 7, device=dev, generator=g)
    u = ops.softmax_entropy(z7)
    lab = torch.randint(0, 7, (4,), device=dev, generator=g)
    ops.ce_weighted_loss(z7, lab, u, 1.0, True, 0.25)
    ops.seg_confusion(lg, y, want_mask=True)
    ops.cls_eval(z7, lab)


@family
def optim():
    w = [torch.randn(64, 64, 3, 3, device=dev, generator=g), torch.randn(128, 64, 3, 3, device=dev, generator=g)]
    v = [torch.zeros_like(t) for t in w]
    gp = [torch.randn(t.numel(), device=dev, generator=g) for t in w]
    packs = [ops.pack_weights(t) for t in w]
    ops.rmsprop_conv_packed(w, v, gp, [p[0] for p in packs], [p[1] for p in packs], 1e-4, 0.99, 1e-8, 1e-4)
    ps = [torch.randn(n, device=dev, generator=g) for n in (64, 7, 1000)]
    ops.rmsprop_multi(ps, [torch.randn_like(t) for t in ps], [torch.zeros_like(t) for t in ps], 1e-4, 0.99, 1e-8, 1e-4)
    ops.adam_multi(ps, [torch.randn_like(t) for t in ps], [torch.zeros_like(t) for t in ps],
                   [torch.zeros_like(t) for t in ps], 1, 1e-3, 0.9, 0.999, 1e-8, 1e-4)


@family
def mlp_head():
    B = 4
    x = act(B, 8, 8, 512)
    w1, b1 = torch.randn(512, 512, device=dev, generator=g) / 20, torch.zeros(512, device=dev)
    w2, b2 = torch.randn(256, 512, device=dev, generator=g) / 20, torch.zeros(256, device=dev)
    w3, b3 = torch.randn(7, 256, device=dev, generator=g) / 20, torch.zeros(7, device=dev)
    u = torch.rand(B, 1280, device=dev, generator=g)
    logits, pooled, h1, h2 = ops.gap_mlp_head_fwd(x, w1, b1, w2, b2, w3, b3, u=u, p=(0.5, 0.3, 0.2))
    ops.mlp_head_bwd(torch.randn_like(logits), w1, w2, w3, pooled, h1, h2, u=u, p=(0.5, 0.3, 0.2))


# ----------------------------------------------------------------------------------------------------------
# Guard bands (`--guards`): compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_unavailable.txt),
# so out-of-bounds WRITES are hunted with canaries of our own: every tensor the op wrappers allocate for a
# kernel to write (torch.empty / empty_like / zeros inside ug_pg_unet_b200.ops) is carved out of a larger
# buffer whose 4 KiB margins hold a byte pattern; after every family the margins must be untouched.  Covers
# plain stores, vector reductions and TMA stores (tensor maps are built on the carved-out pointers).
# ----------------------------------------------------------------------------------------------------------
GUARD_BYTES = 4096
_guards = []


class _TorchWithGuards:
    """Stands in for the `torch` module inside ug_pg_unet_b200.ops: allocation functions add guard bands."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _carve(shape, dtype, device, zero):
        if isinstance(shape, int):
            shape = (shape,)
        n = 1
        for d in shape:
            n *= int(d)
        esz = torch.empty((), dtype=dtype).element_size()
        g = GUARD_BYTES // esz
        raw = torch.empty((n + 2 * g,), dtype=torch.uint8 if False else dtype, device=device)
        raw.view(torch.uint8).fill_(0xA5)
        body = raw[g:g + n].view(tuple(shape))
        if zero:
            body.zero_()
        _guards.append((raw, g, n))
        return body

    def empty(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        return self._carve(tuple(shape), dtype, device, False)

    def zeros(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        return self._carve(tuple(shape), dtype, device, True)

    def empty_like(self, t, **kw):
        return self._carve(tuple(t.shape), kw.get("dtype", t.dtype), t.device, False)


def check_guards(family: str) -> int:
    torch.cuda.synchronize()
    bad = 0
    for raw, g, n in _guards:
        b = raw.view(torch.uint8)
        esz = raw.element_size()
        lo, hi = b[: g * esz], b[(g + n) * esz:]
        if not (bool((lo == 0xA5).all()) and bool((hi == 0xA5).all())):
            bad += 1
            print(f"[GUARD] {family}: a kernel wrote outside a {n}-element {raw.dtype} buffer "
                  f"(below: {int((lo != 0xA5).sum())} bytes, above: {int((hi != 0xA5).sum())} bytes)", flush=True)
    count = len(_guards)
    _guards.clear()
    return bad, count


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    guards = "--guards" in sys.argv
    names = args or list(FAMILIES)
    if guards:
        ops.torch = _TorchWithGuards()
    bad = 0
    for n in names:
        FAMILIES[n]()
        torch.cuda.synchronize()
        if guards:
            b, count = check_guards(n)
            bad += b
            print(f"[ran] {n}: {count} guarded buffers, {b} violated", flush=True)
        else:
            print(f"[ran] {n}", flush=True)
    if guards:
        print(f"sanitize.py --guards: {'OK' if bad == 0 else 'FAILED'} ({bad} violated guard bands)", flush=True)
        if bad:
            raise SystemExit(1)
    else:
        print("sanitize.py: all families launched", flush=True)


if __name__ == "__main__":
    main()
