"""Runs one convolution configuration a few times (for `ncu --set full -k regex:conv_`).

    python tools/profile_conv.py B H W c0 c1 cout [halo=1] [iters=3] [kind=fwd|wgrad]
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ug_pg_unet_b200 import _lib, ops


def main():
    a = sys.argv[1:]
    B, H, W, c0, c1, cout = [int(v) for v in a[:6]]
    halo = int(a[6]) if len(a) > 6 else 1
    iters = int(a[7]) if len(a) > 7 else 3
    kind = a[8] if len(a) > 8 else "fwd"
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, device="cuda", generator=g) / 30
    wf, _ = ops.pack_weights(w)
    _lib.set_option("conv_halo", halo)
    if kind == "fwd":
        out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        st = ops.new_stats(cout, "cuda")
        bias = torch.zeros(cout, device="cuda")
        for _ in range(iters):
            ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, stats=st, out0=out)
    else:
        dz = torch.randn(B, H, W, cout, device="cuda", generator=g).bfloat16()
        dwp = torch.zeros(9 * (c0 + c1) * cout, device="cuda")
        for _ in range(iters):
            ops.conv_wgrad(x0, x1, dz, 9, dwp)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
