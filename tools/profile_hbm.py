"""Launches the HBM-bound kernels of the stage-4 step at their real batch-16 shapes, once each after a
warm-up, for an `ncu --set full` capture (profiles/r02_hbm_*.txt):

    ncu --set full --clock-control none --import-source on -k regex:'upsample|bn_relu|conv_first' -o gpurun_out/hbm python tools/profile_hbm.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ug_pg_unet_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B = 16
which = sys.argv[1:] or ["up", "bn", "first"]


def act(b, h, w, c):
    return torch.randn(b, h, w, c, device=dev, generator=g).bfloat16()


def flush():
    torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()


for rep in range(2):
    if "up" in which:
        for (h, c) in ((128, 64), (64, 128), (32, 256)):
            x = act(B, h, h, c)
            flush()
            up = ops.upsample2x(x)
            flush()
            ops.upsample2x_bwd(up)
    if "bn" in which or "bnpool" in which:
        shapes = ((256, 64, False), (128, 64, False), (64, 128, False), (32, 256, False), (16, 512, False),
                  (128, 128, True), (256, 64, True))
        if "bn" not in which:
            shapes = tuple(s for s in shapes if s[2]) + ((256, 64, False),)
        for (h, c, pool) in shapes:
            z, dA = act(B, h, h, c), act(B, h, h, c)
            dP = act(B, h // 2, h // 2, c) if pool else None
            coef = torch.rand(4, c, device=dev, generator=g) + 0.5
            small = torch.empty(3, c, device=dev)
            flush()
            ops.bn_relu_bwd(z, dA, dP, coef[0], coef[1] - 1.0, coef[2], coef[3], torch.ones(c, device=dev), small[:2], small[2])
            flush()
            ops.bn_relu_apply(z, coef[0], coef[1] - 1.0, pool=pool)
    if "first" in which:       # fused first layer (conv_first.cu): forward (train statistics) and weight gradient
        x = torch.randn(B, 3, 256, 256, device=dev, generator=g)
        w = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
        wf, _ = ops.pack_weights(w, need_dgrad=False)
        rows = ops.first_rows(x)
        st = ops.new_stats(64, dev)
        flush()
        zf = ops.conv_gemm(rows, None, wf, 1, 64, bias=torch.zeros(64, device=dev), stats=st)
        flush()
        ops.conv_wgrad(rows, None, zf, 1, torch.zeros(64 * 64, device=dev))
    torch.cuda.synchronize()
print("profile_hbm done")
