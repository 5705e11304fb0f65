"""Validates and times the halo-staged convolution (csrc/conv_halo.cu) against the per-tap kernel
and torch, for both descriptor base-offset modes.

    python tools/gpu_check_halo.py
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from ug_pg_unet_b200 import _lib, ops

torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nchw(x):
    return x.permute(0, 3, 1, 2).float()


def time_it(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    cases = [  # B, H, W, c0, c1, cout, split
        (1, 4, 128, 64, 0, 64, None),
        (2, 128, 128, 64, 0, 64, None),
        (1, 256, 256, 64, 64, 64, None),
        (2, 128, 128, 128, 0, 128, None),
        (1, 128, 128, 64, 0, 256, 128),
        (1, 256, 256, 64, 0, 128, 64),
    ]
    ok_modes = {}
    for mode in (1,):
        all_ok = True
        for (B, H, W, c0, c1, cout, split) in cases:
            x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
            x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
            cin = c0 + c1
            w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
            bias = torch.randn(cout, device="cuda", generator=g)
            wf, _ = ops.pack_weights(w)
            _lib.set_option("conv_halo", 0)
            st0 = ops.new_stats(cout, "cuda")
            base = ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, stats=st0, split_n=split)
            _lib.set_option("conv_halo", mode)
            st1 = ops.new_stats(cout, "cuda")
            try:
                out = ops.conv_gemm(x0, x1, wf, 9, cout, bias=bias, stats=st1, split_n=split)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"mode {mode} case {(B, H, W, c0, c1, cout, split)}: EXCEPTION {e}")
                all_ok = False
                break
            if split:
                base_c = torch.cat([base[0], base[1]], 3)
                out_c = torch.cat([out[0], out[1]], 3)
            else:
                base_c, out_c = base, out
            xin = nchw(x0) if x1 is None else torch.cat([nchw(x0), nchw(x1)], 1)
            ref = F.conv2d(xin, w.bfloat16().float(), bias, padding=1)
            e_ref = rel(nchw(out_c), ref)
            e_base = rel(out_c, base_c)
            e_st = rel(st1.sum(0), st0.sum(0))
            good = e_ref < 6e-3 and e_st < 1e-4
            all_ok &= good
            print(f"[{'PASS' if good else 'FAIL'}] mode {mode} B{B} {H}x{W} {c0}+{c1}->{cout} split={split}: "
                  f"vs torch {e_ref:.2e}, vs per-tap kernel {e_base:.2e}, stats {e_st:.1e}", flush=True)
        ok_modes[mode] = all_ok
    print("modes ok:", ok_modes)
    good_mode = next((m for m, v in ok_modes.items() if v), 0)
    # timing at the stage-4 batch-16 shapes
    shapes = [(16, 256, 256, 64, 0, 64), (16, 256, 256, 64, 64, 64), (16, 128, 128, 64, 0, 128),
              (16, 128, 128, 128, 0, 128), (16, 128, 128, 128, 128, 64), (16, 128, 128, 64, 0, 64)]
    for (B, H, W, c0, c1, cout) in shapes:
        x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
        x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
        w = torch.randn(cout, c0 + c1, 3, 3, device="cuda", generator=g) / 30
        wf, _ = ops.pack_weights(w)
        out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
        st = ops.new_stats(cout, "cuda")
        flops = 2.0 * B * H * W * cout * 9 * (c0 + c1)
        res = []
        for mode in (0, good_mode) if good_mode else (0,):
            _lib.set_option("conv_halo", mode)
            us = time_it(lambda: ops.conv_gemm(x0, x1, wf, 9, cout, stats=st, out0=out))
            res.append(f"mode {mode}: {us:7.1f} us {flops / us / 1e6:7.1f} TFLOP/s")
        print(f"B{B} {H}x{W} {c0}+{c1}->{cout}:  " + "   ".join(res), flush=True)
    _lib.set_option("conv_halo", 0)


if __name__ == "__main__":
    main()
