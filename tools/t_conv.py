"""Per-layer timing of the tcgen05 convolutions over the stage-4 shapes (CUDA events, L2 flushed).

    python tools/t_conv.py [--batch 16] [--opt conv_pair=0] [--only fwd|dgrad|wgrad]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import ops, _lib

LAYERS = [  # name, c0, c1, cout, res
    ("inc.3", 64, 0, 64, 256), ("down1.0", 64, 0, 128, 128), ("down1.3", 128, 0, 128, 128),
    ("down2.0", 128, 0, 256, 64), ("down2.3", 256, 0, 256, 64), ("down3.0", 256, 0, 512, 32),
    ("down3.3", 512, 0, 512, 32), ("down4.0", 512, 0, 512, 16), ("down4.3", 512, 0, 512, 16),
    ("up1.0", 512, 512, 256, 32), ("up1.3", 256, 0, 256, 32), ("up2.0", 256, 256, 128, 64),
    ("up2.3", 128, 0, 128, 64), ("up3.0", 128, 128, 64, 128), ("up3.3", 64, 0, 64, 128),
    ("up4.0", 64, 64, 64, 256), ("up4.3", 64, 0, 64, 256),
]


def timed(fn, flush, reps=3):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--only", default="")
    ap.add_argument("--layers", default="")
    args = ap.parse_args()
    for o in args.opt:
        k, v = o.split("=")
        _lib.set_option(k, int(v))
    dev = torch.device("cuda", 0)
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = {"fwd": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    for name, c0, c1, cout, R in LAYERS:
        if args.layers and name not in args.layers.split(","):
            continue
        cin = c0 + c1
        x0 = torch.randn(B, R, R, c0, device=dev).bfloat16()
        x1 = torch.randn(B, R, R, c1, device=dev).bfloat16() if c1 else None
        w = torch.randn(cout, cin, 3, 3, device=dev) / (3.0 * cin ** 0.5)
        wf, wd = ops.pack_weights(w)
        stats = ops.new_stats(cout, "cuda")
        dz = torch.randn(B, R, R, cout, device=dev).bfloat16()
        dwp = torch.zeros(9 * cin * cout, device=dev)
        out = torch.empty(B, R, R, cout, dtype=torch.bfloat16, device=dev)
        gf = 2.0 * B * R * R * cout * 9 * cin / 1e9
        line = f"{name:8s} {cin:5d}->{cout:3d} @{R:3d}  {gf:6.1f} GF "
        if args.only in ("", "fwd"):
            t = timed(lambda: ops.conv_gemm(x0, x1, wf, 9, cout, stats=stats, out0=out), flush)
            tot["fwd"][0] += gf; tot["fwd"][1] += t
            line += f"| fwd {t:7.1f} us {gf / t * 1e3:6.0f} TF/s "
        if args.only in ("", "dgrad"):
            sp = c0 if c1 else None
            t = timed(lambda: ops.conv_gemm(dz, None, wd, 9, cin, split_n=sp), flush)
            tot["dgrad"][0] += gf; tot["dgrad"][1] += t
            line += f"| dgrad {t:7.1f} us {gf / t * 1e3:6.0f} TF/s "
        if args.only in ("", "wgrad"):
            t = timed(lambda: ops.conv_wgrad(x0, x1, dz, 9, dwp), flush)
            tot["wgrad"][0] += gf; tot["wgrad"][1] += t
            line += f"| wgrad {t:7.1f} us {gf / t * 1e3:6.0f} TF/s"
        print(line, flush=True)
    for k, (g, t) in tot.items():
        if t:
            print(f"total {k}: {g:.1f} GF in {t:.1f} us = {g / t * 1e3:.0f} TF/s")


if __name__ == "__main__":
    main()
