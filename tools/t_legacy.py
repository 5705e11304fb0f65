"""Timing of the ConvTranspose2d(k=2, s=2) kernels over the legacy UNet's four up-path shapes and of a
whole legacy-UNet training step (forward + mean-BCE backward), CUDA events, L2 flushed per kernel.

    python tools/t_legacy.py [--batch 16]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import ug_pg_unet_b200 as U
from ug_pg_unet_b200 import ops
from t_conv import timed

LAYERS = [("up_convolution_1.up", 1024, 512, 16), ("up_convolution_2.up", 512, 256, 32),
          ("up_convolution_3.up", 256, 128, 64), ("up_convolution_4.up", 128, 64, 128)]   # name, cin, cout, input res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, cin, cout, R in LAYERS:
        x = torch.randn(B, R, R, cin, device=dev).bfloat16()
        w = torch.randn(cin, cout, 2, 2, device=dev) / cin ** 0.5
        b = torch.randn(cout, device=dev)
        wf, wd = ops.pack_weights_tconv(w)
        dout = torch.randn(B, 2 * R, 2 * R, cout, device=dev).bfloat16()
        out = torch.empty_like(dout)
        dwp = torch.zeros(4 * cin * cout, device=dev)
        gf = 2.0 * B * R * R * cin * 4 * cout / 1e9
        mb = (x.numel() + out.numel()) * 2 / 1e6
        tf = timed(lambda: ops.conv_transpose2x2_fwd(x, wf, cout, bias=b, out=out), flush)
        td = timed(lambda: ops.conv_transpose2x2_dgrad(dout, wd, cin), flush)
        tw = timed(lambda: ops.conv_transpose2x2_wgrad(x, dout, dwp), flush)
        print(f"{name:20s} {cin:4d}->{cout:3d} @{R:3d}->{2 * R:3d}  {gf:5.1f} GF {mb:6.1f} MB | fwd {tf:6.1f} us "
              f"{gf / tf * 1e3:5.0f} TF/s {mb / tf * 1e3:5.0f} GB/s | dgrad {td:6.1f} us {gf / td * 1e3:5.0f} TF/s | "
              f"wgrad {tw:6.1f} us {gf / tw * 1e3:5.0f} TF/s", flush=True)
    torch.manual_seed(0)
    net = U.UNet(3, 1).to(dev).train()
    x = torch.randn(B, 3, 256, 256, device=dev)
    y = torch.randint(0, 2, (B, 1, 256, 256), device=dev).float()

    def step():
        net.zero_grad(set_to_none=True)
        F.binary_cross_entropy_with_logits(net(x), y).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(f"legacy UNet(3,1) fwd+bwd, batch {B}, 256x256 (eager launches): {ms:.2f} ms/step, {B / ms * 1e3:.0f} images/s")


if __name__ == "__main__":
    main()
