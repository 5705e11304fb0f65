"""Times ugp_rmsprop_conv_packed over the stage-4 conv shapes (CUDA events, L2 flushed between runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import ops

dev = torch.device("cuda", 0)
shapes = [(64, 64), (128, 64), (128, 128), (256, 128), (256, 256), (512, 256), (512, 512), (512, 512), (512, 512),
          (256, 1024), (256, 256), (128, 512), (128, 128), (64, 256), (64, 64), (64, 128), (64, 64)]
W = [torch.randn(co, ci, 3, 3, device=dev) * 0.05 for co, ci in shapes]
V = [torch.zeros_like(w) for w in W]
D = [torch.randn(w.numel(), device=dev) for w in W]
WF = [torch.empty(9, co, ci, dtype=torch.bfloat16, device=dev) for co, ci in shapes]
WD = [torch.empty(9, ci, co, dtype=torch.bfloat16, device=dev) for co, ci in shapes]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n = sum(w.numel() for w in W)
from ug_pg_unet_b200 import _lib
for dbg in (0, 4, 8, 16, 4 + 8, 4 + 16, 8 + 16, 28):
  _lib.set_option("conv_debug", dbg)
  ts = []
  for it in range(4):
    for _ in range(12):
        flush.zero_()          # keeps the GPU busy while the host prepares the launch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.rmsprop_conv_packed(W, V, D, WF, WD, 1e-4, 0.99, 1e-8, 1e-4, 1.0)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
  print("dbg", dbg, "params", n, "us", [round(t, 1) for t in ts], "GB/s at 28 B/param", round(n * 28 / min(ts) / 1e3, 1))
