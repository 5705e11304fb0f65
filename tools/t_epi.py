"""What-if timings of the halo conv kernel (debug switches; results are intentionally wrong)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import _lib, ops
from tools.gpu_check_halo import time_it
g = torch.Generator(device="cuda").manual_seed(0)
for (B,H,W,c0,c1,cout) in [(16,256,256,64,0,64),(16,256,256,64,64,64),(16,128,128,128,0,128)]:
    x0 = torch.randn(B,H,W,c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B,H,W,c1, device="cuda", generator=g).bfloat16() if c1 else None
    w = torch.randn(cout, c0+c1, 3,3, device="cuda", generator=g)/30
    wf,_ = ops.pack_weights(w)
    out = torch.empty(B,H,W,cout, device="cuda", dtype=torch.bfloat16)
    st = ops.new_stats(cout, "cuda"); bias = torch.zeros(cout, device="cuda")
    flops = 2.0*B*H*W*cout*9*(c0+c1)
    _lib.set_option("conv_halo", 1)
    for dbg, name in ((0, "full"), (1, "no MMA"), (2, "no epilogue"), (3, "loads only")):
        _lib.set_option("conv_debug", dbg)
        us = time_it(lambda: ops.conv_gemm(x0, x1, wf, 9, cout, out0=out, bias=bias))
        print(f"{(B,H,W,c0,c1,cout)} {name:12s} {us:7.1f} us {flops/us/1e6:7.1f} TF/s", flush=True)
    _lib.set_option("conv_debug", 0)
