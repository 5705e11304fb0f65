"""What-if timings of the wgrad kernel (debug switches; results are intentionally wrong)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import _lib, ops
from tools.gpu_check_halo import time_it
g = torch.Generator(device="cuda").manual_seed(0)
for (B,H,W,c0,c1,cout) in [(16,256,256,64,0,64),(16,256,256,64,64,64),(16,128,128,128,0,128),(16,64,64,256,0,256),(16,32,32,512,0,512)]:
    x0 = torch.randn(B,H,W,c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B,H,W,c1, device="cuda", generator=g).bfloat16() if c1 else None
    dz = torch.randn(B,H,W,cout, device="cuda", generator=g).bfloat16()
    dwp = torch.zeros(9*(c0+c1)*cout, device="cuda")
    flops = 2.0*B*H*W*cout*9*(c0+c1)
    for dbg, name in ((0, "full"), (1, "no MMA"), (2, "no reds"), (3, "loads only")):
        _lib.set_option("conv_debug", dbg)
        us = time_it(lambda: ops.conv_wgrad(x0, x1, dz, 9, dwp))
        print(f"{(B,H,W,c0,c1,cout)} {name:12s} {us:7.1f} us {flops/us/1e6:7.1f} TF/s", flush=True)
    _lib.set_option("conv_debug", 0)
