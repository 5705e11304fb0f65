"""tcgen05.mma cost model: cycles per MMA vs N, A-window alignment, CTA pairs (ugp_bench_mma)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import _lib

import ctypes
from ug_pg_unet_b200 import build as _build
lib = ctypes.CDLL(str(_build.build_microbench()))
lib.ugp_bench_mma.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p]
lib.ugp_bench_mma.restype = ctypes.c_int
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 2000
print("   N pair nacc shift ctas | cycles per MMA (min / max over CTAs) | math floor")
for ctas in (1, 148):
    for pair in (0, 1):
        for n in (64, 128, 256):
            for nacc, shift in ((1, 0), (2, 0), (2, 1), (2, 7), (2, 8), (2, 9)):
                if nacc * n > 512:
                    continue
                out.zero_()
                _lib.check(lib.ugp_bench_mma(n, pair, iters, nacc, nacc, shift, ctas, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "bench")
                torch.cuda.synchronize()
                v = out[out > 0].double() / (iters * 4)
                print(f"{n:4d} {pair:4d} {nacc:4d} {shift:5d} {ctas:4d} | {v.min().item():8.1f} {v.max().item():8.1f}   | {n / 2:5.0f}", flush=True)
