"""Is an issue gap between groups of 24 tcgen05.mma hidden by queued MMAs?  (ugp_bench_mma, a_shift = gap << 4)"""
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
iters = 2400
for n in (64, 128):
    for gap in (0, 1, 2, 4, 8, 16):
        out.zero_()
        _lib.check(lib.ugp_bench_mma(n, 0, iters, 2, 2, gap << 4, 1, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench")
        torch.cuda.synchronize()
        cyc = out[0].item()
        print(f"N={n} gap={gap:2d}: {cyc / (iters * 4):6.1f} cycles/MMA, {cyc / (iters / 6):8.0f} cycles per 24-MMA group (MMA floor {24 * max(n // 2, 32 + n // 4)})", flush=True)
