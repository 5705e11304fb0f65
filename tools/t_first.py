"""Fused first layer (ugp_conv_first_fwd / _wgrad): stand-alone times at the headline shape, with the what-if
switches of conv_debug (1 = no TMA store, 4 = no image loads) to see which role bounds the kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import ops
from ug_pg_unet_b200._lib import set_option

dev = torch.device("cuda")
B, H, W = 16, 256, 256
x = torch.randn(B, 3, H, W, device=dev)
w = torch.randn(64, 3, 3, 3, device=dev) * 0.2
bias = torch.randn(64, device=dev)
wf, _ = ops.pack_weights(w, need_dgrad=False)
rows = ops.first_rows(x)
col = rows.materialize()
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
dz = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
dw = torch.zeros(64 * 64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


st = ops.new_stats(64, dev)
for dbg in (0, 1, 4, 5):
    set_option("conv_debug", dbg)
    a = t(lambda: ops.conv_gemm(rows, None, wf, 1, 64, bias=bias, stats=st))
    b = t(lambda: ops.conv_gemm(rows, None, wf, 1, 64, bias=bias, scale=sc, shift=sh, relu=True))
    c = t(lambda: ops.conv_gemm(rows, None, wf, 1, 64))
    print(f"conv_debug={dbg}: fused fwd train(stats) {a:6.1f} us   eval(affine+relu) {b:6.1f} us   plain {c:6.1f} us", flush=True)
set_option("conv_debug", 0)
print(f"packed rows + GEMM: im2col {t(lambda: ops.im2col_first(x)):6.1f} us, gemm train {t(lambda: ops.conv_gemm(col, None, wf, 1, 64, bias=bias, stats=st)):6.1f} us")
print(f"wgrad fused {t(lambda: ops.conv_wgrad(rows, None, dz, 1, dw)):6.1f} us   packed {t(lambda: ops.conv_wgrad(col, None, dz, 1, dw)):6.1f} us")
x2 = torch.randn(B, 3, 256, 256, device=dev)
w2 = torch.randn(128, 3, 3, 3, device=dev) * 0.2
wf2, _ = ops.pack_weights(w2, need_dgrad=False)
r2 = ops.first_rows(x2, 128, 128, cout=128)
c2 = r2.materialize()
sc2, sh2 = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev)
print(f"stage-3 twin (resize 256->128, cout 128, eval): fused {t(lambda: ops.conv_gemm(r2, None, wf2, 1, 128, scale=sc2, shift=sh2, relu=True)):6.1f} us   "
      f"packed {t(lambda: ops.im2col_first(x2, 128, 128)):6.1f} + {t(lambda: ops.conv_gemm(c2, None, wf2, 1, 128, scale=sc2, shift=sh2, relu=True)):6.1f} us")
