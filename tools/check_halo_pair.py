import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
from ug_pg_unet_b200 import _lib, ops
g = torch.Generator(device="cuda").manual_seed(0)
ok = True
for (B, H, W, c0, c1, cout, split) in [(1, 4, 128, 64, 0, 64, None), (2, 128, 128, 64, 64, 128, None), (1, 256, 256, 64, 0, 64, None),
                                (2, 128, 128, 128, 0, 128, 64), (1, 256, 256, 64, 0, 128, 64), (3, 6, 128, 64, 0, 256, None)]:
    x0 = torch.randn(B, H, W, c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B, H, W, c1, device="cuda", generator=g).bfloat16() if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, device="cuda", generator=g) / 30
    wf, _ = ops.pack_weights(w)
    outs = []
    for mode in (1, 3):
        _lib.set_option("conv_pair", mode)
        st = ops.new_stats(cout, "cuda")
        r = ops.conv_gemm(x0, x1, wf, 9, cout, stats=st, split_n=split)
        if split:
            r = torch.cat([r[0], r[1]], dim=3)
        torch.cuda.synchronize()
        outs.append((r.float(), st.sum(0)))
    e = ((outs[1][0] - outs[0][0]).norm() / outs[0][0].norm()).item()
    es = ((outs[1][1] - outs[0][1]).norm() / outs[0][1].norm()).item()
    print((B, H, W, c0, c1, cout, split), "rel", e, "stats rel", es, flush=True)
    ok &= e < 1e-4 and es < 1e-5
print("PAIR HALO OK" if ok else "PAIR HALO FAIL")
