"""Where does run-to-run noise enter the forward?  Compares per-layer tensors of two identical runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ug_pg_unet_b200 as U
from ug_pg_unet_b200 import ops, engine

dev = torch.device("cuda", 0)
torch.manual_seed(100)
m = U.PGUNet2(3, 1).to(dev).train()
g = torch.Generator().manual_seed(7)
x = torch.randn(4, 3, 64, 64, generator=g).to(dev)

def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()

# unit level: first CBR of inc
dc = m.inc.conv
col = ops.im2col_first(x)
res = []
for i in range(3):
    wf, _ = dc._packs[0].get()
    stats = ops.new_stats(256, "cuda")
    z = ops.conv_gemm(col, None, wf, 1, 256, bias=dc.conv_op[0].bias.detach(), stats=stats)
    torch.cuda.synchronize()
    res.append((z.clone(), stats.sum(0)))
print("conv1 z rel", rel(res[1][0], res[0][0]), "stats rel", rel(res[1][1], res[0][1]),
      "max abs stats diff", (res[1][1] - res[0][1]).abs().max().item())
with torch.no_grad():
    f = [[t.clone() for t in m.encode(x)] for _ in range(2)]
for i, (a, b) in enumerate(zip(*f)):
    print("encode feat", i, tuple(a.shape), "rel", rel(a, b), "n diff", (a != b).sum().item(), "of", a.numel())
