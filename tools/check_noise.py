"""Run-to-run reproducibility of one stage-2 trainer forward/backward (same weights, same batch)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ug_pg_unet_b200 as U

from ug_pg_unet_b200 import _lib
for o in filter(None, os.environ.get("UGP_OPTS", "").split(",")):
    k, v = o.split("=")
    _lib.set_option(k, int(v))
dev = torch.device("cuda", 0)
torch.manual_seed(100)
with contextlib.redirect_stdout(io.StringIO()):
    tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
st = 2
tr.current_stage, tr.current_model = st, tr.models[st]
tr.setup_optimizer(st)
tr.current_model.train(); tr.models[1].eval()
g = torch.Generator().manual_seed(7)
x = torch.randn(4, 3, 64, 64, generator=g).to(dev)
y = torch.randint(0, 2, (4, 1, 64, 64), generator=g).float().to(dev)
arena = tr._grad_arena
outs = []
for i in range(3):
    arena.zero_()
    tr.optimizer.zero_grad(set_to_none=True)
    loss, logits, unc, _ = tr._forward_loss(x, y, st)
    loss.backward()
    torch.cuda.synchronize()
    outs.append((arena.clone(), logits.detach().clone(), loss.item()))
for i in (1, 2):
    a, b = outs[0], outs[i]
    print(f"[{os.environ.get('UGP_OPTS', '')}] run {i} vs 0: arena rel {((a[0]-b[0]).norm()/a[0].norm()).item():.3e} logits rel "
          f"{((a[1]-b[1]).norm()/a[1].norm()).item():.3e} loss {a[2]:.8f} {b[2]:.8f}")
