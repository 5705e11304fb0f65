"""Model-level parity on the GPU: kernel-backed PGUNet{1..4} + fused loss vs the oracle restatement
evaluated in fp32 on the same device (TF32 off).  Prints relative L2 errors for logits, loss,
uncertainty map, BN running stats and every parameter gradient.

    python tools/gpu_check_model.py [stage ...]
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch

import ug_oracle as O
import ug_pg_unet_b200 as U

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def check_stage(stage: int, B: int = 2, nc: int = 1):
    dev = "cuda"
    R = O.STAGE_RESOLUTION[stage]
    g = torch.Generator().manual_seed(100 + stage)
    x = torch.randn(B, 3, R, R, generator=g).to(dev)
    y = torch.randint(0, 2, (B, nc, R, R), generator=g).float().to(dev)
    sd = {k: v.to(dev) for k, v in O.seeded_state_dict(stage, 3, nc, seed=stage).items()}
    sd_prev = None
    model = U.STAGE_MODELS[stage](3, nc).to(dev)
    model.load_state_dict(sd)
    prev = None
    if stage > 1:
        sd_prev = {k: v.to(dev) for k, v in O.seeded_state_dict(stage - 1, 3, nc, seed=stage - 1).items()}
        prev = U.STAGE_MODELS[stage - 1](3, nc).to(dev)
        prev.load_state_dict(sd_prev)

    ref = O.train_step(sd, x, y, stage, sd_prev, alpha=1.0, pos_weight=5.0)
    with O.bf16_storage(True):
        emu = O.train_step(sd, x, y, stage, sd_prev, alpha=1.0, pos_weight=5.0)

    model.train()
    ugl = U.UncertaintyGuidedLoss(dev)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([5.0], device=dev), reduction="none")
    logits = model(x)
    A = None
    if stage > 1:
        A = ugl.generate_uncertainty_map(x, prev, O.STAGE_RESOLUTION[stage - 1], R)
    final, base = ugl.apply_uncertainty_weighted_loss(crit, logits, y, A, 1.0)
    final.backward()
    torch.cuda.synchronize()

    worst = 0.0
    rows = []
    e = rel(logits, ref["logits"])
    rows.append(("logits", e))
    rows.append(("final_loss", abs(final.item() - ref["final_loss"])))
    rows.append(("base_loss", abs(base - ref["base_loss"])))
    if A is not None:
        rows.append(("uncertainty max|d|", (A - ref["uncertainty"]).abs().max().item()))
    # argmax mask agreement away from the decision boundary
    refm = ref["logits"] > 0
    mine = logits > 0
    safe = ref["logits"].abs() > 0.05 * ref["logits"].abs().max()
    rows.append(("mask mismatches (|z_ref| > 5% of max)", float((refm != mine)[safe].sum().item())))
    rows.append(("mask mismatches (all)", float((refm != mine).sum().item()) / refm.numel()))
    gerrs = []
    for k, p in model.named_parameters():
        gr = ref["grads"][k]
        if p.grad is None:
            rows.append((f"grad {k}", float("nan")))
            continue
        if k.endswith("conv_op.0.bias") or k.endswith("conv_op.3.bias"):
            # conv bias before train-mode BN: gradient is rounding noise around 0
            scale = ref["grads"][k.replace(".bias", ".weight")].abs().max().item()
            gerrs.append((k, (p.grad - gr).abs().max().item() / (scale + 1e-30), True))
        else:
            gerrs.append((k, rel(p.grad, gr), False))
    for k, e, is_bias in gerrs:
        worst = max(worst, e)
    gerrs.sort(key=lambda t: -t[1])
    for k, e, is_bias in gerrs[:6]:
        rows.append((f"grad {k}{' (abs/|dW|max)' if is_bias else ''}", e))
    rows.append(("grad worst", worst))
    bn_worst = 0.0
    for k, v in ref["bn"].items():
        mine_v = model.state_dict()[k]
        if k.endswith("num_batches_tracked"):
            bn_worst = max(bn_worst, float(abs(int(mine_v) - int(v))))
        else:
            bn_worst = max(bn_worst, rel(mine_v, v))
    rows.append(("bn running stats worst", bn_worst))
    # eval-mode forward
    model.eval()
    with torch.no_grad():
        le = model(x)
        refe = O.pgunet_forward({k: v for k, v in model.state_dict().items()}, x, stage, training=False)
    rows.append(("eval logits", rel(le, refe)))
    # same comparison against the bf16-storage emulation of the oracle (isolates kernel bugs)
    rows.append(("[emu] logits", rel(logits, emu["logits"])))
    rows.append(("[emu] final_loss", abs(final.item() - emu["final_loss"])))
    rows.append(("[emu-vs-fp32 oracle] logits", rel(emu["logits"], ref["logits"])))
    ge = []
    for k, p in model.named_parameters():
        if k.endswith("conv_op.0.bias") or k.endswith("conv_op.3.bias") or p.grad is None:
            continue
        ge.append((k, rel(p.grad, emu["grads"][k]), rel(emu["grads"][k], ref["grads"][k])))
    for k, e1, e2 in ge:
        if k.endswith(".weight") and ("conv_op.0" in k or "conv_op.3" in k or "outc" in k):
            rows.append((f"[emu] grad {k}  (emu-vs-fp32: {e2:.2e})", e1))
    rows.append(("[emu] grad worst", max(e for _, e, _ in ge)))
    print(f"== stage {stage}  B={B} nc={nc} R={R}")
    for name, v in rows:
        print(f"   {name:75s} {v:.3e}")
    return rows


if __name__ == "__main__":
    stages = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    for s in stages:
        check_stage(s, B=2, nc=1)
    check_stage(1, B=4, nc=2)
