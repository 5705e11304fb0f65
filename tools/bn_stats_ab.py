"""Design A/B on the CPU oracle: train-mode BatchNorm statistics from the bf16-ROUNDED conv
outputs (what conv_epilogue.cuh computes) vs from the fp32 accumulators before rounding.  Both variants are
the oracle's bf16-storage emulation; distances are to the fp32 oracle (= the reference)."""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import ug_oracle as O


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


print("stage B | logits rel: rounded-stats / fp32-stats | median grad rel: rounded / fp32 | worst grad rel: rounded / fp32 | "
      "rounded-vs-fp32 variants apart (logits)")
for stage, B in ((1, 4), (2, 4), (3, 2), (4, 2)):
    R = O.STAGE_RESOLUTION[stage]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 3, R, R, generator=g)
    y = torch.randint(0, 2, (B, 1, R, R), generator=g).float()
    sd = O.seeded_state_dict(stage, 3, 1, seed=stage)
    sdp = O.seeded_state_dict(stage - 1, 3, 1, seed=stage - 1) if stage > 1 else None
    ref = O.train_step(sd, x, y, stage, sdp)
    res = []
    for flag in (False, True):
        with O.bf16_storage(True, stats_from_fp32=flag):
            e = O.train_step(sd, x, y, stage, sdp)
        gr = [rel(e["grads"][k], ref["grads"][k]) for k in ref["grads"] if not k.endswith(("conv_op.0.bias", "conv_op.3.bias"))]
        res.append((rel(e["logits"], ref["logits"]), statistics.median(gr), max(gr), e["logits"]))
    print(f"{stage} {B} | {res[0][0]:.3e} / {res[1][0]:.3e} | {res[0][1]:.3e} / {res[1][1]:.3e} | "
          f"{res[0][2]:.3e} / {res[1][2]:.3e} | {rel(res[0][3], res[1][3]):.3e}", flush=True)
