"""One stage-4 trainer step bracketed by cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 16] [--stage 4]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import ug_pg_unet_b200 as U


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--stage", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
    st = args.stage
    tr.current_stage, tr.current_model = st, tr.models[st]
    tr.setup_optimizer(st)
    tr.current_model.train()
    if st > 1:
        tr.models[st - 1].eval()
    R = tr.stage_configs[st]["resolution"]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(args.batch, 3, R, R, generator=g).to(dev)
    y = torch.randint(0, 2, (args.batch, 1, R, R), generator=g).float().to(dev)
    for _ in range(args.warmup):
        tr.train_step(x, y, st)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        tr.train_step(x, y, st)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
