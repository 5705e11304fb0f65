"""Data-parallel check of the Herlev trainer under torchrun (>= 2 GPUs, NCCL): BASELINE.json configs[3]
shape (stage 4, 7 classes, batch 16 per rank), a few optimisation steps, then every rank must hold
bit-identical weights and the step time is reported as the max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp_herlev.py
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ug_pg_unet_b200.herlev as H


def main():
    if os.environ.get("CHECK_DP_DUMP_AFTER"):      # where is every thread if the run stalls?
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["CHECK_DP_DUMP_AFTER"]), exit=True)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = int(os.environ.get("HERLEV_RES", "256"))
    cfg = {"device": dev, "epochs_per_stage": 1, "num_classes": 7, "uncertainty_alpha": 1.0, "log_interval": 0,
           "stage4_resolution": res}
    torch.manual_seed(100 + rank)                 # ranks start different; the trainer broadcasts rank 0
    with contextlib.redirect_stdout(io.StringIO()):
        tr = H.HerlevTrainer(cfg)
        tr.setup_optimizer_scheduler(4)
    model = tr.models[4]
    model.train()
    for m in model.classifier:
        if isinstance(m, torch.nn.Dropout):
            m.eval()                              # rank-local RNG draws would be legitimate but hide bugs
    tr.models[3].eval()
    g = torch.Generator().manual_seed(7 + rank)
    B = 16
    x = torch.randn(B, 3, res, res, generator=g).to(dev)
    y = torch.randint(0, 7, (B,), generator=g).to(dev)
    for _ in range(6):                            # 3 eager warm-ups, the graph capture, 2 replays
        tr.train_step(x, y, 4)
    torch.cuda.synchronize()
    dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    a.record()
    for _ in range(steps):
        out = tr.train_step(x, y, 4)
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = 0.0
    for n, p in model.named_parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, float((p.detach() - ref).abs().max()))
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(w) == 0.0 and bool(torch.isfinite(out).all())
        print(f"check_dp_herlev {'ok' if ok else 'FAILED'}: world {world}, stage 4 at {res}x{res}, batch {B}/rank, "
              f"{float(t):.2f} ms/step ({'graph replay' if any(g.get('graph') is not None for g in tr._graphs.values()) else 'eager'}), "
              f"{world * B / float(t) * 1e3:.0f} images/s, "
              f"max weight difference across ranks {float(w):.1e}")
    # NCCL kernels captured into CUDA graphs keep the communicator busy: destroy_process_group() does not
    # return on this stack (same as bench.py / tools/check_dp.py), so flush and leave
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
