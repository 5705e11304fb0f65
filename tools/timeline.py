"""Timeline of ONE replay of the real (multi-stream) trainer-step graph: every instrumented launch is
bracketed by external event-record nodes on the stream it was captured on; after a replay each launch gets
(start, end) relative to the first event.  Shows where the step's critical path and its gaps are
(profiles/r02_timeline_*.txt).

    python tools/timeline.py [--batch 16] [--out gpurun_out/timeline.txt]
"""
import argparse
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ug_pg_unet_b200 as U
from ug_pg_unet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    tr = U.UncertaintyGuidedProgressiveTrainer(3, 1, device=dev, log_interval=0)
    for s in (2, 3, 4):
        tr.transfer_weights(s - 1, s)
tr.current_stage, tr.current_model = 4, tr.models[4]
tr.setup_optimizer(4)
tr.current_model.train()
tr.models[3].eval()
g = torch.Generator().manual_seed(1)
x = torch.randn(args.batch, 3, 256, 256, generator=g).to(dev)
y = torch.randint(0, 2, (args.batch, 1, 256, 256), generator=g).float().to(dev)
for _ in range(3):
    tr._eager_step(x, y, 4)
torch.cuda.synchronize()

# record which stream each launch was captured on
streams = []
orig_enter = ops._timed.__enter__


def enter(self):
    streams.append(torch.cuda.current_stream().cuda_stream)
    return orig_enter(self)


ops._timed.__enter__ = enter
tr.optimizer.zero_grad(set_to_none=True)
ops.start_profile(external=True)
cap = torch.cuda.Stream(device=dev, priority=-1)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=cap):
    base = torch.cuda.Event(enable_timing=True, external=True)
    base.record()
    tr._eager_step(x, y, 4)
    end = torch.cuda.Event(enable_timing=True, external=True)
    end.record()
evs = list(ops.profile_events())
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    graph.replay()
e1.record()
torch.cuda.synchronize()
rows = []
ids = {}
for (name, fl, by, a, b), st in zip(evs, streams):
    sid = ids.setdefault(st, len(ids))
    rows.append((base.elapsed_time(a) * 1e3, base.elapsed_time(b) * 1e3, sid, name, fl))
total = base.elapsed_time(end) * 1e3
rows.sort()
out = [f"# batch {args.batch}: one replay of the step graph, {total:.0f} us from first to last node "
       f"({e0.elapsed_time(e1) * 100:.0f} us/step over 10 back-to-back replays with the event nodes); "
       f"stream 0 = main (high priority), others = side streams",
       "# start_us  end_us  dur_us  stream  kernel  [TFLOP/s]"]
busy = {}
last_end = {}
gaps = {}
for s, e, sid, name, fl in rows:
    out.append(f"{s:9.1f} {e:9.1f} {e - s:7.1f}   s{sid}  {name:22s} " + (f"{fl / (e - s) / 1e6:7.0f}" if fl else ""))
    busy[sid] = busy.get(sid, 0.0) + (e - s)
    if sid in last_end:
        gaps.setdefault(sid, []).append((s - last_end[sid], name, s))
    last_end[sid] = e
out.append("# per stream: busy us; the ten largest gaps before a launch (us, launch, at)")
for sid in sorted(busy):
    top = sorted(gaps.get(sid, []), reverse=True)[:10]
    out.append(f"#   s{sid}: busy {busy[sid]:.0f} us; " + ", ".join(f"{g:.0f} before {n}@{at:.0f}" for g, n, at in top))
txt = "\n".join(out)
print(txt)
if args.out:
    with open(args.out, "w") as f:
        f.write(txt + "\n")
