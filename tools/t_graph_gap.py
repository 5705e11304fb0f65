"""Cost of one dependent kernel node inside a replayed CUDA graph (the trainer step is a chain of ~150
such nodes on its main stream): N launches of a tiny kernel captured once, replayed and timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ug_pg_unet_b200 import ops

dev = torch.device("cuda", 0)
C = 64
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
sc = torch.empty(C, device=dev); sh = torch.empty(C, device=dev)
for N in (50, 200, 800):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            ops.bn_eval_coeffs(g, b, rm, rv, 1e-5, sc, sh)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            for _ in range(N):
                ops.bn_eval_coeffs(g, b, rm, rv, 1e-5, sc, sh)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        print(f"graph of {N} dependent tiny kernels: {a.elapsed_time(e) / 10 / N * 1e3:.2f} us per node")
