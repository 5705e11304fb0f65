"""RMSprop with the reference's configuration (uncertainty_guided_trainer.py:84-88) as ONE kernel
launch over all parameter tensors.

Subclasses `torch.optim.RMSprop`, so constructor arguments, `param_groups`, `state_dict()` layout
(`step`, `square_avg` per parameter) and checkpoints are those of the stock optimiser; only `step()`
is replaced.  CUDA parameters only; momentum and centered variants are delegated to the stock
implementation.
"""
from __future__ import annotations

import torch

from . import ops


class FusedRMSprop(torch.optim.RMSprop):
    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False, **kw):
        kw.pop("capturable", None)
        kw.pop("foreach", None)
        super().__init__(params, lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum,
                         centered=centered, **kw)
        self._replayed_steps = 0

    def note_replayed_steps(self, n: int = 1) -> None:
        """A CUDA-graph replay ran `step()`'s kernel without running this Python code."""
        self._replayed_steps += n

    def state_dict(self):
        if self._replayed_steps:
            for st in self.state.values():
                if "step" in st:
                    st["step"] = st["step"] + self._replayed_steps
            self._replayed_steps = 0
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group["momentum"] != 0 or group["centered"] or group.get("maximize", False):
                return super().step()
            ps, gs, vs = [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.grad.is_sparse:
                    return super().step()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                ps.append(p.data)
                gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                vs.append(st["square_avg"])
            if ps:
                ops.rmsprop_multi(ps, gs, vs, float(group["lr"]), float(group["alpha"]), float(group["eps"]),
                                  float(group["weight_decay"]))
        return loss
