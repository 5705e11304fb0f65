"""Optimisers of the two trainers as single kernel launches.

FusedRMSprop: RMSprop with the reference's configuration (uncertainty_guided_trainer.py:84-88) as ONE kernel
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
    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False,
                 conv_packs=None, **kw):
        """`conv_packs`: the `engine.PackedWeights` of the model's convolutions.  Those with deferred
        weight gradients (left in the packed layout by backward, see `engine.enable_deferred_wgrad`)
        get their update, the refresh of the bf16 operand packs and the zeroing of the gradient
        accumulator in one kernel launch; all other parameters take the multi-tensor kernel from
        their `.grad`, and the operand packs of non-deferred convolutions are marked stale."""
        kw.pop("capturable", None)
        kw.pop("foreach", None)
        super().__init__(params, lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum,
                         centered=centered, **kw)
        self._replayed_steps = 0
        self._targeted = conv_packs is not None
        self._packs = {id(pk.conv.weight): pk for pk in (conv_packs or [])}
        self.deferred_grad_scale = 1.0      # 1/world_size when the packed gradients were sum-all-reduced
        self.dense_grad_scale = 1.0         # the same for `.grad` tensors a reducer left as sums over ranks

    def note_replayed_steps(self, n: int = 1) -> None:
        """A CUDA-graph replay ran `step()`'s kernels without running this Python code."""
        self._replayed_steps += n
        self._invalidate_packs(())

    def _invalidate_packs(self, rewritten) -> None:
        """The kernels write parameters without advancing autograd's version counters: bf16 operand
        caches of the updated convolutions are stale, except those the packed kernel has rewritten."""
        from .engine import PackedWeights

        if not self._targeted:
            PackedWeights.invalidate_all()
            return
        fresh = {id(pk) for pk in rewritten}
        for pk in self._packs.values():
            pk._eval_key = None          # BN affine parameters / running statistics moved as well
            if id(pk) in fresh:
                pk.grad_pending = False
                pk.mark_fresh()
            elif not pk.defer:
                pk.invalidate()

    def state_dict(self):
        if self._replayed_steps:
            for st in self.state.values():
                if "step" in st:
                    st["step"] = st["step"] + self._replayed_steps
            self._replayed_steps = 0
        return super().state_dict()

    @torch.no_grad()
    def step_deferred(self) -> int:
        """The packed-gradient part of `step()` alone: RMSprop + bf16 operand repack + gradient zeroing of every
        convolution whose packed weight gradient is pending, in one launch on the CURRENT stream.  The trainer
        calls it on its side stream as soon as the last weight gradient and the last input gradient of the
        backward pass are queued, so the 375 MB update runs beside the tail of backward instead of after it;
        the `step()` that follows skips these tensors.  Returns the number of tensors updated."""
        n = 0
        for group in self.param_groups:
            if group["momentum"] != 0 or group["centered"] or group.get("maximize", False):
                return 0
            dw, dv, dg, dwf, dwd, dpk = [], [], [], [], [], []
            for p in group["params"]:
                pk = self._packs.get(id(p))
                if pk is None or not (pk.defer and pk.grad_pending and p.grad is None):
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["square_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                wf, wd = pk.get()
                dw.append(p.data); dv.append(st["square_avg"]); dg.append(pk._dwp)
                dwf.append(wf); dwd.append(wd); dpk.append(pk)
            if dw:
                ops.rmsprop_conv_packed(dw, dv, dg, dwf, dwd, float(group["lr"]), float(group["alpha"]),
                                        float(group["eps"]), float(group["weight_decay"]),
                                        float(self.deferred_grad_scale))
                for pk in dpk:
                    pk.grad_pending = False
                    pk.mark_fresh()
                n += len(dw)
        return n

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        rewritten = []
        for group in self.param_groups:
            if group["momentum"] != 0 or group["centered"] or group.get("maximize", False):
                if any(pk.defer for pk in self._packs.values()):
                    raise RuntimeError("FusedRMSprop: deferred conv gradients need momentum=0, centered=False")
                return super().step()
            ps, gs, vs = [], [], []
            dw, dv, dg, dwf, dwd, dpk = [], [], [], [], [], []
            for p in group["params"]:
                pk = self._packs.get(id(p))
                if pk is not None and pk.defer and pk.grad_pending and p.grad is None:
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = torch.zeros((), dtype=torch.float32)
                        st["square_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["step"] += 1
                    wf, wd = pk.get()
                    dw.append(p.data)
                    dv.append(st["square_avg"])
                    dg.append(pk._dwp)
                    dwf.append(wf)
                    dwd.append(wd)
                    dpk.append(pk)
                    continue
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
                                  float(group["weight_decay"]), float(self.dense_grad_scale))
            if dw:
                ops.rmsprop_conv_packed(dw, dv, dg, dwf, dwd, float(group["lr"]), float(group["alpha"]),
                                        float(group["eps"]), float(group["weight_decay"]),
                                        float(self.deferred_grad_scale))
                rewritten.extend(dpk)
        self._invalidate_packs(rewritten)
        return loss


class FusedAdam(torch.optim.Adam):
    """`torch.optim.Adam` (L2 weight decay, the Herlev trainer's optimiser, Herlev/train_herlev.py:170-176)
    with `step()` as one multi-tensor kernel launch.  Same constructor arguments, `param_groups` and
    `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`); amsgrad / maximize / non-CUDA parameters are
    delegated to the stock implementation.  `dense_grad_scale` multiplies the gradients inside the kernel
    (1/world_size when a reducer left them as sums over ranks).

    The step count that the bias corrections need also lives in a device int64 (per parameter group,
    advanced by a one-thread kernel), so a captured CUDA graph containing `step()` stays correct when
    replayed; `note_replayed_steps` keeps the per-parameter `step` entries of `state_dict()` in sync."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
        kw.pop("capturable", None)
        kw.pop("foreach", None)
        kw.pop("fused", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)
        self.dense_grad_scale = 1.0    # (torch.optim.Adam itself uses the attribute name `grad_scale` for AMP)
        self._dev_steps = {}           # id(group) -> int64 device scalar
        self._replayed_steps = 0

    def note_replayed_steps(self, n: int = 1) -> None:
        """A CUDA-graph replay ran `step()`'s kernels (and advanced the device step count) without
        running this Python code."""
        self._replayed_steps += n
        self._bump_versions([p for g in self.param_groups for p in g["params"] if p in self.state])

    @staticmethod
    def _bump_versions(params) -> None:
        """The kernel wrote these parameters without autograd noticing: advance their version counters so
        that exactly the caches keyed on them (bf16 operand packs, folded BN coefficients) refresh - not
        those of other models, e.g. the frozen previous-stage network."""
        if params:
            torch.autograd.graph.increment_version(params)

    def _sync_host_steps(self) -> None:
        if self._replayed_steps:
            for st in self.state.values():
                if "step" in st:
                    st["step"] = st["step"] + self._replayed_steps
            self._replayed_steps = 0

    def state_dict(self):
        self._sync_host_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._replayed_steps = 0
        self._dev_steps = {}           # rebuilt from the loaded per-parameter step counts

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        updated = []
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._sync_host_steps()
        # eligibility first, nothing mutated yet: anything the kernel cannot take goes to the stock update
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            prospective = {int(self.state[p]["step"]) + 1 if len(self.state[p]) else 1 for p in live}
            if (group["amsgrad"] or group.get("maximize", False) or len(prospective) > 1
                    or any(not p.is_cuda or p.grad.is_sparse or p.dtype != torch.float32 for p in live)):
                self._dev_steps = {}       # rebuilt from the host-side step counts at the next fused step
                return super().step()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            t = 0
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                t = int(st["step"])
                ps.append(p.data)
                updated.append(p)
                gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if not ps:
                continue
            dev = self._dev_steps.get(id(group))
            if dev is None or dev.device != ps[0].device:
                if capturing:
                    raise RuntimeError("FusedAdam: run at least one eager step() before capturing a CUDA graph")
                dev = torch.full((), t - 1, dtype=torch.int64, device=ps[0].device)
                self._dev_steps[id(group)] = dev
            ops.counter_add(dev, 1)
            b1, b2 = group["betas"]
            ops.adam_multi(ps, gs, ms, vs, t, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                           float(group["weight_decay"]), float(self.dense_grad_scale), step_dev=dev)
        self._bump_versions(updated)
        return loss
