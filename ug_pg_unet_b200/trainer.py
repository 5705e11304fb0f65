"""Drop-in `UncertaintyGuidedProgressiveTrainer` on the sm_100a kernels.

Mirrors the public surface of the reference's
`uncertainty_guided_trainer.py::UncertaintyGuidedProgressiveTrainer` (:25-524):
constructor, `stage_configs`, `models`, `current_model`, `current_stage`,
`base_criterion`, `optimizer`, `uncertainty_loss`, `uncertainty_alpha`, `history`,
`setup_optimizer`, `dice_coefficient`, `get_predictions`, `calculate_accuracy`,
`transfer_weights`, `uncertainty_guided_forward_pass`, `train_epoch`,
`validate_epoch`, `train_progressive`, `load_stage_weights` — so subclasses such
as `AugMoNuSegTrainer` (MoNuSegImprove/train_aug_monuseg.py:36-121) keep working.

What changes underneath: the models are the kernel-backed PGUNet1..4, the
loss + all step metrics come out of one fused kernel as a single device tensor
(one host read-back per log interval instead of six syncs per step, SURVEY.md
§3.1), and the step can be sharded across ranks (`world_size` > 1) with bucketed
NCCL gradient all-reduce overlapped with backward.  Plotting / analysis dumps are
out of scope (reporting only).
"""
from __future__ import annotations

import math
import os
import time
from pathlib import Path
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

from .engine import PackedWeights, WgradOverlap, conv_packs, enable_deferred_wgrad, nvtx_range
from .optim import FusedRMSprop
from .parallel import GradBucketReducer, broadcast_module_state
from .unet import PGUNet1, PGUNet2, PGUNet3, PGUNet4, UncertaintyGuidedLoss, _pos_weight_value, transfer_state


class UncertaintyGuidedProgressiveTrainer:
    def __init__(self, in_channels=3, num_classes=1, device="cuda", uncertainty_alpha=1.0,
                 distributed: Optional[bool] = None, bucket_mb: float = 16.0, log_interval: int = 10,
                 use_cuda_graph: bool = True):
        self.device = device
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.uncertainty_alpha = uncertainty_alpha
        self.log_interval = log_interval
        self.bucket_mb = bucket_mb
        self.use_cuda_graph = use_cuda_graph and os.environ.get("UGP_NO_CUDA_GRAPH", "0") != "1"
        self.defer_conv_grads = os.environ.get("UGP_NO_DEFERRED_WGRAD", "0") != "1"
        # two side streams inside the step: the previous stage's uncertainty forward beside the current
        # model's forward, and the weight-gradient convolutions beside the dgrad / BN-backward chain
        self.overlap_streams = os.environ.get("UGP_NO_STREAM_OVERLAP", "0") != "1"
        self._side_streams = None
        self._capture_stream = None
        self._copy_stream = None
        self._graphs = {}                 # captured step graphs, keyed by everything a capture bakes in
        self.max_step_graphs = 4          # each capture owns a memory pool; further shapes run eagerly
        self._last_step_losses = None
        self._grad_arena = None
        self._arena_packs = []
        self._stage_ring = [None, None, None]      # device staging buffers of _prefetch
        # UGP_EARLY_OPTIMIZER=1: packed-weight optimiser launched from inside backward, beside its tail, instead
        # of after it.  Off by default: same-box A/B at N = 1, 2 and 8 GPUs showed no step-time change (the tail
        # of backward is HBM-bound like the update itself), so the simpler order stays.
        self.early_optimizer = os.environ.get("UGP_EARLY_OPTIMIZER", "0") == "1"
        # weight gradients of consecutive layers on ONE side stream (1) or alternating between two (2)
        self.wgrad_streams = int(os.environ.get("UGP_WGRAD_STREAMS", "1"))

        # reference uncertainty_guided_trainer.py:42-47
        self.stage_configs = {
            1: {"resolution": 32, "epochs_per_stage": 40, "lr": 3e-4},
            2: {"resolution": 64, "epochs_per_stage": 40, "lr": 1e-4},
            3: {"resolution": 128, "epochs_per_stage": 40, "lr": 1e-4},
            4: {"resolution": 256, "epochs_per_stage": 40, "lr": 1e-4},
        }
        self.models = {
            1: PGUNet1(in_channels, num_classes).to(device),
            2: PGUNet2(in_channels, num_classes).to(device),
            3: PGUNet3(in_channels, num_classes).to(device),
            4: PGUNet4(in_channels, num_classes).to(device),
        }
        self.distributed = dist.is_initialized() and dist.get_world_size() > 1 if distributed is None else distributed
        self.world_size = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        if self.distributed:
            for m in self.models.values():
                broadcast_module_state(m)
        self._reducer = None

        self.current_stage = 1
        self.current_model = self.models[1]
        self.uncertainty_loss = UncertaintyGuidedLoss(device)

        pos_weight = torch.tensor([5.0]).to(device)                       # :64
        self.base_criterion = nn.BCEWithLogitsLoss(pos_weight=pos_weight, reduction="none")
        self.optimizer = None
        self.setup_optimizer(1)

        self.history = {
            "train_loss": [], "val_loss": [], "train_dice": [], "val_dice": [],
            "uncertainty_weights_mean": [], "uncertainty_weights_std": [], "base_loss": [],
            "stage_transitions": [],
        }

    # ------------------------------------------------------------------ optimiser / weights
    def setup_optimizer(self, stage):
        """RMSprop(lr, weight_decay=1e-4) on the current model — reference :81-88."""
        lr = self.stage_configs[stage]["lr"]
        on_cuda = next(self.current_model.parameters()).is_cuda
        # same hyper-parameters and state_dict layout as the reference's optim.RMSprop; on CUDA the
        # update of all 80 tensors is one kernel launch (FusedRMSprop subclasses optim.RMSprop)
        for g in self._graphs.values():      # graphs of the previous optimiser / model address freed tensors
            g.pop("out", None)
            g["graph"] = None
        self._graphs.clear()
        if self._reducer is not None:
            self._reducer.remove()
            self._reducer = None
        # 3x3 conv weight gradients stay in the packed layout the wgrad kernel reduces into (one flat
        # arena); FusedRMSprop consumes them there and rewrites the bf16 operand packs in the same
        # launch.  Only the model being trained defers; every other stage model is restored.
        for m in self.models.values():
            enable_deferred_wgrad(m, False)
        packs, arena = enable_deferred_wgrad(self.current_model, on_cuda and self.defer_conv_grads)
        self._grad_arena = arena
        self._arena_packs = packs            # arena order = the order backward produces the weight gradients
        if on_cuda:
            self.optimizer = FusedRMSprop(self.current_model.parameters(), lr=lr, weight_decay=1e-4,
                                          conv_packs=conv_packs(self.current_model))
            self.optimizer.deferred_grad_scale = 1.0 / self.world_size
        else:
            self.optimizer = optim.RMSprop(self.current_model.parameters(), lr=lr, weight_decay=1e-4)
        if self.distributed:
            fused = isinstance(self.optimizer, FusedRMSprop) and os.environ.get("UGP_DP_AVERAGE_IN_REDUCER", "0") != "1"
            self._reducer = GradBucketReducer(list(self.current_model.parameters()), bucket_mb=self.bucket_mb,
                                              conv_packs=packs, arena=arena, average=not fused)
            if fused:       # sums stay in the flat bucket, the optimiser kernel applies 1/world_size
                self.optimizer.dense_grad_scale = 1.0 / self.world_size

    def transfer_weights(self, prev_stage, new_stage):
        """reference :125-141, without the throw-away 49 M-parameter ProgressiveUNet."""
        print(f"Transferring weights from stage {prev_stage} to stage {new_stage}")
        prev_dict = self.models[prev_stage].state_dict()
        new_dict = self.models[new_stage].state_dict()
        self.models[new_stage].load_state_dict(transfer_state(prev_dict, new_dict, verbose_stage=new_stage))
        print(f"Weight transfer completed for stage {new_stage}")

    def load_stage_weights(self, stage, checkpoint_path):
        checkpoint = torch.load(checkpoint_path, map_location=self.device)
        sd = checkpoint["model_state_dict"] if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint \
            else checkpoint
        self.models[stage].load_state_dict(sd)
        print(f"Loaded weights for stage {stage} from {checkpoint_path}")

    # ------------------------------------------------------------------ metrics (reference :90-123)
    def dice_coefficient(self, pred, target, smooth=1):
        pred = pred.to(target.device).contiguous().float()
        target = target.contiguous().float()
        p = pred.view(pred.size(0), -1)
        t = target.view(target.size(0), -1)
        inter = (p * t).sum(dim=1)
        return ((2.0 * inter + smooth) / (p.sum(dim=1) + t.sum(dim=1) + smooth)).mean()

    def get_predictions(self, output_batch):
        return (output_batch > 0).float().squeeze(1)          # sigmoid(z) > 0.5  <=>  z > 0

    def calculate_accuracy(self, pred, target):
        pred = pred.to(target.device)
        assert pred.size() == target.size()
        return 1 - pred.ne(target).sum().item() / pred.numel()

    # ------------------------------------------------------------------ the step
    def _forward_loss(self, data, target, stage):
        """Forward + fused loss.  Returns (final_loss tensor, logits, metrics device tensor or None,
        base_loss float or None)."""
        uncertainty_map = None
        if data.dtype == torch.uint8:     # loader bytes: /255, Normalize and the stage resize run in the packer
            res = self.stage_configs[stage]["resolution"]
            run_model = lambda d: self.current_model(d, resolution=res)
        else:
            run_model = self.current_model
        side = self._streams(data)[0] if (stage > 1 and self.overlap_streams and data.is_cuda) else None
        if side is not None:
            # independent of the current model until the loss: runs beside its forward
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                uncertainty_map = self.uncertainty_loss.generate_uncertainty_map(
                    data, self.models[stage - 1], self.stage_configs[stage - 1]["resolution"],
                    self.stage_configs[stage]["resolution"])
            output_current = run_model(data)
            main.wait_stream(side)
        else:
            output_current = run_model(data)
            if stage > 1:
                uncertainty_map = self.uncertainty_loss.generate_uncertainty_map(
                    data, self.models[stage - 1], self.stage_configs[stage - 1]["resolution"],
                    self.stage_configs[stage]["resolution"])
        self.uncertainty_loss.last_metrics = None
        ul = self.uncertainty_loss
        if ul.fused_criterion(self.base_criterion, output_current, uncertainty_map):
            final_loss, _ = ul.weighted_loss_device(self.base_criterion, output_current, target,
                                                    uncertainty_map, self.uncertainty_alpha)
            base = None
        else:
            final_loss, base = ul.apply_uncertainty_weighted_loss(
                self.base_criterion, output_current, target, uncertainty_map, self.uncertainty_alpha)
        return final_loss, output_current, uncertainty_map, base

    def _metrics_from_device(self, m: torch.Tensor, n_elem: int, batch: int, has_unc: bool):
        """Turns the fused kernel's accumulators (double[5+3B]) into the reference's step metrics."""
        v = m.tolist()
        final, base = v[0] / n_elem, v[1] / n_elem
        unc_mean = v[2] / n_elem if has_unc else 0.0
        unc_std = 0.0
        if has_unc and n_elem > 1:
            var = max(v[3] - v[2] * v[2] / n_elem, 0.0) / (n_elem - 1)     # torch.std: unbiased
            unc_std = math.sqrt(var)
        acc = 1.0 - v[4] / n_elem
        dice = 0.0
        for b in range(batch):
            pt, p, t = v[5 + 3 * b: 8 + 3 * b]
            dice += (2.0 * pt + 1.0) / (p + t + 1.0)
        return final, base, dice / batch, acc, unc_mean, unc_std

    def uncertainty_guided_forward_pass(self, data, target, stage):
        """reference :143-184 — same return value `(final_loss, metrics)`."""
        final_loss, output_current, unc, base = self._forward_loss(data, target, stage)
        m = self.uncertainty_loss.last_metrics
        if m is not None:
            f, b, _, _, um, us = self._metrics_from_device(m, output_current.numel(), output_current.shape[0],
                                                            unc is not None)
            metrics = {"final_loss": f, "base_loss": b, "output": output_current,
                       "uncertainty_weight_mean": um, "uncertainty_weight_std": us}
        else:
            metrics = {"final_loss": final_loss.item(), "base_loss": base, "output": output_current,
                       "uncertainty_weight_mean": torch.mean(unc).item() if unc is not None else 0.0,
                       "uncertainty_weight_std": torch.std(unc).item() if unc is not None else 0.0}
        return final_loss, metrics

    def _resize(self, data, target, resolution):
        """reference :208-209; identity when the loader already yields the stage resolution.  uint8 HWC
        image bytes are left alone: their resize happens inside the first-layer packer."""
        if data.dtype == torch.uint8:
            if target.shape[-2:] != (resolution, resolution):
                target = F.interpolate(target, size=(resolution, resolution), mode="nearest")
            return data, target
        if data.shape[-2:] != (resolution, resolution):
            data = F.interpolate(data, size=(resolution, resolution), mode="bilinear", align_corners=True)
        if target.shape[-2:] != (resolution, resolution):
            target = F.interpolate(target, size=(resolution, resolution), mode="nearest")
        return data, target

    def _streams(self, like):
        if self._side_streams is None:
            dev = like.device
            self._side_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev),
                                  torch.cuda.Stream(device=dev))
        return self._side_streams

    def _early_optimizer_step(self):
        """Called from backward once the LAST packed weight gradient (the network's first deferred layer) and
        its input gradient are queued: starts the packed-weight optimiser on the weight-gradient side stream,
        beside the tail of the backward pass (the first BatchNorm's backward and the image layer's weight
        gradient on the main stream), instead of after it.  The update rewrites the bf16 input-gradient packs,
        so the side stream first waits for everything queued on the main stream (every dgrad) and - data
        parallel - for the all-reduce of every arena bucket."""
        side = WgradOverlap.stream
        main = torch.cuda.current_stream()
        if WgradOverlap.stream2 is not None:
            side.wait_stream(WgradOverlap.stream2)
        if self._reducer is not None:
            with torch.cuda.stream(side):            # the reducer forks from the stream the wgrads run on,
                self._reducer.flush_arena()          # and makes it wait for every arena bucket's all-reduce
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.optimizer.step_deferred()

    def _eager_step(self, data, target, stage):
        self.optimizer.zero_grad(set_to_none=True)
        with nvtx_range(f"step stage {stage}: forward + uncertainty + loss"):
            final_loss, output_current, unc, base = self._forward_loss(data, target, stage)
        with nvtx_range("step: backward"):
            if self.overlap_streams and data.is_cuda and self._grad_arena is not None:
                early = self._arena_packs[-1] if (self._arena_packs and self.early_optimizer
                                                  and isinstance(self.optimizer, FusedRMSprop)) else None
                WgradOverlap.begin(self._streams(data)[1], last_pack=early,
                                   on_last=self._early_optimizer_step if early is not None else None,
                                   stream2=self._streams(data)[2] if self.wgrad_streams == 2 else None)
                try:
                    final_loss.backward()
                finally:
                    WgradOverlap.end()
            else:
                final_loss.backward()
        with nvtx_range("step: gradient exchange + optimiser"):
            if self._reducer is not None:
                self._reducer.wait()
            self.optimizer.step()
        # a custom criterion has no fused metrics: keep this step's own (detached) losses so that the epoch
        # loop can log them without a second forward (reference :221-234 logs the step's own numbers)
        self._last_step_losses = None if self.uncertainty_loss.last_metrics is not None else (final_loss.detach(), base)
        # detached: a caller holding the logits must not keep this step's autograd graph (and its
        # AccumulateGrad nodes, created on the warm-up stream) alive into the CUDA-graph capture
        return self.uncertainty_loss.last_metrics, output_current.detach(), unc

    def train_step(self, data, target, stage):
        """One optimisation step on already-resident device tensors; no host synchronisation.
        Returns (fused-metrics device tensor, logits, uncertainty map); see `_metrics_from_device`.

        The step launches ~230 kernels; after a few eager warm-up steps the whole step (forward,
        uncertainty forward, loss, backward, optimiser) is captured once into a CUDA graph per
        (stage, shapes) and replayed, so the GPU is not left waiting for the host.  Tensors returned
        from a replay are static buffers that the next step overwrites (the metrics are cloned)."""
        crit = self.base_criterion
        fused = isinstance(crit, nn.BCEWithLogitsLoss) and crit.reduction == "none" and crit.weight is None \
            and (crit.pos_weight is None or crit.pos_weight.numel() == 1)
        dp_graph_ok = self._reducer is None or os.environ.get("UGP_DP_CUDA_GRAPH", "1") == "1"
        if not (self.use_cuda_graph and data.is_cuda and dp_graph_ok and fused):
            return self._eager_step(data, target, stage)
        # everything a capture bakes in as a constant: shapes / dtypes, the objects whose tensors it addresses,
        # and the Python floats passed to the kernels by value (alpha, pos_weight, learning rate, momentum mode)
        pw = 1.0 if crit.pos_weight is None else _pos_weight_value(crit)
        key = (stage, tuple(data.shape), data.dtype, tuple(target.shape), target.dtype, id(self.current_model),
               id(self.optimizer), self.current_model.training, float(self.uncertainty_alpha), id(crit), pw,
               tuple(float(gp["lr"]) for gp in self.optimizer.param_groups))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.max_step_graphs:
                # e.g. a stream of ragged batch shapes: run them eagerly, keep the captured graphs
                return self._eager_step(data, target, stage)
            self._graphs[key] = g = {"key": key, "warm": 0, "graph": None}
        if g["graph"] is None:
            if g["warm"] < 3:
                g["warm"] += 1
                return self._eager_step(data, target, stage)
            g["x"], g["y"] = data.clone(), target.clone()
            self.optimizer.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # The step is captured on a HIGH-priority stream while its side streams (uncertainty forward,
            # weight gradients, all-reduce) keep the default priority: a wgrad CTA needs all 512 TMEM
            # columns, so wgrad and dgrad kernels never share an SM and the block scheduler serialises them
            # - with priorities the critical-path kernel (main stream) gets the SMs first and the weight
            # gradient fills in beside the HBM-bound BatchNorm backward that follows.
            cap = None
            if os.environ.get("UGP_STREAM_PRIORITY", "1") == "1":
                if self._capture_stream is None:
                    self._capture_stream = torch.cuda.Stream(device=data.device, priority=-1)
                cap = self._capture_stream
            with torch.cuda.graph(graph, stream=cap):
                g["out"] = self._eager_step(g["x"], g["y"], stage)
            g["graph"] = graph
            g["counted"] = True            # the capture pass already advanced the optimiser's step counters
        g["x"].copy_(data, non_blocking=True)
        g["y"].copy_(target, non_blocking=True)
        g["graph"].replay()
        # the replayed optimizer step changed the fp32 weights behind autograd's version counters
        if isinstance(self.optimizer, FusedRMSprop):
            self.optimizer.note_replayed_steps(0 if g.pop("counted", False) else 1)
        else:
            PackedWeights.invalidate_all()
        m, logits, unc = g["out"]
        return m.clone(), logits, unc

    # ------------------------------------------------------------------ teardown
    def close(self, destroy_process_group: bool = False):
        """Releases the captured step graphs (and with them the NCCL kernels they reference), then -
        optionally - tears the process group down.  `dist.destroy_process_group()` does not return while a
        live CUDA graph still holds captured collectives of the communicator, so a data-parallel run must
        drop its graphs first: call this (or use the trainer as a context manager) before leaving."""
        for g in self._graphs.values():
            g.pop("out", None)
            g["graph"] = None
        self._graphs.clear()
        if self._reducer is not None:
            self._reducer.remove()
            self._reducer = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if destroy_process_group and dist.is_initialized():
            dist.destroy_process_group()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _prefetch(self, dataloader):
        """Yields device batches with the host -> device copy of batch i+1 issued on a copy stream while
        step i runs (reference :204-205 copies synchronously inside the step)."""
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        if dev.type != "cuda" or not self.overlap_streams:
            for data, target in dataloader:
                yield data.to(self.device, non_blocking=True), target.to(self.device, non_blocking=True)
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy = self._copy_stream

        # Device staging buffers are a small ring reused across batches (GPU-side events order a slot's next
        # copy after the step that consumed it): the host runs an epoch ahead of the GPU, and a fresh
        # `.to(device)` per batch then means one cudaMalloc per 16.8 MB fp32 batch - measured as stalls of
        # 100+ ms per epoch on some boxes.  Steady state allocates nothing.
        ring = self._stage_ring

        def stage_in(batch, k):
            data, target = batch
            slot = ring[k % len(ring)]
            fits = (slot is not None and slot["d"].shape == data.shape and slot["d"].dtype == data.dtype
                    and slot["t"].shape == target.shape and slot["t"].dtype == target.dtype)
            with torch.cuda.stream(copy):
                if not fits:                      # first use, or a ragged / differently typed batch
                    slot = {"d": torch.empty(data.shape, dtype=data.dtype, device=dev),
                            "t": torch.empty(target.shape, dtype=target.dtype, device=dev), "done": None}
                    ring[k % len(ring)] = slot
                elif slot["done"] is not None:
                    copy.wait_event(slot["done"])        # the step that read this slot has finished
                slot["d"].copy_(data, non_blocking=True)
                slot["t"].copy_(target, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return slot, ev

        it = iter(dataloader)
        k = 0
        try:
            nxt = stage_in(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            slot, ev = nxt
            k += 1
            try:
                nxt = stage_in(next(it), k)       # overlaps the step that consumes this slot
            except StopIteration:
                nxt = None
            main = torch.cuda.current_stream(dev)
            main.wait_event(ev)
            slot["d"].record_stream(main)
            slot["t"].record_stream(main)
            yield slot["d"], slot["t"]
            done = torch.cuda.Event()             # the consumer's work on this slot is queued on `main` by now
            done.record(torch.cuda.current_stream(dev))
            slot["done"] = done

    def _epoch(self, dataloader, stage, train: bool):
        resolution = self.stage_configs[stage]["resolution"]
        totals = [0.0] * 6
        pending = []                      # fused-metric device tensors not yet read back
        num_batches = 0

        def drain():
            last = None
            for item in pending:
                last = item if isinstance(item, tuple) and len(item) == 6 else \
                    self._metrics_from_device(*item)
                for i in range(6):
                    totals[i] += last[i]
            pending.clear()
            return last

        for batch_idx, (data, target) in enumerate(self._prefetch(dataloader)):
            data, target = self._resize(data, target, resolution)
            if train:
                m, out, unc = self.train_step(data, target, stage)
            else:
                final_loss, out, unc, base = self._forward_loss(data, target, stage)
                m = self.uncertainty_loss.last_metrics
            if m is not None:
                pending.append((m, out.numel(), out.shape[0], unc is not None))
            else:
                # custom criterion: the reference's own metric arithmetic (:221-234) on THIS step's logits and
                # losses - one forward per step, as in the reference (BN running statistics advance once)
                with torch.no_grad():
                    fl, bl = self._last_step_losses if train else (final_loss.detach(), base)
                    pred = self.get_predictions(out)
                    tsq = target.squeeze(1)
                    pending.append((float(fl), float(bl),
                                    self.dice_coefficient(pred.float(), tsq.float()).item(),
                                    self.calculate_accuracy(pred, tsq.long()),
                                    torch.mean(unc).item() if unc is not None else 0.0,
                                    torch.std(unc).item() if unc is not None else 0.0))
            num_batches += 1
            if train and self.log_interval and batch_idx % self.log_interval == 0:
                f, b, d, a, um, _ = drain()
                info = f", Unc_mean: {um:.4f}" if stage > 1 else ""
                print(f"Stage {stage}, Batch {batch_idx}, Loss: {f:.4f}, Base_Loss: {b:.4f}, "
                      f"Dice: {d:.4f}, Acc: {a:.4f}{info}")
        drain()
        n = max(num_batches, 1)
        kind = "training" if train else "validation"
        print(f"Stage {stage} {kind} epoch completed. Batches processed: {num_batches}")
        return tuple(t / n for t in totals)

    def train_epoch(self, dataloader, stage):
        """reference :186-256 — returns (loss, base_loss, dice, accuracy, unc_mean, unc_std) averages."""
        self.current_model.train()
        if stage > 1:
            self.models[stage - 1].eval()
        return self._epoch(dataloader, stage, train=True)

    def validate_epoch(self, dataloader, stage):
        """reference :258-314."""
        self.current_model.eval()
        if stage > 1:
            self.models[stage - 1].eval()
        with torch.no_grad():
            return self._epoch(dataloader, stage, train=False)

    def train_progressive(self, train_loader, val_loader, max_stages=4, save_dir="./uncertainty_guided_weights"):
        """reference :316-398 (checkpoint dict layout preserved; plots omitted)."""
        save_path = Path(save_dir)
        save_path.mkdir(exist_ok=True)
        print("Starting Uncertainty-Guided Progressive Growing U-Net Training")
        print("=" * 60)
        for stage in range(1, max_stages + 1):
            res = self.stage_configs[stage]["resolution"]
            print(f"\nStarting Stage {stage}")
            print(f"Resolution: {res}x{res}")
            if stage > 1:
                print(f"Using uncertainty-guided loss weighting (alpha={self.uncertainty_alpha})")
                self.transfer_weights(stage - 1, stage)
            print("-" * 40)
            self.current_stage = stage
            self.current_model = self.models[stage]
            self.setup_optimizer(stage)
            self.history["stage_transitions"].append(len(self.history["train_loss"]))
            epochs = self.stage_configs[stage]["epochs_per_stage"]
            best_val_dice = 0
            for epoch in range(epochs):
                start_time = time.time()
                tr = self.train_epoch(train_loader, stage)
                va = self.validate_epoch(val_loader, stage)
                self.history["train_loss"].append(tr[0])
                self.history["val_loss"].append(va[0])
                self.history["train_dice"].append(tr[2])
                self.history["val_dice"].append(va[2])
                self.history["uncertainty_weights_mean"].append(va[4])
                self.history["uncertainty_weights_std"].append(va[5])
                self.history["base_loss"].append(va[1])
                print(f"Stage {stage}, Epoch {epoch + 1}/{epochs} ({time.time() - start_time:.2f}s)")
                print(f"Train - Loss: {tr[0]:.4f}, Base: {tr[1]:.4f}, Dice: {tr[2]:.4f}, Acc: {tr[3]:.4f}")
                print(f"Val   - Loss: {va[0]:.4f}, Base: {va[1]:.4f}, Dice: {va[2]:.4f}, Acc: {va[3]:.4f}")
                if stage > 1:
                    print(f"Uncertainty - Mean: {va[4]:.4f}, Std: {va[5]:.4f}")
                if va[2] > best_val_dice and self.rank == 0:
                    best_val_dice = va[2]
                    torch.save({
                        "stage": stage, "epoch": epoch,
                        "model_state_dict": self.current_model.state_dict(),
                        "optimizer_state_dict": self.optimizer.state_dict(),
                        "val_dice": va[2], "train_dice": tr[2],
                        "uncertainty_alpha": self.uncertainty_alpha, "history": self.history,
                    }, save_path / f"ug_pgunet_stage{stage}_best.pth")
                print("-" * 60)
        print("Uncertainty-guided progressive training completed!")
