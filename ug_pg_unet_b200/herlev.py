"""Herlev classification variant on the sm_100a kernels — interface mirror of the reference's
`Herlev/train_herlev.py`: `HerlevClassificationModel` (:29-120) and `HerlevTrainer` (:123-463).

The encoder of the stage's PGUNet (`inc`, downs up to `down3`) runs on the tcgen05 implicit-GEMM
convolutions exactly as in the segmentation path (NHWC bf16, fused BN/ReLU/pool); the
`AdaptiveAvgPool2d(1)` in front of the MLP head is a hand-written reduction kernel; the softmax
entropy of the previous stage (`u = H(softmax)/ln K`, :253-257) and the uncertainty-weighted
cross entropy with its gradient and the accuracy count (:266-281, :318-320) are one kernel each.
The 3-layer MLP head (512 -> 512 -> 256 -> K on a [B, 512] vector, < 0.002 % of the step's FLOPs)
stays stock `nn.Linear` / `nn.Dropout`, so `state_dict` keys (`unet.*`, `classifier.{3,6,9}.*`)
and checkpoints are the reference's; in eval mode (the previous stage's uncertainty pass, validation,
evaluators) GAP + the three Linear layers run as ONE kernel (`ugp_gap_mlp_head`).
"""
from __future__ import annotations

import json
import math
import os
from datetime import datetime
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import engine, ops
from .optim import FusedAdam
from .parallel import GradBucketReducer, broadcast_module_state
from .unet import STAGE_MODELS, _PGUNetBase
from .unet_parts import _require_cuda

_RES = {1: 32, 2: 64, 3: 128, 4: 256}


def _encode_resized(unet: _PGUNetBase, x: torch.Tensor, resolution: int, depth: int) -> torch.Tensor:
    """Inference-only encoder forward of `F.interpolate(x, resolution, bilinear, align_corners=True)`
    with the resize folded into the first-layer packer; returns the deepest NHWC bf16 feature map."""
    col = unet.first_rows(x, resolution)     # loader bytes [B,H,W,3]: /255 + Normalize folded into the row builder
    dc = unet.inc.conv
    downs = [getattr(unet, n).mpconv[1] for n in unet._downs][:depth]
    a, _, _ = engine.cbr_forward([col], 1, dc.conv_op[0], dc.conv_op[1], dc._packs[0], False, False)
    a, p, _ = engine.cbr_forward([a], 9, dc.conv_op[3], dc.conv_op[4], dc._packs[1], bool(downs), False)
    for i, d in enumerate(downs):
        a, _, _ = engine.cbr_forward([p], 9, d.conv_op[0], d.conv_op[1], d._packs[0], False, False)
        a, p, _ = engine.cbr_forward([a], 9, d.conv_op[3], d.conv_op[4], d._packs[1], i + 1 < len(downs), False)
    return a


class HerlevClassificationModel(nn.Module):
    """U-Net encoder as feature extractor + MLP head — reference Herlev/train_herlev.py:29-120."""

    def __init__(self, stage: int, num_classes: int, pretrained_unet_path: Optional[str] = None):
        super().__init__()
        if stage not in STAGE_MODELS:
            raise ValueError("Stage must be 1, 2, 3, or 4")
        self.stage = stage
        self.num_classes = num_classes
        self.unet = STAGE_MODELS[stage](in_channels=3, num_classes=1)
        if pretrained_unet_path and os.path.exists(pretrained_unet_path):
            print(f"Loading pretrained U-Net weights from: {pretrained_unet_path}")
            state_dict = torch.load(pretrained_unet_path, map_location="cpu")
            if "model_state_dict" in state_dict:
                state_dict = state_dict["model_state_dict"]
            self.unet.load_state_dict(state_dict)
        # the reference probes this with a forward pass (:58-63); the encoder always ends at 512 channels
        feature_dim = 512
        self.classifier = nn.Sequential(
            nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Dropout(0.5),
            nn.Linear(feature_dim, 512), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(512, 256), nn.ReLU(), nn.Dropout(0.2),
            nn.Linear(256, num_classes),
        )
        if pretrained_unet_path:
            self._freeze_encoder()

    @property
    def _depth(self) -> int:
        return self.stage - 1            # inc only | +down3 | +down2,down3 | +down1,down2,down3  (:83-102)

    def encoder_parameters(self):
        """Parameters that `_extract_features` touches (what receives gradients)."""
        mods = [self.unet.inc] + [getattr(self.unet, n) for n in self.unet._downs][: self._depth]
        for m in mods:
            yield from m.parameters()

    def _features_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "HerlevClassificationModel")
        return self.unet.encode(x, self._depth)[-1]

    def _extract_features(self, x):
        """Reference :83-102 — fp32 NCHW feature map of the last encoder block."""
        a = self._features_nhwc(x)
        return engine.nhwc_as_nchw(a).float().contiguous()

    def _freeze_encoder(self):
        for p in self.unet.parameters():
            p.requires_grad = False

    def unfreeze_encoder(self):
        for p in self.unet.parameters():
            p.requires_grad = True

    def _head(self, pooled: torch.Tensor) -> torch.Tensor:
        """Stock-module head on an already pooled vector (kept for classifiers that were edited)."""
        for m in list(self.classifier)[2:]:       # [0] pool and [1] flatten ran in the kernel
            pooled = m(pooled)
        return pooled

    def _fused_head_layout(self):
        """(fc1, fc2, fc3, (p0, p1, p2)) when `classifier` still is the reference's Sequential (:66-77):
        the fused kernel then replaces pool + flatten + 3 dropouts + 3 linears + 2 ReLUs; else None."""
        c = list(self.classifier)
        kinds = (nn.AdaptiveAvgPool2d, nn.Flatten, nn.Dropout, nn.Linear, nn.ReLU, nn.Dropout, nn.Linear, nn.ReLU,
                 nn.Dropout, nn.Linear)
        if len(c) != len(kinds) or not all(isinstance(m, k) for m, k in zip(c, kinds)):
            return None
        fcs = (c[3], c[6], c[9])
        if any(fc.bias is None or not fc.weight.is_cuda or fc.weight.dtype != torch.float32 for fc in fcs):
            return None
        if c[0].output_size not in (1, (1, 1)) or fcs[0].in_features % 64 or fcs[0].out_features % 4 or fcs[1].out_features % 4:
            return None
        ps = tuple(float(d.p) if d.training else 0.0 for d in (c[2], c[5], c[8]))
        if any(pp >= 1.0 for pp in ps):
            return None
        return fcs[0], fcs[1], fcs[2], ps

    def _classify(self, feats_nhwc: torch.Tensor) -> torch.Tensor:
        lay = self._fused_head_layout() if os.environ.get("UGP_NO_FUSED_HEAD", "0") != "1" else None
        if lay is None or feats_nhwc.shape[0] > 64 or feats_nhwc.shape[3] != lay[0].in_features:
            return self._head(engine.GlobalAvgPoolFn.apply(feats_nhwc))
        fc1, fc2, fc3, ps = lay
        u = None
        if any(pp > 0.0 for pp in ps):      # train mode: one launch of uniforms for all three Dropout layers
            u = torch.rand((feats_nhwc.shape[0], fc1.in_features + fc1.out_features + fc2.out_features),
                           dtype=torch.float32, device=feats_nhwc.device)
        return engine.GapMlpHeadFn.apply(feats_nhwc, fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight,
                                         fc3.bias, u, *ps)

    def forward(self, x):
        return self._classify(self._features_nhwc(x))

    @torch.no_grad()
    def forward_resized(self, x, resolution: int):
        """`self(F.interpolate(x, resolution, bilinear, align_corners=True))` in eval mode, resize
        folded into the first-layer packer (reference :243-248)."""
        if x.dtype != torch.uint8 and x.shape[-1] == resolution and x.shape[-2] == resolution:
            return self(x)
        a = _encode_resized(self.unet, x, resolution, self._depth)
        return self._classify(a)


class HerlevTrainer:
    """Progressive trainer for Herlev cervical cell classification — reference :123-463.

    Same constructor (`config` dict), attributes (`models`, `optimizers`, `schedulers`, `criterion`,
    `uncertainty_alpha`, `history`, `stage_configs`) and methods.  With `torch.distributed`
    initialised the step is batch-sharded: gradients of the encoder + classifier are averaged with
    the bucketed NCCL all-reduce of `parallel.GradBucketReducer`."""

    def __init__(self, config):
        self.config = config
        self.device = config["device"]
        self.stage_configs = {
            1: {"resolution": 32, "epochs": config["epochs_per_stage"], "lr": 3e-4},
            2: {"resolution": 64, "epochs": config["epochs_per_stage"], "lr": 1e-4},
            3: {"resolution": 128, "epochs": config["epochs_per_stage"], "lr": 1e-4},
            # the reference trains stage 4 at 224 (:138): 224 -> 112 -> 56 -> 28 pixel maps run on the
            # general-size tiles of the conv kernels (boxes of <= 128 pixels whose sides divide the map);
            # BASELINE.json configs[3] quotes 256 (power-of-two tiles, the fast halo / pair kernels)
            4: {"resolution": config.get("stage4_resolution", 224), "epochs": config["epochs_per_stage"], "lr": 1e-4},
        }
        self.current_stage = 1
        self.models, self.optimizers, self.schedulers = {}, {}, {}
        for stage in range(1, 5):
            self.models[stage] = HerlevClassificationModel(
                stage=stage, num_classes=config["num_classes"],
                pretrained_unet_path=config.get("pretrained_unet_paths", {}).get(stage)).to(self.device)
        self.distributed = dist.is_initialized() and dist.get_world_size() > 1
        self.world_size = dist.get_world_size() if self.distributed else 1
        if self.distributed:
            for m in self.models.values():
                broadcast_module_state(m)
        self._reducers = {}
        self._reducer_sig = {}
        self._graphs = {}                 # captured step graphs keyed by everything a capture bakes in
        self.max_step_graphs = 4
        self._capture_stream = None
        self.use_cuda_graph = config.get("use_cuda_graph", True) and os.environ.get("UGP_NO_CUDA_GRAPH", "0") != "1"
        self.setup_loss_function()
        from .unet import UncertaintyGuidedLoss
        self.uncertainty_loss = UncertaintyGuidedLoss(self.device)
        self.uncertainty_alpha = config.get("uncertainty_alpha", 1.0)
        self.history = {"train_loss": [], "val_loss": [], "train_acc": [], "val_acc": [],
                        "uncertainty_weights_mean": [], "uncertainty_weights_std": [], "base_loss": [],
                        "stage_transitions": []}

    def setup_loss_function(self):
        if self.config.get("class_weights") is not None:
            cw = torch.FloatTensor(self.config["class_weights"]).to(self.device)
            self.criterion = nn.CrossEntropyLoss(weight=cw)
        else:
            self.criterion = nn.CrossEntropyLoss()

    def setup_optimizer_scheduler(self, stage):
        model = self.models[stage]
        lr = self.stage_configs[stage]["lr"]
        # same hyper-parameters / state_dict layout as the reference's optim.Adam; on CUDA the update of
        # all tensors is one kernel launch (FusedAdam subclasses optim.Adam)
        self._drop_graphs()
        adam = FusedAdam if torch.device(self.device).type == "cuda" else torch.optim.Adam
        self.optimizers[stage] = adam(model.parameters(), lr=lr, weight_decay=self.config.get("weight_decay", 1e-4))
        self.schedulers[stage] = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.optimizers[stage], mode="min", factor=0.5, patience=5)
        self._build_reducer(stage)

    def _grad_signature(self, stage):
        model = self.models[stage]
        return tuple(p.requires_grad for p in list(model.encoder_parameters()) + list(model.classifier.parameters()))

    def _build_reducer(self, stage):
        """(Re)builds the data-parallel gradient reducer of `stage` over the parameters that receive
        gradients NOW; called again from the step when `requires_grad` flags changed since (e.g.
        `unfreeze_encoder()` after `setup_optimizer_scheduler`), so newly trainable tensors are reduced too."""
        if not self.distributed:
            return
        model = self.models[stage]
        old = self._reducers.pop(stage, None)
        if old is not None:
            old.remove()
        self._drop_graphs()
        used = [p for p in list(model.encoder_parameters()) + list(model.classifier.parameters())
                if p.requires_grad]
        fused = isinstance(self.optimizers[stage], FusedAdam)
        self._reducers[stage] = GradBucketReducer(used, bucket_mb=self.config.get("bucket_mb", 16.0),
                                                  average=not fused)
        self._reducer_sig[stage] = self._grad_signature(stage)
        if fused:       # .grad stays a view of the summed bucket; the optimiser kernel applies 1/world_size
            self.optimizers[stage].dense_grad_scale = 1.0 / self.world_size

    def _drop_graphs(self):
        for g in self._graphs.values():
            g.pop("out", None)
            g["graph"] = None
        self._graphs.clear()

    def close(self, destroy_process_group: bool = False):
        """Releases the captured step graphs and the reducers, then (optionally) the process group:
        `dist.destroy_process_group()` does not return while a live CUDA graph holds captured NCCL kernels."""
        self._drop_graphs()
        for r in self._reducers.values():
            r.remove()
        self._reducers.clear()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if destroy_process_group and dist.is_initialized():
            dist.destroy_process_group()

    def transfer_weights(self, prev_stage, current_stage):
        """reference :200-217 — copies classifier tensors whose shapes match."""
        print(f"Transferring weights from stage {prev_stage} to {current_stage}")
        prev_c, cur_c = self.models[prev_stage].classifier, self.models[current_stage].classifier
        with torch.no_grad():
            for (pn, pp), (cn, cp) in zip(prev_c.named_parameters(), cur_c.named_parameters()):
                if pp.shape == cp.shape:
                    cp.copy_(pp)
                    print(f"  Transferred {cn}")

    # ------------------------------------------------------------------ the step
    def _uncertainty(self, data, stage):
        """Per-sample uncertainty from the previous stage (reference :232-263), detached [B]."""
        prev = self.models[stage - 1]
        prev.eval()
        prev_res = self.stage_configs[stage - 1]["resolution"]
        with torch.no_grad():
            out_prev = prev.forward_resized(data, prev_res).float().contiguous()
            K = self.config["num_classes"]
            if K > 2:
                return ops.softmax_entropy(out_prev)
            # binary head: the reference's sigmoid margin on the [B, 2] logits (:258-261), stock ops
            return 1.0 - 2.0 * torch.abs(torch.sigmoid(out_prev) - 0.5)

    def _forward_loss(self, data, target, stage):
        """-> (final_loss, base_loss tensor, logits, uncertainty weights or None, fused out or None)."""
        model = self.models[stage]
        logits = model(data)
        u = self._uncertainty(data, stage) if stage > 1 else None
        fused = self.criterion.weight is None and logits.is_cuda and (u is None or u.dim() == 1) \
            and getattr(self.criterion, "label_smoothing", 0.0) == 0.0
        if fused:
            final, out = engine.CEWeightedLossFn.apply(logits.float(), target, u, float(self.uncertainty_alpha))
            base = (out[1] / logits.shape[0]).float()
            w = None if u is None else 1.0 + self.uncertainty_alpha * u
            return final, base, logits, w, out
        # class-weighted / binary configurations: the reference's own arithmetic (:265-281)
        base = self.criterion(logits, target)
        if u is not None:
            w = 1.0 + self.uncertainty_alpha * u.squeeze()
            if w.dim() == 0:
                w = w.unsqueeze(0)
            final = torch.mean(F.cross_entropy(logits, target, reduction="none") * w.detach())
        else:
            final, w = base, None
        return final, base, logits, w, None

    def uncertainty_guided_forward_pass(self, data, target, stage):
        """reference :219-296 — returns (final_loss, metrics dict)."""
        final, base, logits, w, _ = self._forward_loss(data, target, stage)
        metrics = {
            "final_loss": final.item(), "base_loss": base.item(), "output": logits,
            "uncertainty_weight_mean": torch.mean(w).item() if w is not None else 0.0,
            "uncertainty_weight_std": torch.std(w).item() if w is not None and w.numel() > 1 else 0.0,
        }
        return final, metrics

    def train_step(self, data, target, stage):
        """One optimisation step on device tensors; returns a device tensor (final, base, #correct, w mean,
        w std) — no host synchronisation.  The step is ~250 short kernels and an eager pass is bound by the
        host's launch rate, so after three eager warm-ups it is captured once into a CUDA graph per
        (stage, shapes, learning rate) and replayed (forward, previous-stage uncertainty, fused loss,
        backward, gradient all-reduce, FusedAdam with its device-side step count)."""
        opt = self.optimizers[stage]
        if self.distributed and self._reducer_sig.get(stage) != self._grad_signature(stage):
            self._build_reducer(stage)          # requires_grad changed since the reducer was built
        # (data parallel: the all-reduces are captured too; verified on 2 GPUs - rank-identical weights.
        # Call close() before dist.destroy_process_group(): a live graph holding captured NCCL kernels keeps
        # the communicator busy.  UGP_HERLEV_DP_GRAPH=0 keeps DP steps eager.)
        use_graph = (self.use_cuda_graph and data.is_cuda and isinstance(opt, FusedAdam)
                     and self.criterion.weight is None and self.config["num_classes"] > 2
                     and (not self.distributed or os.environ.get("UGP_HERLEV_DP_GRAPH", "1") == "1"))
        if not use_graph:
            return self._eager_step(data, target, stage)
        key = (stage, tuple(data.shape), data.dtype, tuple(target.shape), id(self.models[stage]), id(opt),
               tuple(float(gp["lr"]) for gp in opt.param_groups), self.models[stage].training,
               tuple(bool(m.training) for m in self.models[stage].classifier), float(self.uncertainty_alpha),
               id(self.criterion), self._grad_signature(stage))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.max_step_graphs:
                return self._eager_step(data, target, stage)       # ragged shapes: keep the captured graphs
            self._graphs[key] = g = {"key": key, "warm": 0, "graph": None}
        if g["graph"] is None:
            if g["warm"] < 3:
                g["warm"] += 1
                return self._eager_step(data, target, stage)
            g["x"], g["y"] = data.clone(), target.clone()
            opt.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(device=data.device, priority=-1)
            with torch.cuda.graph(graph, stream=self._capture_stream):
                g["out"] = self._eager_step(g["x"], g["y"], stage)
            g["graph"] = graph
            g["counted"] = True            # the capture pass already advanced the host-side step counts
        g["x"].copy_(data, non_blocking=True)
        g["y"].copy_(target, non_blocking=True)
        g["graph"].replay()
        opt.note_replayed_steps(0 if g.pop("counted", False) else 1)
        return g["out"].clone()

    def _eager_step(self, data, target, stage):
        opt = self.optimizers[stage]
        opt.zero_grad(set_to_none=True)
        final, base, logits, w, out = self._forward_loss(data, target, stage)
        final.backward()
        red = self._reducers.get(stage)
        if red is not None:
            red.wait()
        opt.step()
        with torch.no_grad():
            correct = out[2] if out is not None else (logits.argmax(dim=1) == target).sum()
            zero = torch.zeros((), device=logits.device)
            wm = torch.mean(w) if w is not None else zero
            ws = torch.std(w) if w is not None and w.numel() > 1 else zero
        return torch.stack([final.detach().double(), base.detach().double(), correct.double(), wm.double(),
                            ws.double()])

    def _epoch(self, dataloader, stage, train: bool):
        acc = None
        total = 0
        log_every = self.config.get("log_interval", 10)
        for batch_idx, (data, target) in enumerate(dataloader):
            data = data.to(self.device, non_blocking=True)
            target = target.to(self.device, non_blocking=True)
            if train:
                m = self.train_step(data, target, stage)
            else:
                final, base, logits, w, out = self._forward_loss(data, target, stage)
                correct = out[2] if out is not None else (logits.argmax(dim=1) == target).sum()
                zero = torch.zeros((), device=logits.device)
                m = torch.stack([final.double(), base.double(), correct.double(),
                                 (torch.mean(w) if w is not None else zero).double(),
                                 (torch.std(w) if w is not None and w.numel() > 1 else zero).double()])
            acc = m if acc is None else acc + m
            total += target.size(0)
            if train and log_every and batch_idx % log_every == 0:
                f, b, c, um, us = m.tolist()
                info = f", Unc_mean: {um:.4f}, Unc_std: {us:.4f}" if stage > 1 else ""
                print(f"Stage {stage}, Batch {batch_idx}/{len(dataloader)} Loss: {f:.6f} Base_Loss: {b:.6f} "
                      f"Acc: {100.0 * acc[2].item() / total:.2f}%{info}")
        n = len(dataloader)
        sums = (acc if acc is not None else torch.zeros(5, dtype=torch.float64)).double()
        if self.distributed:
            # Every rank must see the SAME epoch numbers: they drive ReduceLROnPlateau, the best-model
            # decision and early stopping (train_progressive), and BatchNorm running statistics are
            # rank-local, so per-rank validation losses differ.  Rank-local numbers would let ranks pick
            # different learning rates (silently diverging weights, graphs re-captured at different times)
            # or leave the epoch loop at different epochs (the others then block in the all-reduce).
            pack = torch.cat([sums.to(self.device), torch.tensor([float(total), float(n)], dtype=torch.float64,
                                                                 device=self.device)])
            dist.all_reduce(pack, op=dist.ReduceOp.SUM)
            sums, total, n = pack[:5], int(pack[5].item()), int(pack[6].item())
        n = max(n, 1)
        f, b, c, um, us = sums.tolist()
        return f / n, b / n, 100.0 * c / max(total, 1), um / n, us / n

    def train_epoch(self, dataloader, stage):
        """reference :298-357 — (avg_loss, avg_base_loss, accuracy %, unc mean, unc std)."""
        self.models[stage].train()
        if stage > 1:
            self.models[stage - 1].eval()
        return self._epoch(dataloader, stage, True)

    def validate_epoch(self, dataloader, stage):
        """reference :359-400."""
        self.models[stage].eval()
        if stage > 1:
            self.models[stage - 1].eval()
        with torch.no_grad():
            return self._epoch(dataloader, stage, False)

    def train_progressive(self, train_loaders, val_loaders, save_dir):
        """reference :402-497 (checkpoint dict layout and history file preserved)."""
        os.makedirs(save_dir, exist_ok=True)
        rank0 = (not self.distributed) or dist.get_rank() == 0
        for stage in self.config["stages"]:
            print(f"\n{'=' * 60}\nTraining Stage {stage} - Resolution: {self.stage_configs[stage]['resolution']}\n{'=' * 60}")
            self.current_stage = stage
            self.setup_optimizer_scheduler(stage)
            if stage > 1 and (stage - 1) in self.models:
                self.transfer_weights(stage - 1, stage)
            best_val_loss, best_val_acc, stale = float("inf"), 0, 0
            for epoch in range(self.stage_configs[stage]["epochs"]):
                print(f"\nStage {stage}, Epoch {epoch + 1}/{self.stage_configs[stage]['epochs']}")
                tl, tb, ta, tum, tus = self.train_epoch(train_loaders[stage], stage)
                vl, vb, va, vum, vus = self.validate_epoch(val_loaders[stage], stage)
                self.schedulers[stage].step(vl)
                for k, v in (("train_loss", tl), ("val_loss", vl), ("train_acc", ta), ("val_acc", va),
                             ("uncertainty_weights_mean", vum), ("uncertainty_weights_std", vus), ("base_loss", vb)):
                    self.history[k].append(v)
                print(f"Train Loss: {tl:.4f}, Base Loss: {tb:.4f}, Train Acc: {ta:.2f}%")
                print(f"Val Loss: {vl:.4f}, Base Loss: {vb:.4f}, Val Acc: {va:.2f}%")
                if va > best_val_acc:
                    best_val_loss, best_val_acc, stale = vl, va, 0
                    if rank0:
                        torch.save({"model_state_dict": self.models[stage].state_dict(),
                                    "optimizer_state_dict": self.optimizers[stage].state_dict(),
                                    "stage": stage, "epoch": epoch + 1, "train_loss": tl, "val_loss": vl,
                                    "train_acc": ta, "val_acc": va, "config": self.config},
                                   os.path.join(save_dir, f"herlev_stage{stage}_best.pth"))
                        print(f"New best model saved! Val Acc: {va:.2f}%")
                else:
                    stale += 1
                if stale >= self.config.get("early_stopping_patience", 15):
                    print(f"Early stopping after {stale} epochs without improvement")
                    break
            self.history["stage_transitions"].append({"stage": stage, "completed_at": datetime.now().isoformat(),
                                                      "best_val_acc": best_val_acc, "best_val_loss": best_val_loss})
            print(f"Stage {stage} completed. Best Val Acc: {best_val_acc:.2f}%")
        if rank0:
            with open(os.path.join(save_dir, "training_history.json"), "w") as f:
                json.dump(self.history, f, indent=2)
