"""Device-side evaluators — the compute of the reference's `MoNuSegEvaluator`
(MoNuSegImprove/test_monuseg.py:100-297) and `HerlevEvaluator` (Herlev/test_herlev.py:36-260) without
the per-image `.cpu()` + numpy round trips: thresholding, the confusion counts behind IoU / Dice /
accuracy / precision / recall / specificity, the mean-probability confidence, softmax / argmax and the
class confusion matrix are single-pass kernels over the logits (`ugp_seg_confusion`, `ugp_cls_eval`).

File handling (PIL, XML polygons, matplotlib, sklearn reports) stays with the caller: these classes take
tensors — fp32 NCHW batches as the reference's datasets yield them, or uint8 HWC image bytes, which
skip the float image entirely (`/255`, Normalize and the resize run inside the first-layer packer).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.nn as nn

from . import engine, ops
from .herlev import HerlevClassificationModel
from .unet import STAGE_MODELS, _PGUNetBase, _forward_resized

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # Herlev/test_herlev.py:122-123
IMAGENET_STD = (0.229, 0.224, 0.225)
_EPS = 1e-8                                 # test_monuseg.py:281


def metrics_from_counts(tp, n_pred, n_gt, n_pixels) -> Dict[str, torch.Tensor]:
    """calculate_metrics (test_monuseg.py:264-297) from the confusion counts; works on python floats
    and on float64 tensors (one entry per image)."""
    fp = n_pred - tp
    fn = n_gt - tp
    tn = n_pixels - tp - fp - fn
    return {
        "iou": (tp + _EPS) / (tp + fp + fn + _EPS),
        "dice": (2 * tp + _EPS) / (2 * tp + fp + fn + _EPS),
        "accuracy": (tp + tn + _EPS) / (tp + tn + fp + fn + _EPS),
        "precision": (tp + _EPS) / (tp + fp + _EPS),
        "recall": (tp + _EPS) / (tp + fn + _EPS),
        "specificity": (tn + _EPS) / (tn + fp + _EPS),
    }


def classification_metrics_from_confusion(confusion: torch.Tensor) -> Dict[str, object]:
    """What HerlevEvaluator.calculate_metrics derives from labels and predictions
    (Herlev/test_herlev.py:217-249: accuracy, sklearn's classification_report with zero_division=0,
    confusion_matrix, per-class accuracy) computed from the K x K confusion matrix
    counts[label][prediction] that `ugp_cls_eval` accumulates on the device."""
    cm = confusion.detach().cpu().double()
    total = float(cm.sum())
    tp = cm.diag()
    support = cm.sum(1)
    predicted = cm.sum(0)
    zero = torch.zeros_like(tp)
    precision = torch.where(predicted > 0, tp / predicted.clamp(min=1), zero)
    recall = torch.where(support > 0, tp / support.clamp(min=1), zero)
    f1 = torch.where(precision + recall > 0, 2 * precision * recall / (precision + recall).clamp(min=1e-300), zero)
    wsum = max(float(support.sum()), 1.0)
    return {
        "accuracy": float(tp.sum()) / max(total, 1.0),
        "confusion_matrix": confusion.detach().cpu(),
        "precision": precision.tolist(), "recall": recall.tolist(), "f1": f1.tolist(),
        "support": support.long().tolist(),
        # classes absent from the labels have no per-class accuracy entry in the reference (:243-247)
        "per_class_accuracy": {i: float(recall[i]) for i in range(cm.shape[0]) if support[i] > 0},
        "macro_avg": {"precision": float(precision.mean()), "recall": float(recall.mean()), "f1": float(f1.mean())},
        "weighted_avg": {"precision": float((precision * support).sum() / wsum),
                         "recall": float((recall * support).sum() / wsum),
                         "f1": float((f1 * support).sum() / wsum)},
        "total_samples": int(total),
    }


class MoNuSegEvaluator:
    """Mirror of MoNuSegImprove/test_monuseg.py:100-297 for tensors.

    `model_or_path`: a PGUNet module, or a checkpoint path in either of the reference's two formats
    (`{'model_state_dict', 'stage', ...}` or a raw state_dict, assumed stage 4; :120-152)."""

    def __init__(self, model_or_path, device: str = "cuda", use_cuda_graph: bool = True):
        self.device = torch.device(device)
        self.model = model_or_path if isinstance(model_or_path, nn.Module) else self.load_model(model_or_path)
        self.model.to(self.device).eval()
        # The inference forward is ~60 short kernels: at batch 1 an eager pass is bound by the host's
        # launch rate, so `predict_batch` replays one CUDA graph per input signature (forward +
        # thresholding / confidence kernel).  The bf16 operand packs and folded BN coefficients the graph
        # reads are refreshed in place before every replay, so in-place weight updates (load_state_dict,
        # optimizer steps) are honoured; `reset_graphs()` is needed only if parameter storage was replaced
        # (`.to()` to another device, swapping modules).
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}

    def reset_graphs(self) -> None:
        self._graphs = {}

    def _predict_graphed(self, images: torch.Tensor, target_size: Optional[int]):
        key = (tuple(images.shape), images.dtype, target_size)
        g = self._graphs.get(key)
        if g is None:
            self._graphs[key] = g = {"warm": 0, "graph": None}
        if g["graph"] is None:
            if g["warm"] < 2:                      # eager passes first: bf16 weight packs, BN coefficients, autotune-free
                g["warm"] += 1
                z = self.logits(images, target_size)
                return z, ops.seg_confusion(z.contiguous(), None, want_mask=True)
            g["x"] = images.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                z = self.logits(g["x"], target_size)
                g["out"] = (z, ops.seg_confusion(z.contiguous(), None, want_mask=True))
            g["graph"] = graph
        engine.refresh_eval_caches(self.model)     # weight / BN updates since the capture (in place, no-op if none)
        g["x"].copy_(images, non_blocking=True)
        g["graph"].replay()
        z, (counts, mask) = g["out"]
        return z, (counts.clone(), mask.clone())   # static buffers: the next replay overwrites them

    def load_model(self, model_path: str) -> nn.Module:
        checkpoint = torch.load(model_path, map_location="cpu")
        if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint:
            stage, state_dict = checkpoint.get("stage", 4), checkpoint["model_state_dict"]
        elif isinstance(checkpoint, dict) and all(isinstance(v, torch.Tensor) for v in checkpoint.values()):
            stage, state_dict = 4, checkpoint
        else:
            raise RuntimeError(f"Unrecognized checkpoint format for: {model_path}")
        model = STAGE_MODELS.get(stage, STAGE_MODELS[4])(in_channels=3, num_classes=1)
        model.load_state_dict(state_dict)
        return model

    @torch.no_grad()
    def logits(self, images: torch.Tensor, target_size: Optional[int] = None) -> torch.Tensor:
        """fp32 NCHW batch or uint8 HWC bytes [B,H,W,3] -> fp32 logits [B,1,S,S]; `target_size` resizes
        inside the first-layer packer (predict_image resizes to 256 before the model, :176-181)."""
        images = images.to(self.device, non_blocking=True)
        if isinstance(self.model, _PGUNetBase):
            size = target_size if target_size is not None else (images.shape[1] if images.dtype == torch.uint8
                                                                else images.shape[-1])
            return _forward_resized(self.model, images, size)
        return self.model(images.float())

    @torch.no_grad()
    def predict_batch(self, images: torch.Tensor, target_size: Optional[int] = None,
                      out_size: Optional[Tuple[int, int]] = None):
        """-> (uint8 masks [B,1,h,w], confidences float64 [B]) — predict_image (:163-201) for a batch:
        mask = sigmoid(z) > 0.5, confidence = mean sigmoid(z); `out_size` applies the reference's
        nearest resize back to the original image size (:191-194)."""
        images = images.to(self.device, non_blocking=True)
        if self.use_cuda_graph and isinstance(self.model, _PGUNetBase):
            z, (counts, mask) = self._predict_graphed(images, target_size)
        else:
            z = self.logits(images, target_size)
            counts, mask = ops.seg_confusion(z.contiguous(), None, want_mask=True)
        conf = counts[:, 3] / (z.numel() // z.shape[0])
        if out_size is not None and tuple(mask.shape[-2:]) != tuple(out_size):
            mask = nn.functional.interpolate(mask, size=tuple(out_size), mode="nearest")
        return mask, conf

    def calculate_metrics(self, pred_mask: torch.Tensor, gt_mask: torch.Tensor) -> Dict[str, float]:
        """calculate_metrics (:264-297) for one image given as 0/1 tensors (any dtype / device)."""
        p = pred_mask.to(self.device).float().flatten()
        g = gt_mask.to(self.device).float().flatten()
        tp = float((p * g).sum())
        return {k: float(v) for k, v in metrics_from_counts(tp, float(p.sum()), float(g.sum()), p.numel()).items()}

    @torch.no_grad()
    def evaluate_batches(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]):
        """evaluate_dataset (:203-262) over an iterable of (images, masks) batches: per-image metrics
        from one fused pass over the logits, then mean and std over images (np.mean / np.std: population
        std).  Returns (avg_metrics, std_metrics).  Nothing is read back until the end."""
        per_image = []
        for images, masks in batches:
            z = self.logits(images).contiguous()
            gt = masks.to(self.device, non_blocking=True)
            if gt.dtype not in (torch.uint8, torch.float32):
                gt = gt.float()
            counts, _ = ops.seg_confusion(z, gt.contiguous().view(z.shape[0], -1))
            n = z.numel() // z.shape[0]
            per_image.append((counts, n))
        if not per_image:
            raise ValueError("evaluate_batches: no batches")
        rows = {k: [] for k in ("iou", "dice", "accuracy", "precision", "recall", "specificity")}
        for counts, n in per_image:
            m = metrics_from_counts(counts[:, 0], counts[:, 1], counts[:, 2], float(n))
            for k in rows:
                rows[k].append(m[k])
        avg, std = {}, {}
        for k, parts in rows.items():
            v = torch.cat(parts)
            avg[k] = float(v.mean())
            std[k] = float(v.std(unbiased=False))
        return avg, std


class HerlevEvaluator:
    """Mirror of Herlev/test_herlev.py:36-260 for tensors: batched softmax / argmax / confusion matrix
    on the device.  `model`: a HerlevClassificationModel (or any module returning [B,K] logits)."""

    def __init__(self, model: nn.Module, device: str = "cuda", normalize_bytes: bool = True):
        self.device = torch.device(device)
        self.model = model.to(self.device).eval()
        self.num_classes = getattr(model, "num_classes", None)
        if normalize_bytes and isinstance(model, HerlevClassificationModel):
            model.unet.set_input_norm(IMAGENET_MEAN, IMAGENET_STD)     # test_herlev.py:119-124

    @torch.no_grad()
    def logits(self, images: torch.Tensor, image_size: Optional[int] = None) -> torch.Tensor:
        images = images.to(self.device, non_blocking=True)
        if isinstance(self.model, HerlevClassificationModel):
            if image_size is None:
                image_size = images.shape[1] if images.dtype == torch.uint8 else images.shape[-1]
            return self.model.forward_resized(images, image_size).float().contiguous()
        return self.model(images.float()).float().contiguous()

    @torch.no_grad()
    def predict_batch(self, images: torch.Tensor, image_size: Optional[int] = None):
        """-> (predicted classes int64 [B], probabilities fp32 [B,K], confidences fp32 [B]) —
        predict_single (:104-135) for a batch."""
        probs, preds, _ = ops.cls_eval(self.logits(images, image_size))
        return preds, probs, probs.gather(1, preds[:, None]).squeeze(1)

    @torch.no_grad()
    def evaluate_batches(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]):
        """evaluate_dataset (:137-215): accumulates the K x K confusion matrix on the device and derives
        accuracy and per-class precision / recall / F1 from it (what sklearn's report computes at
        :217-260).  Returns a dict with python / CPU values."""
        confusion = None
        total = 0
        for images, labels in batches:
            z = self.logits(images)
            labels = labels.to(self.device, non_blocking=True).long().contiguous()
            _, _, confusion = ops.cls_eval(z, labels, confusion)
            total += z.shape[0]
        if confusion is None:
            raise ValueError("evaluate_batches: no batches")
        return classification_metrics_from_confusion(confusion)
