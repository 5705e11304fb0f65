"""CPU oracle for the UG-PG-UNet hot path — TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement, in plain torch-on-CPU ops, of the reference's
algorithm for the path this repository accelerates.  It is the checker for
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference
arm; nothing under `ug_pg_unet_b200/` imports it and the product path never
routes through it.

Pinning: `tests/golden/*.pt` were produced by importing the UNMODIFIED
reference (`/root/reference`, PyTorch modules) in the build container with
`tools/make_goldens.py` — five trainer steps (stages 1-4, BASELINE config 1
shape included), two Herlev steps, primitive semantics (max-pool indices,
bilinear resize, BCE pos_weight, uncertainty map, weight transfer), the
legacy `UNet` / `UpSample` (ConvTranspose2d) forward + backward, and the
`MoNuSegEvaluator` / `HerlevEvaluator` metric arithmetic; `tests/test_oracle_golden.py` checks every
function here against those vectors.  The reference itself ships no tests or golden
vectors, so these reference-generated fixtures are the pin.

Every function cites the reference lines it restates (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

STAGE_RESOLUTION = {1: 32, 2: 64, 3: 128, 4: 256}   # uncertainty_guided_trainer.py:42-47
_UP_OUT = [256, 128, 64, 64]                          # UG_unet.py:184,204-205,236-238,274-277


# --------------------------------------------------------------------------- architecture
def encoder_widths(stage: int) -> List[int]:
    """Channel widths inc -> ... -> down4 (UG_unet.py:182-183, 200-202, 230-233, 267-271)."""
    return {1: [512, 512], 2: [256, 512, 512], 3: [128, 256, 512, 512], 4: [64, 128, 256, 512, 512]}[stage]


def down_names(stage: int) -> List[str]:
    return [f"down{k}" for k in range(5 - (stage), 5)]      # stage 1: down4; stage 4: down1..down4


def block_specs(stage: int, in_channels: int) -> List[Tuple[str, int, int]]:
    """(state_dict prefix of the DoubleConv, Cin, Cout) in forward order."""
    enc = encoder_widths(stage)
    specs = [("inc.conv", in_channels, enc[0])]
    for i, name in enumerate(down_names(stage)):
        specs.append((f"{name}.mpconv.1", enc[i], enc[i + 1]))
    prev = enc[-1]
    for j in range(1, stage + 1):
        skip = enc[stage - j]
        specs.append((f"up{j}.conv", skip + prev, _UP_OUT[j - 1]))
        prev = _UP_OUT[j - 1]
    return specs


def head_names(stage: int) -> List[str]:
    return ["outc"] if stage == 1 else [f"outc{j}" for j in range(1, stage + 1)]   # UG_unet.py:185,206-207


def state_dict_shapes(stage: int, in_channels: int, num_classes: int) -> Dict[str, Tuple[int, ...]]:
    """Every key/shape of PGUNet{stage}.state_dict() (UG_unet_parts.py:9-16,88)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    for prefix, cin, cout in block_specs(stage, in_channels):
        for conv_i, bn_i, ci in ((0, 1, cin), (3, 4, cout)):
            shapes[f"{prefix}.conv_op.{conv_i}.weight"] = (cout, ci, 3, 3)
            shapes[f"{prefix}.conv_op.{conv_i}.bias"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.weight"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.bias"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.running_mean"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.running_var"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.num_batches_tracked"] = ()
    for j, name in enumerate(head_names(stage)):
        c = _UP_OUT[stage - 1] if stage == 1 else _UP_OUT[j]
        shapes[f"{name}.conv.weight"] = (num_classes, c, 1, 1)
        shapes[f"{name}.conv.bias"] = (num_classes,)
    return shapes


def seeded_state_dict(stage: int, in_channels: int, num_classes: int, seed: int = 0) -> StateDict:
    """Deterministic weights that depend only on (key, shape, seed) — identical wherever they are
    regenerated (golden generator, tests, bench), with no dependence on module construction order."""
    sd: StateDict = {}
    for key, shape in state_dict_shapes(stage, in_channels, num_classes).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in) * 1.4
        elif ".conv_op.1." in key or ".conv_op.4." in key:
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75 if key.endswith("weight") \
                else torch.randn(shape, generator=g) * 0.1
        else:   # conv / outconv biases
            sd[key] = torch.randn(shape, generator=g) * 0.05
    return sd


# --------------------------------------------------------------------------- bf16 storage emulation
class _RoundBF16(torch.autograd.Function):
    """Round-to-bf16 in forward AND backward: models a tensor (and its gradient) being stored in
    bf16 between kernels while all arithmetic stays fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _q(x: Tensor, bf16_storage: bool) -> Tensor:
    return _RoundBF16.apply(x) if bf16_storage else x


def _tf32(x: Tensor) -> Tensor:
    """Round-to-nearest to TF32's 10-bit mantissa (what a `kind::tf32` tensor-core operand keeps)."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _RoundTF32Fwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundTF32Bwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _tf32(g)


# set by `tf32_operands(True)`: fp32 storage everywhere, but every convolution operand (activation,
# weight, and in backward the output gradient) is rounded to TF32 - the arithmetic of cuDNN's default
# TF32 convolutions and of a `tcgen05.mma.kind::tf32` path.  Only used to show where north_star's
# alternative 1e-3 (TF32) bar lands on these networks; no kernel implements it.
_TF32_OPERANDS = False


class tf32_operands:
    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        global _TF32_OPERANDS
        self.prev, _TF32_OPERANDS = _TF32_OPERANDS, self.on

    def __exit__(self, *a):
        global _TF32_OPERANDS
        _TF32_OPERANDS = self.prev


# set by `bf16_storage(True)`: every activation / conv weight is rounded to bf16 where the sm_100a
# path stores it in bf16 (conv inputs, conv outputs, post-ReLU activations, upsampled tensors and
# their gradients); arithmetic stays fp32.  Used to separate kernel bugs from precision effects.
_BF16_STORAGE = False


# A/B of a kernel design choice: train-mode BatchNorm statistics taken from the fp32
# accumulators BEFORE the conv output is rounded to bf16 (True) instead of from the stored bf16 values
# (False = what conv_epilogue.cuh does).  Only meaningful inside bf16_storage(True).
_BN_STATS_FROM_FP32 = False


class bf16_storage:
    def __init__(self, on: bool = True, stats_from_fp32: bool = False):
        self.on, self.stats_from_fp32 = on, stats_from_fp32

    def __enter__(self):
        global _BF16_STORAGE, _BN_STATS_FROM_FP32
        self.prev = (_BF16_STORAGE, _BN_STATS_FROM_FP32)
        _BF16_STORAGE, _BN_STATS_FROM_FP32 = self.on, self.stats_from_fp32

    def __exit__(self, *a):
        global _BF16_STORAGE, _BN_STATS_FROM_FP32
        _BF16_STORAGE, _BN_STATS_FROM_FP32 = self.prev


# --------------------------------------------------------------------------- blocks
def double_conv(x: Tensor, sd: StateDict, prefix: str, training: bool, bn_out: Optional[StateDict] = None,
                momentum: float = 0.1, eps: float = 1e-5) -> Tensor:
    """(Conv3x3 pad1 + bias -> BatchNorm2d -> ReLU) x 2 — UG_unet_parts.py:5-19.
    In training mode batch statistics normalise (biased variance) and, when `bn_out` is given,
    the updated running statistics (unbiased variance, momentum 0.1) are written there."""
    bq = _BF16_STORAGE
    x = _q(x, bq)
    for conv_i, bn_i in ((0, 1), (3, 4)):
        w = sd[f"{prefix}.conv_op.{conv_i}.weight"]
        if _TF32_OPERANDS:
            x = _RoundTF32Bwd.apply(F.conv2d(_RoundTF32Fwd.apply(x), _RoundTF32Fwd.apply(w),
                                             sd[f"{prefix}.conv_op.{conv_i}.bias"], padding=1))
        else:
            x = F.conv2d(x, _q(w, bq), sd[f"{prefix}.conv_op.{conv_i}.bias"], padding=1)
        x_acc = x
        if bq and training:
            x = _q(x, True)        # train: raw conv output stored in bf16; eval: BN+ReLU fused in the epilogue
        rm = sd[f"{prefix}.conv_op.{bn_i}.running_mean"].clone()
        rv = sd[f"{prefix}.conv_op.{bn_i}.running_var"].clone()
        if bq and training and _BN_STATS_FROM_FP32:
            # design A/B: normalise the STORED (rounded) values with the statistics of the unrounded ones
            mean = x_acc.mean(dim=(0, 2, 3), keepdim=True)
            var = x_acc.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
            g_ = sd[f"{prefix}.conv_op.{bn_i}.weight"].view(1, -1, 1, 1)
            b_ = sd[f"{prefix}.conv_op.{bn_i}.bias"].view(1, -1, 1, 1)
            x = (x - mean) * torch.rsqrt(var + eps) * g_ + b_
        else:
            x = F.batch_norm(x, rm, rv, sd[f"{prefix}.conv_op.{bn_i}.weight"], sd[f"{prefix}.conv_op.{bn_i}.bias"],
                             training=training, momentum=momentum, eps=eps)
        if training and bn_out is not None:
            bn_out[f"{prefix}.conv_op.{bn_i}.running_mean"] = rm
            bn_out[f"{prefix}.conv_op.{bn_i}.running_var"] = rv
            bn_out[f"{prefix}.conv_op.{bn_i}.num_batches_tracked"] = sd[f"{prefix}.conv_op.{bn_i}.num_batches_tracked"] + 1
        x = _q(F.relu(x), bq)
    return x


def up_block(x_low: Tensor, x_skip: Tensor, sd: StateDict, prefix: str, training: bool, bn_out=None) -> Tensor:
    """bilinear x2 (align_corners=True) -> cat([skip, up]) -> DoubleConv — UG_unet_parts.py:70-81."""
    up = _q(F.interpolate(x_low, scale_factor=2, mode="bilinear", align_corners=True), _BF16_STORAGE)
    return double_conv(torch.cat([x_skip, up], dim=1), sd, prefix, training, bn_out)


def pgunet_forward(sd: StateDict, x: Tensor, stage: int, training: bool = True, bn_out: Optional[StateDict] = None,
                   return_features: bool = False):
    """PGUNet{stage}.forward — UG_unet.py:188-193, 209-223, 242-260, 282-304."""
    feats = [double_conv(x, sd, "inc.conv", training, bn_out)]                       # InConv, UG_unet_parts.py:21-28
    for name in down_names(stage):                                                    # Down, UG_unet_parts.py:44-54
        feats.append(double_conv(F.max_pool2d(feats[-1], 2), sd, f"{name}.mpconv.1", training, bn_out))
    dec = []
    cur = feats[-1]
    for j in range(1, stage + 1):
        cur = up_block(cur, feats[stage - j], sd, f"up{j}.conv", training, bn_out)
        dec.append(cur)
    logits = None
    for j, name in enumerate(head_names(stage)):                                      # OutConv, UG_unet_parts.py:84-91
        d = dec[-1] if stage == 1 else dec[j]
        l = F.conv2d(d, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"])
        factor = x.shape[-1] // l.shape[-1]
        if factor > 1:                                                                # UG_unet.py:218,254-255,299-301
            l = F.interpolate(l, scale_factor=factor, mode="bilinear", align_corners=True)
        logits = l if logits is None else logits + l
    if return_features:
        return logits, feats, dec
    return logits


# --------------------------------------------------------------------------- legacy UNet (ConvTranspose2d up path)
_LEGACY_DOWN = [("down_convolution_1", None, 64), ("down_convolution_2", 64, 128),
                ("down_convolution_3", 128, 256), ("down_convolution_4", 256, 512)]   # UG_unet.py:435-438
_LEGACY_UP = [("up_convolution_1", 1024, 512), ("up_convolution_2", 512, 256),
              ("up_convolution_3", 256, 128), ("up_convolution_4", 128, 64)]           # UG_unet.py:442-445


def legacy_unet_shapes(in_channels: int, num_classes: int) -> Dict[str, Tuple[int, ...]]:
    """Every key/shape of UNet(in_channels, num_classes).state_dict() — UG_unet.py:430-447,
    UG_unet_parts.py:31-41 (DownSample), :57-67 (UpSample)."""
    shapes: Dict[str, Tuple[int, ...]] = {}

    def dconv(prefix, cin, cout):
        for conv_i, bn_i, ci in ((0, 1, cin), (3, 4, cout)):
            shapes[f"{prefix}.conv_op.{conv_i}.weight"] = (cout, ci, 3, 3)
            shapes[f"{prefix}.conv_op.{conv_i}.bias"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.weight"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.bias"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.running_mean"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.running_var"] = (cout,)
            shapes[f"{prefix}.conv_op.{bn_i}.num_batches_tracked"] = ()

    for name, cin, cout in _LEGACY_DOWN:
        dconv(f"{name}.conv", in_channels if cin is None else cin, cout)
    dconv("bottle_neck", 512, 1024)
    for name, cin, cout in _LEGACY_UP:
        shapes[f"{name}.up.weight"] = (cin, cin // 2, 2, 2)      # nn.ConvTranspose2d: [Cin, Cout, kH, kW]
        shapes[f"{name}.up.bias"] = (cin // 2,)
        dconv(f"{name}.conv", cin, cout)
    shapes["out.weight"] = (num_classes, 64, 1, 1)
    shapes["out.bias"] = (num_classes,)
    return shapes


def seeded_legacy_state_dict(in_channels: int, num_classes: int, seed: int = 0) -> StateDict:
    """Deterministic legacy-UNet weights, a pure function of (key, shape, seed) like seeded_state_dict."""
    sd: StateDict = {}
    for key, shape in legacy_unet_shapes(in_channels, num_classes).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith(".up.weight"):
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / shape[0]) * 1.4
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in) * 1.4
        elif ".conv_op.1." in key or ".conv_op.4." in key:
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75 if key.endswith("weight") \
                else torch.randn(shape, generator=g) * 0.1
        else:
            sd[key] = torch.randn(shape, generator=g) * 0.05
    return sd


def conv_transpose2x2(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) — UG_unet_parts.py:61,65:
    out[n, co, 2y+i, 2x+j] = b[co] + sum_ci x[n, ci, y, x] * w[ci, co, i, j]."""
    bq = _BF16_STORAGE
    return _q(F.conv_transpose2d(_q(x, bq), _q(w, bq), b, stride=2), bq)


def legacy_unet_forward(sd: StateDict, x: Tensor, training: bool = True, bn_out: Optional[StateDict] = None) -> Tensor:
    """UNet.forward — UG_unet.py:449-462."""
    skips = []
    cur = x
    for name, _, _ in _LEGACY_DOWN:                                    # DownSample, UG_unet_parts.py:38-41
        d = double_conv(cur, sd, f"{name}.conv", training, bn_out)
        skips.append(d)
        cur = F.max_pool2d(d, 2)
    cur = double_conv(cur, sd, "bottle_neck", training, bn_out)
    for (name, _, _), skip in zip(_LEGACY_UP, reversed(skips)):         # UpSample, UG_unet_parts.py:64-67
        up = conv_transpose2x2(cur, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"])
        cur = double_conv(torch.cat([up, skip], dim=1), sd, f"{name}.conv", training, bn_out)
    return F.conv2d(cur, sd["out.weight"], sd["out.bias"])


# --------------------------------------------------------------------------- uncertainty + loss
def uncertainty_map(x: Tensor, sd_prev: StateDict, prev_stage: int, cur_res: int) -> Tensor:
    """A = 1 - 2|up(sigmoid(M_prev(down(x)))) - 0.5|, previous model in eval mode — UG_unet.py:19-59."""
    prev_res = STAGE_RESOLUTION[prev_stage]
    with torch.no_grad():
        xp = F.interpolate(x, size=(prev_res, prev_res), mode="bilinear", align_corners=True)
        p = torch.sigmoid(pgunet_forward(sd_prev, xp, prev_stage, training=False))
        p = F.interpolate(p, size=(cur_res, cur_res), mode="bilinear", align_corners=True)
        return 1.0 - 2.0 * torch.abs(p - 0.5)


def bce_with_logits(z: Tensor, y: Tensor, pos_weight: float) -> Tensor:
    """BCEWithLogitsLoss(pos_weight, reduction='none') closed form —
    uncertainty_guided_trainer.py:64-65: l = (1-y) z + (1+(pw-1) y) softplus(-z)."""
    return (1 - y) * z + (1 + (pos_weight - 1) * y) * (torch.log1p(torch.exp(-z.abs())) + torch.clamp(-z, min=0))


def weighted_loss(z: Tensor, y: Tensor, A: Optional[Tensor], alpha: float, pos_weight: float = 5.0):
    """final = mean(l * (1 + alpha A)) (or mean(l) at stage 1), base = mean(l) — UG_unet.py:61-94."""
    pl = bce_with_logits(z, y, pos_weight)
    final = pl.mean() if A is None else (pl * (1.0 + alpha * A).detach()).mean()
    return final, pl.mean()


def dice_coefficient(pred: Tensor, target: Tensor, smooth: float = 1.0) -> Tensor:
    """uncertainty_guided_trainer.py:90-107."""
    p = pred.reshape(pred.shape[0], -1).float()
    t = target.reshape(target.shape[0], -1).float()
    return ((2.0 * (p * t).sum(1) + smooth) / (p.sum(1) + t.sum(1) + smooth)).mean()


def predictions(z: Tensor) -> Tensor:
    """sigmoid(z) > 0.5 — uncertainty_guided_trainer.py:109-113."""
    return (torch.sigmoid(z) > 0.5).float().squeeze(1)


def accuracy(pred: Tensor, target: Tensor) -> float:
    """uncertainty_guided_trainer.py:115-123."""
    return 1.0 - pred.ne(target).sum().item() / pred.numel()


def seg_metrics(pred: Tensor, gt: Tensor) -> Dict[str, float]:
    """IoU / Dice / accuracy / precision / recall / specificity of one 0/1 prediction against one 0/1
    mask with eps = 1e-8 — MoNuSegEvaluator.calculate_metrics, MoNuSegImprove/test_monuseg.py:264-297."""
    p, g = pred.double().flatten(), gt.double().flatten()
    tp = float((p * g).sum())
    fp = float(p.sum()) - tp
    fn = float(g.sum()) - tp
    tn = p.numel() - tp - fp - fn
    eps = 1e-8
    return {"iou": (tp + eps) / (tp + fp + fn + eps), "dice": (2 * tp + eps) / (2 * tp + fp + fn + eps),
            "accuracy": (tp + tn + eps) / (tp + tn + fp + fn + eps), "precision": (tp + eps) / (tp + fp + eps),
            "recall": (tp + eps) / (tp + fn + eps), "specificity": (tn + eps) / (tn + fp + eps)}


def evaluate_segmentation(sd: StateDict, stage: int, images_u8: Tensor, gts: Tensor):
    """MoNuSegEvaluator.evaluate_dataset / predict_image on tensors — test_monuseg.py:176-201, 236-258:
    x = bytes/255 (HWC -> CHW), eval-mode logits, mask = sigmoid > 0.5, confidence = mean sigmoid,
    per-image metrics, then np.mean / np.std (population) over images.
    Returns (masks [B,1,H,W] float, confidences [B], per-image metric dicts, avg, std, logits)."""
    x = images_u8.permute(0, 3, 1, 2).float() / 255.0
    with torch.no_grad():
        logits = pgunet_forward(sd, x, stage, training=False)
    probs = torch.sigmoid(logits)
    masks = (probs > 0.5).float()
    conf = probs.flatten(1).mean(1)
    per = [seg_metrics(masks[i], gts[i]) for i in range(x.shape[0])]
    avg = {k: sum(m[k] for m in per) / len(per) for k in per[0]}
    std = {k: math.sqrt(sum((m[k] - avg[k]) ** 2 for m in per) / len(per)) for k in per[0]}
    return masks, conf, per, avg, std, logits


def classification_metrics(y_true: Tensor, y_pred: Tensor, num_classes: int) -> Dict[str, object]:
    """HerlevEvaluator.calculate_metrics — Herlev/test_herlev.py:217-249: accuracy = mean(y_true == y_pred);
    sklearn classification_report(zero_division=0): per class precision = tp/(tp+fp), recall = tp/(tp+fn),
    f1 = 2pr/(p+r) (0 where the denominator is 0), support = #labels, macro / weighted averages;
    confusion_matrix[label][prediction]; per-class accuracy for the classes present in the labels."""
    K = num_classes
    cm = torch.zeros(K, K, dtype=torch.int64)
    for t, p in zip(y_true.tolist(), y_pred.tolist()):
        cm[t, p] += 1
    per = {}
    for k in range(K):
        tp = int(cm[k, k]); fp = int(cm[:, k].sum()) - tp; fn = int(cm[k, :].sum()) - tp
        pr = tp / (tp + fp) if tp + fp else 0.0
        rc = tp / (tp + fn) if tp + fn else 0.0
        per[k] = {"precision": pr, "recall": rc, "f1": 2 * pr * rc / (pr + rc) if pr + rc else 0.0, "support": tp + fn}
    n = max(len(y_true), 1)
    macro = {m: sum(per[k][m] for k in range(K)) / K for m in ("precision", "recall", "f1")}
    weighted = {m: sum(per[k][m] * per[k]["support"] for k in range(K)) / n for m in ("precision", "recall", "f1")}
    return {"accuracy": float((y_true == y_pred).double().mean()), "per_class": per, "confusion_matrix": cm,
            "macro_avg": macro, "weighted_avg": weighted,
            "per_class_accuracy": {k: per[k]["recall"] for k in range(K) if per[k]["support"] > 0}}


def softmax_entropy_uncertainty(logits: Tensor) -> Tensor:
    """u = H(softmax(z)) / ln K per sample — Herlev/train_herlev.py:253-257."""
    p = F.softmax(logits, dim=1)
    return -(p * torch.log(p + 1e-8)).sum(1) / math.log(logits.shape[1])


def weighted_cross_entropy(logits: Tensor, labels: Tensor, u: Optional[Tensor], alpha: float) -> Tensor:
    """mean(CE_i * (1 + alpha u_i)) — Herlev/train_herlev.py:274-281."""
    ce = F.cross_entropy(logits, labels, reduction="none")
    return ce.mean() if u is None else (ce * (1.0 + alpha * u).detach()).mean()


# --------------------------------------------------------------------------- Herlev variant
def seeded_herlev_state_dict(stage: int, num_classes: int, seed: int = 0) -> StateDict:
    """`HerlevClassificationModel(stage, num_classes).state_dict()` layout (Herlev/train_herlev.py:47,
    66-77): `unet.*` = PGUNet{stage}(3, 1), `classifier.{3,6,9}.{weight,bias}` = the three Linear layers."""
    sd = {"unet." + k: v for k, v in seeded_state_dict(stage, 3, 1, seed).items()}
    for idx, (fo, fi) in ((3, (512, 512)), (6, (256, 512)), (9, (num_classes, 256))):
        for kind, shape in (("weight", (fo, fi)), ("bias", (fo,))):
            key = f"classifier.{idx}.{kind}"
            g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
            bound = 1.0 / math.sqrt(fi)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def herlev_forward(sd: StateDict, x: Tensor, stage: int, training: bool = True,
                   bn_out: Optional[StateDict] = None) -> Tensor:
    """HerlevClassificationModel.forward with the Dropout layers inactive — encoder blocks up to
    `down3` (Herlev/train_herlev.py:83-102), AdaptiveAvgPool2d(1)+Flatten, Linear-ReLU-Linear-ReLU-
    Linear (:66-77)."""
    usd = {k[5:]: v for k, v in sd.items() if k.startswith("unet.")}
    ubn: Optional[StateDict] = {} if bn_out is not None else None
    f = double_conv(x, usd, "inc.conv", training, ubn)
    for name in down_names(stage)[: stage - 1]:
        f = double_conv(F.max_pool2d(f, 2), usd, f"{name}.mpconv.1", training, ubn)
    if bn_out is not None:
        bn_out.update({"unet." + k: v for k, v in ubn.items()})
    h = f.mean(dim=(2, 3))
    h = F.relu(F.linear(h, sd["classifier.3.weight"], sd["classifier.3.bias"]))
    h = F.relu(F.linear(h, sd["classifier.6.weight"], sd["classifier.6.bias"]))
    return F.linear(h, sd["classifier.9.weight"], sd["classifier.9.bias"])


def herlev_train_step(sd: StateDict, x: Tensor, y: Tensor, stage: int, sd_prev: Optional[StateDict] = None,
                      alpha: float = 1.0, prev_resolution: Optional[int] = None):
    """HerlevTrainer.uncertainty_guided_forward_pass + backward for num_classes > 2, no class weights —
    Herlev/train_herlev.py:219-296 (previous stage in eval mode on the bilinearly down-sized input)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
              and not k.endswith(("running_mean", "running_var"))}
    full = dict(sd)
    full.update(params)
    bn_out: StateDict = {}
    logits = herlev_forward(full, x, stage, training=True, bn_out=bn_out)
    u = None
    if stage > 1:
        assert sd_prev is not None
        pr = STAGE_RESOLUTION[stage - 1] if prev_resolution is None else prev_resolution   # stage_configs, :134-139
        with torch.no_grad():
            xp = F.interpolate(x, size=(pr, pr), mode="bilinear", align_corners=True)
            u = softmax_entropy_uncertainty(herlev_forward(sd_prev, xp, stage - 1, training=False))
    base = F.cross_entropy(logits, y)
    final = weighted_cross_entropy(logits, y, u, alpha)
    final.backward()
    w = None if u is None else 1.0 + alpha * u
    return {
        "logits": logits.detach(), "final_loss": final.item(), "base_loss": base.item(), "u": u,
        "grads": {k: p.grad for k, p in params.items()}, "bn": bn_out,
        "correct": int((logits.argmax(1) == y).sum()),
        "uncertainty_weight_mean": 0.0 if w is None else w.mean().item(),
        "uncertainty_weight_std": 0.0 if w is None else w.std().item(),
    }


# --------------------------------------------------------------------------- trainer step
def train_step(sd: StateDict, x: Tensor, y: Tensor, stage: int, sd_prev: Optional[StateDict] = None,
               alpha: float = 1.0, pos_weight: float = 5.0):
    """One `uncertainty_guided_forward_pass` + backward — uncertainty_guided_trainer.py:143-184,214-217.
    Returns a dict with logits, losses, metrics, gradients per state_dict key and updated BN buffers."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
              and not k.endswith(("running_mean", "running_var"))}
    full = dict(sd)
    full.update(params)
    bn_out: StateDict = {}
    logits = pgunet_forward(full, x, stage, training=True, bn_out=bn_out)
    A = None
    if stage > 1:
        assert sd_prev is not None
        A = uncertainty_map(x, sd_prev, stage - 1, STAGE_RESOLUTION[stage])
    final, base = weighted_loss(logits, y, A, alpha, pos_weight)
    final.backward()
    out = {
        "logits": logits.detach(), "final_loss": final.item(), "base_loss": base.item(),
        "uncertainty": A, "grads": {k: p.grad for k, p in params.items()}, "bn": bn_out,
        "uncertainty_weight_mean": 0.0 if A is None else A.mean().item(),
        "uncertainty_weight_std": 0.0 if A is None else A.std().item(),     # unbiased, as torch.std
    }
    if logits.shape[1] == 1:
        pred = predictions(logits.detach())
        out["dice"] = dice_coefficient(pred, y.squeeze(1)).item()
        out["accuracy"] = accuracy(pred, y.squeeze(1))
    return out


def rmsprop_step(p: Tensor, g: Tensor, v: Tensor, lr: float, alpha: float = 0.99, eps: float = 1e-8,
                 weight_decay: float = 1e-4) -> None:
    """torch.optim.RMSprop as configured at uncertainty_guided_trainer.py:84-88 (in place)."""
    g = g + weight_decay * p
    v.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    p.addcdiv_(g, v.sqrt().add_(eps), value=-lr)


def transfer_weights(prev: StateDict, cur: StateDict) -> Tuple[StateDict, List[str]]:
    """Same-key copy with leading-slice partial copy on shape mismatch — UG_unet.py:345-411."""
    new = {k: v.clone() for k, v in cur.items()}
    copied = []
    for k, pv in prev.items():
        if k not in cur:
            continue
        cv = cur[k]
        if pv.shape == cv.shape:
            new[k] = pv.clone()
        elif pv.ndim == cv.ndim and pv.ndim in (1, 2, 4):
            sl = tuple(slice(0, min(a, b)) for a, b in zip(pv.shape[:2], cv.shape[:2])) if pv.ndim > 1 \
                else (slice(0, min(pv.shape[0], cv.shape[0])),)
            t = cv.clone()
            t[sl] = pv[sl]
            new[k] = t
        else:
            continue
        copied.append(k)
    return new, copied
