"""PGUNet1..4, ProgressiveUNet and UncertaintyGuidedLoss on the sm_100a kernels.

Interface mirror of the reference's `UG_unet.py`: same class names, constructor
arguments, sub-module attribute names (`inc`, `down1..4`, `up1..4`,
`outc | outc1..4`, `stage1..4`) and therefore the same `state_dict` keys, so
reference checkpoints load here and vice versa.  `forward` takes an fp32 NCHW
CUDA tensor and returns raw fp32 NCHW logits (no sigmoid / softmax), like the
reference (UG_unet.py:188-193, 209-223, 242-260, 282-304).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine, ops
from .unet_parts import DoubleConv, Down, DownSample, InConv, OutConv, Up, UpSample, _require_cuda


class _PGUNetBase(nn.Module):
    """Shared forward for the four stage models: encoder (`inc`, downs), decoder (ups), one OutConv per
    decoder scale, coarse logits bilinearly upsampled (align_corners=True) and summed."""

    _downs: tuple = ()
    _ups: tuple = ()
    _heads: tuple = ()
    # Normalisation folded into the first-layer packer for uint8 inputs: (byte/255 - mean) / std per
    # channel (ToTensor + optional Normalize of the reference's loaders); None = ToTensor only.
    input_mean = None
    input_std = None

    def set_input_norm(self, mean=None, std=None):
        """Per-channel Normalize(mean, std) applied to uint8 image inputs (python sequences or None)."""
        self.input_mean = None if mean is None else tuple(float(v) for v in mean)
        self.input_std = None if std is None else tuple(float(v) for v in std)
        return self

    def pack_bytes(self, x: torch.Tensor, resolution: Optional[int] = None) -> torch.Tensor:
        """uint8 HWC image bytes [B,H,W,C] -> first-layer rows at `resolution` (bilinear,
        align_corners=True, like uncertainty_guided_trainer.py:208) with /255 + Normalize folded in."""
        if x.dim() != 4 or x.shape[-1] != self.inc.conv.conv_op[0].in_channels:
            raise ValueError(f"{type(self).__name__}: uint8 input must be [B, H, W, "
                             f"{self.inc.conv.conv_op[0].in_channels}] image bytes, got {tuple(x.shape)}")
        return ops.im2col_first_u8(x.contiguous(), resolution, resolution, self.input_mean, self.input_std)

    def first_rows(self, x: torch.Tensor, resolution: Optional[int] = None):
        """The first convolution's source for an fp32 NCHW image or uint8 HWC loader bytes at `resolution`:
        an `ops.FirstRows` handle (rows built inside the fused first-layer kernels) where those cover the
        shape, else the packed rows of `pack_bytes` / `ops.im2col_first`."""
        c1 = self.inc.conv.conv_op[0]
        if x.dtype == torch.uint8:
            if x.dim() != 4 or x.shape[-1] != c1.in_channels:
                return self.pack_bytes(x, resolution)          # raises the shape error
            return ops.first_rows(x.contiguous(), resolution, resolution, self.input_mean, self.input_std,
                                  cout=c1.out_channels)
        if x.dtype != torch.float32:
            x = x.float()
        return ops.first_rows(x.contiguous(), resolution, resolution, cout=c1.out_channels)

    def encode(self, x: torch.Tensor, depth: Optional[int] = None, resolution: Optional[int] = None):
        """Runs `inc` and the first `depth` Down blocks (all by default) on an fp32 NCHW image - or on
        uint8 HWC image bytes, see `pack_bytes` - and returns the list of NHWC bf16 feature maps.  The
        max-pool feeding each Down is fused into the producer's BN/ReLU pass."""
        _require_cuda(x, type(self).__name__)
        downs = [getattr(self, n) for n in self._downs]
        if depth is not None:
            downs = downs[:depth]
        if x.dtype == torch.uint8:
            kind, x = "col", self.first_rows(x, resolution)
        else:
            kind = "first"
            if x.dtype != torch.float32:
                x = x.float()
            if resolution is not None and tuple(x.shape[-2:]) != (resolution, resolution):
                kind, x = "col", self.first_rows(x, resolution)
            else:
                x = x.contiguous()
        feats = []
        if downs:
            a, p = self.inc.conv._run(kind, x, emit_pool=True)
        else:
            a, p = self.inc.conv._run(kind, x), None
        feats.append(a)
        for i, d in enumerate(downs):
            dc = d.mpconv[1]
            if i + 1 < len(downs):
                a, p = dc._run("plain", p, emit_pool=True)
            else:
                a, p = dc._run("plain", p), None
            feats.append(a)
        return feats

    def forward(self, x, resolution: Optional[int] = None):
        """fp32 NCHW image (the reference's call) or uint8 HWC bytes -> raw fp32 NCHW logits.
        `resolution` (optional) resizes the input inside the first-layer packer."""
        feats = self.encode(x, resolution=resolution)
        cur = feats[-1]
        n = len(feats) - 1
        heads = [getattr(self, h) for h in self._heads]
        # head j reads decoder output j (a single head reads the last one); in train mode it is computed
        # inside the block's second BN/ReLU pass (forward) and BatchNorm backward (engine.head_is_fusable)
        head_of = {len(self._ups) - 1: heads[0]} if len(heads) == 1 else dict(enumerate(heads))
        maps = []
        for j, name in enumerate(self._ups):
            dc = getattr(self, name).conv
            h = head_of.get(j)
            if h is not None and engine.head_is_fusable(dc, h.conv):
                cur, lmap = dc._run("up", cur, feats[n - 1 - j], head=h.conv)
                maps.append(lmap)
            else:
                cur = dc._run("up", cur, feats[n - 1 - j])
                if h is not None:
                    maps.append(h._run(cur))
        if len(maps) == 1:
            return maps[0]
        return engine.HeadCombineFn.apply(*maps)


class PGUNet1(_PGUNetBase):
    """Stage 1, 32x32 — reference UG_unet.py:178-193."""
    _downs, _ups, _heads = ("down4",), ("up1",), ("outc",)

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.inc = InConv(in_channels, 512)
        self.down4 = Down(512, 512)
        self.up1 = Up(1024, 256)
        self.outc = OutConv(256, num_classes)


class PGUNet2(_PGUNetBase):
    """Stage 2, 64x64 — reference UG_unet.py:196-223."""
    _downs, _ups, _heads = ("down3", "down4"), ("up1", "up2"), ("outc1", "outc2")

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.inc = InConv(in_channels, 256)
        self.down3 = Down(256, 512)
        self.down4 = Down(512, 512)
        self.up1 = Up(1024, 256)
        self.up2 = Up(512, 128)
        self.outc1 = OutConv(256, num_classes)
        self.outc2 = OutConv(128, num_classes)


class PGUNet3(_PGUNetBase):
    """Stage 3, 128x128 — reference UG_unet.py:226-260."""
    _downs, _ups, _heads = ("down2", "down3", "down4"), ("up1", "up2", "up3"), ("outc1", "outc2", "outc3")

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.inc = InConv(in_channels, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        self.down4 = Down(512, 512)
        self.up1 = Up(1024, 256)
        self.up2 = Up(512, 128)
        self.up3 = Up(256, 64)
        self.outc1 = OutConv(256, num_classes)
        self.outc2 = OutConv(128, num_classes)
        self.outc3 = OutConv(64, num_classes)


class PGUNet4(_PGUNetBase):
    """Stage 4, 256x256 — reference UG_unet.py:263-304."""
    _downs = ("down1", "down2", "down3", "down4")
    _ups = ("up1", "up2", "up3", "up4")
    _heads = ("outc1", "outc2", "outc3", "outc4")

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.inc = InConv(in_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        self.down4 = Down(512, 512)
        self.up1 = Up(1024, 256)
        self.up2 = Up(512, 128)
        self.up3 = Up(256, 64)
        self.up4 = Up(128, 64)
        self.outc1 = OutConv(256, num_classes)
        self.outc2 = OutConv(128, num_classes)
        self.outc3 = OutConv(64, num_classes)
        self.outc4 = OutConv(64, num_classes)


STAGE_MODELS = {1: PGUNet1, 2: PGUNet2, 3: PGUNet3, 4: PGUNet4}


def transfer_state(prev_stage_dict, current_stage_dict, verbose_stage=None):
    """Same-key copy, leading-slice partial copy on shape mismatch — reference UG_unet.py:345-411.
    Works on the state dicts where they live (device-side copies, no throw-away model)."""
    new_state = {k: v.clone() for k, v in current_stage_dict.items()}
    copied = []
    for k, pv in prev_stage_dict.items():
        cv = current_stage_dict.get(k)
        if cv is None or not isinstance(pv, torch.Tensor) or not isinstance(cv, torch.Tensor):
            continue
        pv = pv.to(cv.device)
        if pv.shape == cv.shape:
            new_state[k] = pv.clone()
        elif pv.ndim == cv.ndim and pv.ndim in (1, 2, 4):
            lead = 2 if pv.ndim > 1 else 1
            sl = tuple(slice(0, min(a, b)) for a, b in zip(pv.shape[:lead], cv.shape[:lead]))
            tmp = cv.clone()
            tmp[sl] = pv[sl]
            new_state[k] = tmp
        else:
            continue
        copied.append(k)
    if verbose_stage is not None:
        print(f"transfer_weights(stage={verbose_stage}): copied {len(copied)} keys (examples: {copied[:5]})")
    return new_state


class ProgressiveUNet(nn.Module):
    """All four stages behind one module with dynamic stage switching — reference UG_unet.py:307-426.

    Accepts the code's signature `(in_channels, num_classes)` and the README's
    `(in_channels, out_channels, stage, task_type)` (reference README.md:46-49)."""

    def __init__(self, in_channels, num_classes=None, stage=1, task_type="segmentation", out_channels=None):
        super().__init__()
        if num_classes is None:
            num_classes = out_channels
        if num_classes is None:
            raise TypeError("ProgressiveUNet: num_classes (or out_channels) is required")
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.task_type = task_type
        self.stage_resolutions = {1: 32, 2: 64, 3: 128, 4: 256}
        self.stage1 = PGUNet1(in_channels, num_classes)
        self.stage2 = PGUNet2(in_channels, num_classes)
        self.stage3 = PGUNet3(in_channels, num_classes)
        self.stage4 = PGUNet4(in_channels, num_classes)
        self.stages = {1: self.stage1, 2: self.stage2, 3: self.stage3, 4: self.stage4}
        self.current_stage = 1
        self.set_stage(stage)

    def set_stage(self, stage):
        if stage not in [1, 2, 3, 4]:
            raise ValueError("Stage must be 1, 2, 3, or 4")
        self.current_stage = stage

    def get_current_resolution(self):
        return self.stage_resolutions[self.current_stage]

    def transfer_weights(self, prev_stage_dict, current_stage_dict, stage):
        return transfer_state(prev_stage_dict, current_stage_dict, verbose_stage=stage)

    def forward(self, x, target_resolution=None):
        if target_resolution is None:
            target_resolution = self.get_current_resolution()
        if x.shape[-2:] != (target_resolution, target_resolution):
            x = F.interpolate(x, size=(target_resolution, target_resolution), mode="bilinear", align_corners=True)
        return self.stages[self.current_stage](x)


class UncertaintyGuidedLoss(nn.Module):
    """Uncertainty-guided loss weighting — reference UG_unet.py:8-94.

    Same two methods / signatures / return types.  `generate_uncertainty_map` fuses the input
    down-resize into the previous model's first-layer packer and sigmoid + up-resize + margin into one
    kernel; `apply_uncertainty_weighted_loss` fuses BCE(pos_weight) x (1 + alpha*A) -> mean, the base
    loss and dlogits into a single pass when `loss_fn` is a `BCEWithLogitsLoss(reduction='none')`
    (any other callable is honoured through the generic path)."""

    def __init__(self, device="cuda", task_type="segmentation"):
        super().__init__()
        self.device = device
        self.task_type = task_type
        self.last_metrics = None     # float64 device tensor from the fused loss kernel

    @torch.no_grad()
    def generate_uncertainty_map(self, input_current, model_prev, prev_resolution, current_resolution):
        model_prev.eval()
        x = input_current
        _require_cuda(x, "UncertaintyGuidedLoss")
        if isinstance(model_prev, _PGUNetBase):
            logits_prev = _forward_resized(model_prev, x, prev_resolution)
        else:   # foreign module: honour the reference's data flow with stock ops
            xp = F.interpolate(x, size=(prev_resolution, prev_resolution), mode="bilinear", align_corners=True)
            logits_prev = model_prev(xp).float().contiguous()
        return ops.uncertainty_margin(logits_prev.contiguous(), current_resolution, current_resolution)

    @staticmethod
    def fused_criterion(loss_fn, output_current, uncertainty_map) -> bool:
        """True when `loss_fn` is the trainer's BCEWithLogitsLoss(pos_weight scalar, 'none') and the
        single-pass kernel applies."""
        return (isinstance(loss_fn, nn.BCEWithLogitsLoss) and loss_fn.reduction == "none"
                and loss_fn.weight is None and output_current.is_cuda
                and (loss_fn.pos_weight is None or loss_fn.pos_weight.numel() == 1)
                and (uncertainty_map is None or uncertainty_map.shape == output_current.shape))

    def weighted_loss_device(self, loss_fn, output_current, target_current, uncertainty_map=None, alpha=1.0):
        """Fused path without any host synchronisation: returns (final_loss, metrics) where metrics is
        the float64 device tensor documented at ugp_bce_uncertainty_loss (also kept in
        `self.last_metrics`)."""
        if target_current.shape != output_current.shape:
            raise ValueError(f"Target size ({target_current.size()}) must be the same as input size "
                             f"({output_current.size()})")
        pw = 1.0 if loss_fn.pos_weight is None else _pos_weight_value(loss_fn)
        if target_current.dtype != torch.uint8:      # uint8 masks feed the kernel as bytes
            target_current = target_current.float()
        final_loss, out = engine.BCEUncertaintyLossFn.apply(
            output_current, target_current, uncertainty_map, float(alpha), pw)
        self.last_metrics = out
        return final_loss, out

    def apply_uncertainty_weighted_loss(self, loss_fn, output_current, target_current, uncertainty_map=None,
                                        alpha=1.0):
        """reference UG_unet.py:61-94 — returns (final_loss tensor, base_loss python float)."""
        if not self.fused_criterion(loss_fn, output_current, uncertainty_map):
            pixel_loss = loss_fn(output_current, target_current)
            if uncertainty_map is None:
                final_loss = torch.mean(pixel_loss)
            else:
                final_loss = torch.mean(pixel_loss * (1.0 + alpha * uncertainty_map).detach())
            return final_loss, torch.mean(pixel_loss).item()
        final_loss, out = self.weighted_loss_device(loss_fn, output_current, target_current, uncertainty_map, alpha)
        return final_loss, (out[1] / output_current.numel()).item()


_PW_CACHE = {}


def _pos_weight_value(loss_fn) -> float:
    """pos_weight as a python float, read back once per (tensor, version)."""
    pwt = loss_fn.pos_weight
    key = (pwt.data_ptr(), pwt._version)
    v = _PW_CACHE.get(id(loss_fn))
    if v is None or v[0] != key:
        v = (key, float(pwt.reshape(-1)[0].item()))
        _PW_CACHE[id(loss_fn)] = v
    return v[1]


def _forward_resized(model: _PGUNetBase, x: torch.Tensor, resolution: int) -> torch.Tensor:
    """model(F.interpolate(x, resolution, bilinear, align_corners=True)) with the resize folded into the
    first-layer packer (UG_unet.py:36-43)."""
    if x.dtype == torch.uint8:
        return _forward_from_col(model, model.first_rows(x, resolution))
    if x.shape[-1] == resolution and x.shape[-2] == resolution:
        return model(x)
    return _forward_from_col(model, model.first_rows(x, resolution))


def _forward_from_col(model: _PGUNetBase, col) -> torch.Tensor:
    """Inference forward starting from the first layer's rows - packed, or an ops.FirstRows handle (no autograd)."""
    dc = model.inc.conv
    c1, b1, c2, b2 = dc.conv_op[0], dc.conv_op[1], dc.conv_op[3], dc.conv_op[4]
    downs = [getattr(model, n).mpconv[1] for n in model._downs]
    a, _, _ = engine.cbr_forward([col], 1, c1, b1, dc._packs[0], False, False)
    a, p, _ = engine.cbr_forward([a], 9, c2, b2, dc._packs[1], bool(downs), False)
    feats = [a]
    for i, d in enumerate(downs):
        a, _, _ = engine.cbr_forward([p], 9, d.conv_op[0], d.conv_op[1], d._packs[0], False, False)
        a, p, _ = engine.cbr_forward([a], 9, d.conv_op[3], d.conv_op[4], d._packs[1], i + 1 < len(downs), False)
        feats.append(a)
    cur = feats[-1]
    dec = []
    n = len(feats) - 1
    for j, name in enumerate(model._ups):
        u = getattr(model, name).conv
        up = ops.upsample2x(cur)
        a, _, _ = engine.cbr_forward([feats[n - 1 - j], up], 9, u.conv_op[0], u.conv_op[1], u._packs[0], False, False)
        cur, _, _ = engine.cbr_forward([a], 9, u.conv_op[3], u.conv_op[4], u._packs[1], False, False)
        dec.append(cur)
    heads = [getattr(model, h) for h in model._heads]
    if len(heads) == 1:
        h = heads[0].conv
        return ops.outconv_fwd(dec[-1], h.weight.detach().reshape(h.out_channels, -1).contiguous(), h.bias.detach())
    maps = [ops.outconv_fwd(d, h.conv.weight.detach().reshape(h.conv.out_channels, -1).contiguous(),
                            h.conv.bias.detach()) for h, d in zip(heads, dec)]
    return ops.head_combine_fwd(maps)


class UNet(nn.Module):
    """The reference's legacy U-Net (UG_unet.py:430-462): DownSample x4, a 512->1024 bottleneck, four
    UpSample blocks (ConvTranspose2d(k=2, s=2) + concat + DoubleConv) and a 1x1 head; same attribute
    names and `state_dict` keys.  fp32 NCHW image in, raw fp32 NCHW logits out; H and W must be
    multiples of 256 (the 16x-downsampled bottleneck runs 16-pixel-wide GEMM tiles)."""

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.down_convolution_1 = DownSample(in_channels, 64)
        self.down_convolution_2 = DownSample(64, 128)
        self.down_convolution_3 = DownSample(128, 256)
        self.down_convolution_4 = DownSample(256, 512)
        self.bottle_neck = DoubleConv(512, 1024)
        self.up_convolution_1 = UpSample(1024, 512)
        self.up_convolution_2 = UpSample(512, 256)
        self.up_convolution_3 = UpSample(256, 128)
        self.up_convolution_4 = UpSample(128, 64)
        self.out = nn.Conv2d(in_channels=64, out_channels=num_classes, kernel_size=1)
        if not 1 <= num_classes <= 8:
            raise ValueError(f"UNet: num_classes={num_classes} unsupported (1..8)")

    def forward(self, x):
        _require_cuda(x, "UNet")
        x = x.float().contiguous()
        d1, p1 = self.down_convolution_1._run(x)
        d2, p2 = self.down_convolution_2._run(p1)
        d3, p3 = self.down_convolution_3._run(p2)
        d4, p4 = self.down_convolution_4._run(p3)
        b = self.bottle_neck._run("plain", p4)
        u1 = self.up_convolution_1._run(b, d4)
        u2 = self.up_convolution_2._run(u1, d3)
        u3 = self.up_convolution_3._run(u2, d2)
        u4 = self.up_convolution_4._run(u3, d1)
        return engine.OutConvFn.apply(u4, self.out.weight, self.out.bias)
