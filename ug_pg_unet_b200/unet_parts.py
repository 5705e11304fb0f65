"""Drop-in U-Net building blocks backed by the sm_100a kernels.

Same constructors, attribute names and `state_dict` keys as the reference's
`UG_unet_parts.py` (DoubleConv :5-19, InConv :21-28, DownSample :31-41, Down :44-54,
UpSample :57-67, Up :70-81, OutConv :84-91): the stock `nn.Conv2d` / `nn.BatchNorm2d` objects are kept as
parameter containers (initialisation, `.to()`, `load_state_dict`, checkpoints
all behave as in the reference), while `forward` runs the tcgen05 implicit-GEMM
convolution and the fused BN/ReLU/pool/upsample kernels.

Tensor formats at a block boundary
  * fp32 (or any float) logical NCHW in  -> fp32 NCHW out   ("compat", default)
  * bf16 logical NCHW with channels_last strides in -> same format out ("fast",
    zero-copy between blocks; what PGUNet1..4 use internally)
CUDA only: a CPU tensor raises (there is no fallback path by design).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine
from .engine import PackedWeights, is_fast_act, nhwc_as_nchw, run_double_conv, to_nhwc_bf16


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: ug_pg_unet_b200 runs on CUDA (sm_100a) only; got a {x.device} tensor. "
                           "There is deliberately no CPU fallback.")


def _in(x: torch.Tensor):
    """-> (NHWC bf16 tensor, was_fast)."""
    if is_fast_act(x):
        return x.permute(0, 2, 3, 1), True
    return to_nhwc_bf16(x), False


def _out(a: torch.Tensor, fast: bool, like: torch.Tensor):
    y = nhwc_as_nchw(a)
    if fast:
        return y
    return y.to(like.dtype if like.is_floating_point() else torch.float32).contiguous()


class DoubleConv(nn.Module):
    """(conv => BN => ReLU) * 2 — reference UG_unet_parts.py:5-19."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv_op = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )
        first = in_channels % 64 != 0
        if first and 9 * in_channels > 64:
            raise ValueError(f"DoubleConv: in_channels={in_channels} unsupported "
                             "(must be a multiple of 64, or <= 7 for the image layer)")
        if out_channels % 64 != 0 or out_channels > 1024:
            raise ValueError(f"DoubleConv: out_channels={out_channels} unsupported (multiple of 64, <= 1024)")
        self._first = first
        # private bf16 operand caches (not parameters, not in state_dict)
        self._packs = (PackedWeights(self.conv_op[0], need_dgrad=not first),
                       PackedWeights(self.conv_op[3], need_dgrad=True))

    # -- internal entry points used by Down / Up / PGUNet (NHWC bf16 in/out) --
    def _run(self, kind, x, skip=None, emit_pool=False, head=None):
        return run_double_conv(self, kind, x, skip, emit_pool, head=head)

    def forward(self, x):
        _require_cuda(x, "DoubleConv")
        if self._first:
            if x.dtype != torch.float32:
                x = x.float()
            a = self._run("first", x.contiguous())
            return _out(a, False, x)
        xin, fast = _in(x)
        return _out(self._run("plain", xin), fast, x)


class InConv(nn.Module):
    """Initial convolution block — reference UG_unet_parts.py:21-28."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x):
        return self.conv(x)


class DownSample(nn.Module):
    """DoubleConv then MaxPool2d(2); returns (features, pooled) — reference UG_unet_parts.py:31-41
    (legacy UNet encoder block).  The pool is fused into the second BN/ReLU pass."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = DoubleConv(in_channels, out_channels)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    def _run(self, x):
        """fp32 NCHW image (first block) or NHWC bf16 -> (features, pooled), both NHWC bf16."""
        return self.conv._run("first" if self.conv._first else "plain", x, emit_pool=True)

    def forward(self, x):
        _require_cuda(x, "DownSample")
        if self.conv._first:
            x = x.float().contiguous()
            a, p = self._run(x)
            return _out(a, False, x), _out(p, False, x)
        xin, fast = _in(x)
        a, p = self._run(xin)
        return _out(a, fast, x), _out(p, fast, x)


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv — reference UG_unet_parts.py:44-54."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        _require_cuda(x, "Down")
        xin, fast = _in(x)
        return _out(self.mpconv[1]._run("pool", xin), fast, x)


class UpSample(nn.Module):
    """ConvTranspose2d(in, in//2, 2, 2) of x1, cat([x1, x2]), DoubleConv — reference
    UG_unet_parts.py:57-67 (legacy UNet decoder block).  The transposed convolution is a tcgen05 GEMM
    whose epilogue writes the pixel-shuffled result; the concat is walked as two sources."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)
        if in_channels % 128 != 0 or in_channels > 1024:
            raise ValueError(f"UpSample: in_channels={in_channels} unsupported (multiple of 128, <= 1024)")
        self._tpack = engine.TConvPackedWeights(self.up)

    def _run(self, low, skip):
        """NHWC bf16 (low-res, skip) -> NHWC bf16."""
        up = engine.ConvTranspose2x2Fn.apply(low, self.up.weight, self.up.bias, self._tpack)
        return self.conv._run("cat", up, skip)

    def forward(self, x1, x2):
        _require_cuda(x1, "UpSample")
        low, fast1 = _in(x1)
        skip, fast2 = _in(x2)
        if low.shape[3] // 2 + skip.shape[3] != self.conv.conv_op[0].in_channels:
            raise RuntimeError(f"UpSample: expected {self.conv.conv_op[0].in_channels} concatenated channels, got "
                               f"{low.shape[3] // 2} + {skip.shape[3]}")
        return _out(self._run(low, skip), fast1 and fast2, x1)


class Up(nn.Module):
    """Bilinear x2 (align_corners=True) of x1, cat([x2, x1]), DoubleConv — reference
    UG_unet_parts.py:70-81.  `in_channels` is the concatenated width; the `bilinear` argument is
    accepted and ignored exactly as in the reference."""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        _require_cuda(x1, "Up")
        low, fast1 = _in(x1)
        skip, fast2 = _in(x2)
        if low.shape[3] + skip.shape[3] != self.conv.conv_op[0].in_channels:
            raise RuntimeError(f"Up: expected {self.conv.conv_op[0].in_channels} concatenated channels, got "
                               f"{skip.shape[3]} + {low.shape[3]}")
        return _out(self.conv._run("up", low, skip), fast1 and fast2, x1)


class OutConv(nn.Module):
    """Final 1x1 convolution — reference UG_unet_parts.py:84-91.  Always returns fp32 NCHW logits."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        if in_channels not in (64, 128, 256) or not 1 <= out_channels <= 8:
            raise ValueError(f"OutConv: ({in_channels} -> {out_channels}) unsupported "
                             "(in_channels in 64/128/256, 1..8 classes)")

    def _run(self, x_nhwc):
        return engine.OutConvFn.apply(x_nhwc, self.conv.weight, self.conv.bias)

    def forward(self, x):
        _require_cuda(x, "OutConv")
        xin, _ = _in(x)
        return self._run(xin)
