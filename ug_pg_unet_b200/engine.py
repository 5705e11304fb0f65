"""Autograd glue between the nn.Module surface and the C-ABI kernels.

The unit of work is Conv3x3(+bias) -> BatchNorm2d -> ReLU ("CBR"), the building
block of the reference's DoubleConv (UG_unet_parts.py:5-19).  Forward keeps the
raw conv output z (bf16) and the per-channel BN coefficients; backward
recomputes the activation from z, so the ReLU mask / max-pool arg-max never
need to be stored.

Activations travel as NHWC bf16 tensors [B, H, W, C].
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import set_option

bf16 = torch.bfloat16

# What-if timing switches: a step is timed with a piece of work REMOVED to bound what
# fusing it away could gain.  Results are wrong while one is set; both are off unless the variable is set.
#   UGP_WHATIF_SKIP_BN1_REDUCE: no reduction pass in the BatchNorm backward of each DoubleConv's first BN
#   UGP_WHATIF_POOL_AS_PLAIN:   pooled gradient dropped, so the pooled BatchNorm backward runs the plain kernels
_WHATIF_SKIP_BN1_REDUCE = os.environ.get("UGP_WHATIF_SKIP_BN1_REDUCE", "0") == "1"
_WHATIF_POOL_AS_PLAIN = os.environ.get("UGP_WHATIF_POOL_AS_PLAIN", "0") == "1"
# UGP_FUSED_BN_BWD_REDUCE=1: the reduction pass of every DoubleConv's first BatchNorm backward runs inside the
# epilogue of the dgrad that produces its input (ugp_conv_dgrad_bnstats_bf16) instead of as its own kernel.
# Built, parity-tested and measured: the kernels of a step take 0.10 ms less, but the step
# does not get faster - the stand-alone reduction pass is HBM-bound and hides under the weight-gradient
# convolutions of the side stream, while the longer, epilogue-bound dgrad (64 -> 64 @256^2: 77 -> 116 us)
# competes with them for the tensor pipe.  Off by default.
_HEAD_ADJOINT_ON_SIDE_STREAM = os.environ.get("UGP_HEAD_ADJOINT_MAIN", "0") != "1"      # =1: on the main stream (A/B)
_MATERIALIZE_GRADS = os.environ.get("UGP_MATERIALIZE_GRADS", "0") == "1"     # autograd's default (A/B only)
_FUSED_BN_BWD_REDUCE = os.environ.get("UGP_FUSED_BN_BWD_REDUCE", "0") == "1"


# NVTX ranges (UGP_NVTX=1): one range per DoubleConv block and direction, one per trainer-step phase, so a
# timeline (nsys, or ncu --nvtx) of the ~190-kernel step reads as blocks.  Off by default: the pushes / pops
# are host calls on a host-bound eager step (they do not exist inside a replayed CUDA graph).
NVTX = os.environ.get("UGP_NVTX", "0") == "1"


class nvtx_range:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


class PackedWeights:
    """bf16 GEMM operands of one nn.Conv2d weight, refreshed when the fp32 master changes
    (optimizer.step / load_state_dict / .to(device) all bump `_version` or the storage), plus the
    persistent per-layer accumulators (BN statistics, packed fp32 weight gradient).

    Deferred weight gradients: with `defer = True` (set by `enable_deferred_wgrad`, used by the
    trainer together with `FusedRMSprop`) backward leaves the weight gradient in the PACKED layout
    the tensor-core kernel reduces into and returns no `.grad` for the conv weight; the optimiser
    kernel consumes the packed gradient, rewrites these bf16 packs and zeroes the accumulator.  The
    default (`defer = False`) materialises the OIHW `.grad` like the reference's autograd."""

    def __init__(self, conv: nn.Conv2d, need_dgrad: bool):
        self.conv = conv
        self.need_dgrad = need_dgrad
        self.key = None
        self.wf = None
        self.wd = None
        self._stats = None   # persistent fp32 [SMs, 2, C] per-CTA partial BN statistics (bn_finalize re-zeroes it)
        self._dwp = None     # persistent packed fp32 weight-gradient accumulator (its consumer re-zeroes it)
        self._eval_coef = None   # cached eval-mode BN scale / shift of the BatchNorm behind this conv
        self._eval_key = None
        self.defer = False
        self.grad_pending = False      # a backward accumulated into _dwp since the last optimiser step
        self.on_grad_ready = None      # data-parallel reducer callback, called after the wgrad launch

    def stats_buffer(self, cout: int, device) -> torch.Tensor:
        if self._stats is None or self._stats.device != device or self._stats.shape[2] != cout:
            self._stats = ops.new_stats(cout, device)
        return self._stats

    def dwp_buffer(self, numel: int, device) -> torch.Tensor:
        if self._dwp is None or self._dwp.device != device or self._dwp.numel() != numel:
            self._dwp = torch.zeros((numel,), dtype=torch.float32, device=device)
        return self._dwp

    def eval_coef(self, bn: nn.BatchNorm2d, device) -> torch.Tensor:
        ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in ts) + (float(bn.eps), PackedWeights.epoch)
        if self._eval_coef is None or self._eval_key != key or self._eval_coef.device != device:
            coef = torch.empty((2, bn.num_features), dtype=torch.float32, device=device) \
                if self._eval_coef is None or self._eval_coef.device != device else self._eval_coef
            ops.bn_eval_coeffs(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                               float(bn.eps), coef[0], coef[1])
            self._eval_coef, self._eval_key = coef, key
        return self._eval_coef

    def assign_dwp(self, buf: Optional[torch.Tensor]) -> None:
        """Use `buf` (a zeroed slice of a flat gradient arena) as the packed-gradient accumulator."""
        self._dwp = buf

    def can_defer(self) -> bool:
        w = self.conv.weight
        return w.is_cuda and w.shape[1] % 64 == 0 and w.shape[0] % 32 == 0 and tuple(w.shape[2:]) == (3, 3)

    # bumped by whoever changes weights behind autograd's back (a CUDA-graph replay that contains
    # the optimizer step, or FusedRMSprop's kernels, do not advance `_version`)
    epoch = 0

    @classmethod
    def invalidate_all(cls):
        cls.epoch += 1

    def _current_key(self):
        w = self.conv.weight
        return (w.data_ptr(), w._version, w.device, PackedWeights.epoch)

    def invalidate(self) -> None:
        self.key = None

    def mark_fresh(self) -> None:
        """The packs were just rewritten from the current weights (by the optimiser kernel)."""
        if self.wf is not None:
            self.key = self._current_key()

    def get(self):
        key = self._current_key()
        if key != self.key:
            w = self.conv.weight
            same = self.wf is not None and self.wf.device == w.device and self.wf.numel() == \
                (w.shape[0] * 64 if w.shape[1] % 64 else w.numel())
            # refresh IN PLACE when possible: captured CUDA graphs hold these addresses
            self.wf, self.wd = ops.pack_weights(w.detach(), self.need_dgrad, wf=self.wf if same else None,
                                                wd=self.wd if same else None)
            self.key = key
        return self.wf, self.wd


def conv_packs(model: nn.Module) -> List[PackedWeights]:
    """All PackedWeights of `model` in forward (registration) order."""
    out: List[PackedWeights] = []
    for m in model.modules():
        pk = getattr(m, "_packs", None)
        if pk is not None:
            out.extend(pk)
    return out


def refresh_eval_caches(model: nn.Module) -> None:
    """Brings the private inference caches of every DoubleConv in `model` up to date IN PLACE (bf16
    operand packs, folded eval-mode BatchNorm scale / shift): a captured CUDA graph of the inference pass
    holds their addresses and does not re-run this host logic on replay."""
    for m in model.modules():
        pk = getattr(m, "_packs", None)
        ops_ = getattr(m, "conv_op", None)
        if pk is None or ops_ is None:
            continue
        for pack, bn in zip(pk, (ops_[1], ops_[4])):
            pack.get()
            pack.eval_coef(bn, pack.conv.weight.device)


def enable_deferred_wgrad(model: nn.Module, enable: bool = True):
    """Switches every eligible 3x3 conv of `model` to deferred (packed) weight gradients and lays
    their accumulators out in ONE flat fp32 arena, in reverse forward order = the order backward
    fills them, so a data-parallel bucket is a contiguous arena range.

    Returns (packs in arena order, arena tensor) — ([], None) when disabled / nothing eligible."""
    packs = [pk for pk in conv_packs(model)]
    if not enable:
        for pk in packs:
            if pk.defer:
                pk.defer, pk.grad_pending, pk.on_grad_ready = False, False, None
                pk.assign_dwp(None)
        return [], None
    elig = [pk for pk in reversed(packs) if pk.can_defer()]
    if not elig:
        return [], None
    dev = elig[0].conv.weight.device
    total = sum(pk.conv.weight.numel() for pk in elig)
    arena = torch.zeros((total,), dtype=torch.float32, device=dev)
    off = 0
    for pk in elig:
        n = pk.conv.weight.numel()
        pk.assign_dwp(arena[off:off + n])
        pk.defer, pk.grad_pending = True, False
        off += n
    return elig, arena


class WgradOverlap:
    """Optional second stream for the weight-gradient convolutions (set by the trainer around
    `backward()`): wgrad(l) only feeds the optimiser, so it can run beside the dgrad(l) -> BN-backward(l-1)
    chain on the main stream (tensor-bound next to HBM-bound work).  Tensors the side stream reads are
    kept alive in `keep` until the trainer has joined the streams."""
    stream: Optional[torch.cuda.Stream] = None
    keep: list = []
    # the pack whose weight gradient is produced LAST (the first deferred layer of the network) and what to
    # do once it and its input gradient are queued: the trainer starts the packed-weight optimiser there, on
    # the side stream, beside the tail of the backward pass (trainer._early_optimizer_step)
    last_pack = None
    on_last = None
    # optional second side stream: consecutive weight gradients alternate between the two, so that the
    # split-K reduction tail of one overlaps the main loop of the next (UGP_WGRAD_STREAMS=2)
    stream2: Optional[torch.cuda.Stream] = None
    _turn = 0
    head_event = None       # recorded after the coarse head-map adjoints launched on the side stream (HeadCombineFn)
    head_direct_ptr = 0     # data_ptr of dlogits, the finest map's gradient (never on the side stream)

    @classmethod
    def begin(cls, stream, last_pack=None, on_last=None, stream2=None):
        cls.stream, cls.keep = stream, []
        cls.last_pack, cls.on_last = last_pack, on_last
        cls.stream2, cls._turn, cls.head_event = stream2, 0, None

    @classmethod
    def pick(cls):
        """The side stream of the next weight gradient."""
        if cls.stream is None or cls.stream2 is None:
            return cls.stream
        cls._turn ^= 1
        return cls.stream if cls._turn else cls.stream2

    @classmethod
    def streams(cls):
        return [s for s in (cls.stream, cls.stream2) if s is not None]

    @classmethod
    def end(cls):
        """Joins the side stream(s) into the current one and releases the kept tensors."""
        sts = cls.streams()
        cls.stream = cls.stream2 = None
        cls.last_pack = cls.on_last = cls.head_event = None
        for st in sts:
            torch.cuda.current_stream().wait_stream(st)
        cls.keep = []


class CBRSaved:
    __slots__ = ("srcs", "z", "coef", "taps", "pooled")

    def __init__(self, srcs, z, coef, taps, pooled):
        self.srcs = srcs      # list of 1-2 NHWC bf16 inputs of the conv
        self.z = z            # raw conv output (+bias), bf16
        self.coef = coef      # fp32 [4, C]: scale, shift, mean, rstd
        self.taps = taps
        self.pooled = pooled  # whether a fused 2x2 max-pool output was emitted


def cbr_forward(srcs: Sequence[torch.Tensor], taps: int, conv: nn.Conv2d, bn: nn.BatchNorm2d,
                packs: PackedWeights, emit_pool: bool, save: bool, head=None):
    """One Conv -> BN -> ReLU unit.  Returns (a, pooled or None, CBRSaved or None).
    `head` = (w [nc, C], b [nc]) fp32: train mode only, the 1x1 head on the activation is computed in the
    BN/ReLU pass and returned in place of the pooled tensor."""
    src0 = srcs[0]
    src1 = srcs[1] if len(srcs) > 1 else None
    B, H, W, _ = src0.shape
    cout = conv.out_channels
    wf, _ = packs.get()
    dev = src0.device
    bias = conv.bias.detach() if conv.bias is not None else None
    if bn.training:
        stats = packs.stats_buffer(cout, dev)
        z = ops.conv_gemm(src0, src1, wf, taps, cout, bias=bias, stats=stats)
        coef = torch.empty((4, cout), dtype=torch.float32, device=dev)
        momentum = 0.0 if bn.momentum is None else float(bn.momentum)
        track = bn.track_running_stats and bn.running_mean is not None
        if track and bn.momentum is None:   # cumulative moving average
            momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
        packs._eval_key = None      # bn_finalize updates the running statistics behind autograd's back
        ops.bn_finalize(stats, B * H * W, bn.weight.detach(), bn.bias.detach(), float(bn.eps), momentum,
                        bn.running_mean if track else None, bn.running_var if track else None,
                        bn.num_batches_tracked if track else None, coef[0], coef[1], coef[2], coef[3])
        if head is not None:
            a, p = ops.bn_relu_apply_head(z, coef[0], coef[1], head[0], head[1])
        elif emit_pool:
            a, p = ops.bn_relu_apply(z, coef[0], coef[1], pool=True)
        else:
            a, p = ops.bn_relu_apply(z, coef[0], coef[1]), None
        saved = CBRSaved(list(srcs), z, coef, taps, emit_pool) if save else None
        return a, p, saved
    # eval: BatchNorm folds into the conv epilogue (running statistics); the folded scale / shift are
    # cached until any of the four BN tensors changes (a frozen previous-stage model never recomputes)
    coef = packs.eval_coef(bn, dev)
    a = ops.conv_gemm(src0, src1, wf, taps, cout, bias=bias, scale=coef[0], shift=coef[1], relu=True)
    p = ops.maxpool2x2(a) if emit_pool else None
    return a, p, None   # nothing saved: backward through eval-mode BatchNorm is not supported


def cbr_backward(saved: CBRSaved, dA: Optional[torch.Tensor], dP: Optional[torch.Tensor], conv: nn.Conv2d,
                 bn: nn.BatchNorm2d, packs: PackedWeights, need_dgrad: bool, head=None, raw=None, below=None):
    """Backward of one unit.  Returns (d_srcs list or None, (dW, dbias, dgamma, dbeta)).
    `head` = (dl [B, nc, H, W] fp32, w [nc, C]): gradient of the fused 1x1 head's map; its weight / bias
    gradients are left in `head[2]` (a list the caller reads).
    `raw`: the reduction pass of this unit's BatchNorm backward already ran inside the convolution that
    produced dA (ops.conv_dgrad_bnstats): only the apply pass runs here.
    `below` = CBRSaved of the unit whose output this unit's single source is (the first unit of the same
    DoubleConv): the input gradient is then produced together with THAT unit's BatchNorm-backward sums and
    returned as (d_srcs, grads, raw_below)."""
    src0 = saved.srcs[0]
    src1 = saved.srcs[1] if len(saved.srcs) > 1 else None
    cout = conv.out_channels
    cin = conv.in_channels
    dev = src0.device
    coef = saved.coef
    small = torch.empty((3, cout), dtype=torch.float32, device=dev)   # dgamma, dbeta, dbias (kernel-zeroed)
    if _WHATIF_POOL_AS_PLAIN and dA is not None:
        dP = None
    if head is not None:
        dz, hg = ops.bn_relu_bwd_head(saved.z, dA, coef[0], coef[1], coef[2], coef[3], bn.weight.detach(),
                                      small[:2], small[2], head[0], head[1])
        head[2].append(hg)
    elif raw is not None:
        dz = ops.bn_relu_bwd_apply(saved.z, dA, raw, coef[0], coef[1], coef[2], coef[3], bn.weight.detach(),
                                   small[:2], small[2])
    else:
        dz = ops.bn_relu_bwd(saved.z, dA, dP, coef[0], coef[1], coef[2], coef[3], bn.weight.detach(),
                             small[:2], small[2])
    kdim = sum(s.shape[3] for s in saved.srcs)
    dwp = packs.dwp_buffer(saved.taps * kdim * cout, dev)
    side = WgradOverlap.pick() if packs.defer else None
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())
        WgradOverlap.keep.append((src0, src1, dz))
        with torch.cuda.stream(side):
            ops.conv_wgrad(src0, src1, dz, saved.taps, dwp)
            dW = None
            packs.grad_pending = True
            if packs.on_grad_ready is not None:
                packs.on_grad_ready(packs)         # the reducer forks from the stream the wgrad runs on
    else:
        ops.conv_wgrad(src0, src1, dz, saved.taps, dwp)
        if packs.defer:
            dW = None                      # stays packed in the arena; FusedRMSprop consumes it
            packs.grad_pending = True
            if packs.on_grad_ready is not None:
                packs.on_grad_ready(packs)
        else:
            dW = ops.unpack_wgrad(dwp, cout, cin)
    d_srcs = None
    raw_below = None
    if need_dgrad:
        _, wd = packs.get()
        if below is not None and src1 is None and _FUSED_BN_BWD_REDUCE and \
                ops.dgrad_bnstats_supported(dz.shape[1], dz.shape[2]):
            d, raw_below = ops.conv_dgrad_bnstats(dz, wd, kdim, below.z, below.coef[0], below.coef[1])
            d_srcs = [d]
        elif src1 is None:
            d_srcs = [ops.conv_gemm(dz, None, wd, 9, kdim)]
        else:
            d0, d1 = ops.conv_gemm(dz, None, wd, 9, kdim, split_n=src0.shape[3])
            d_srcs = [d0, d1]
    if side is not None and packs is WgradOverlap.last_pack and WgradOverlap.on_last is not None:
        WgradOverlap.on_last()          # every packed weight gradient and every dgrad is queued now
    if below is not None:
        return d_srcs, (dW, small[2], small[0], small[1]), raw_below
    return d_srcs, (dW, small[2], small[0], small[1])


# ----------------------------------------------------------------------------- layout helpers
def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor (any float dtype / strides) -> contiguous NHWC bf16."""
    return x.permute(0, 2, 3, 1).to(bf16).contiguous()


def nhwc_as_nchw(x: torch.Tensor) -> torch.Tensor:
    """NHWC tensor -> logical NCHW view (channels_last strides, no copy)."""
    return x.permute(0, 3, 1, 2)


def is_fast_act(x: torch.Tensor) -> bool:
    """bf16, logical NCHW with channels_last strides: the zero-copy inter-block format."""
    return x.dtype == bf16 and x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()


# ----------------------------------------------------------------------------- DoubleConv function
class DoubleConvFn(torch.autograd.Function):
    """(optional pre-op) -> CBR -> CBR as one autograd node.

    kind: 'plain'  x NHWC bf16
          'first'  x fp32 NCHW image, first conv runs through im2col (9*Cin <= 64)
          'col'    x already packed first-layer rows [B,H,W,64] (ops.im2col_first[_u8]) or an ops.FirstRows handle
          'pool'   MaxPool2d(2) then DoubleConv  (Down, UG_unet_parts.py:44-54)
          'up'     bilinear x2 of x, concat [skip, up], DoubleConv (Up, UG_unet_parts.py:70-81)
          'cat'    concat [x, skip] walked as two sources (UpSample, UG_unet_parts.py:64-67)
    emit_pool: also return MaxPool2d(2) of the output (fused into the BN/ReLU pass).
    """

    @staticmethod
    def forward(ctx, mod, kind: str, emit_pool: bool, head_mod, x, skip, *params):
        """`head_mod`: an nn.Conv2d(C, nc, 1) applied to this block's output (the PGUNet decoder heads,
        UG_unet.py:294-297) fused into the second BN/ReLU pass (train mode); its weight and bias are the last
        two entries of `params` and the head's map is returned beside the activation."""
        conv1, bn1, conv2, bn2 = mod.conv_op[0], mod.conv_op[1], mod.conv_op[3], mod.conv_op[4]
        need_grad = any(ctx.needs_input_grad)
        # an output nobody differentiates (the last decoder block's activation: only its fused head map is
        # used) arrives in backward as None, not as a materialised tensor of zeros - the BatchNorm-backward
        # kernels take a null activation gradient (saves a 134 MB fill and two 134 MB reads per step)
        ctx.set_materialize_grads(_MATERIALIZE_GRADS)
        head = None
        if head_mod is not None:
            hw, hb = params[8], params[9]
            head = (hw.detach().reshape(hw.shape[0], -1).contiguous(), hb.detach().contiguous())
        pre = None
        if kind == "first":
            srcs, taps = [ops.first_rows(x, cout=conv1.out_channels)], 1      # fused first layer: no packed rows in HBM
        elif kind == "col":
            srcs, taps = [x], 1
        elif kind == "pool":
            pre = x
            srcs, taps = [ops.maxpool2x2(x)], 9
        elif kind == "up":
            srcs, taps = [skip, ops.upsample2x(x)], 9
        elif kind == "cat":
            srcs, taps = [x, skip], 9
        else:
            srcs, taps = [x], 9
        with nvtx_range(f"fwd {kind} {conv1.in_channels}->{conv2.out_channels} @{srcs[0].shape[1]}x{srcs[0].shape[2]}"):
            a1, _, s1 = cbr_forward(srcs, taps, conv1, bn1, mod._packs[0], False, need_grad)
            a2, p2, s2 = cbr_forward([a1], 9, conv2, bn2, mod._packs[1], emit_pool, need_grad, head=head)
        ctx.mod, ctx.kind, ctx.emit_pool = mod, kind, emit_pool
        ctx.s1, ctx.s2, ctx.pre = s1, s2, pre
        ctx.head_w = head[0] if head is not None else None
        ctx.head_shape = tuple(params[8].shape) if head is not None else None
        ctx.x_needs_grad = kind not in ("first", "col") and ctx.needs_input_grad[4]
        ctx.skip_needs_grad = skip is not None and ctx.needs_input_grad[5]
        if emit_pool or head is not None:
            return a2, p2
        return a2

    @staticmethod
    def backward(ctx, dA, dP=None):
        mod = ctx.mod
        conv1, bn1, conv2, bn2 = mod.conv_op[0], mod.conv_op[1], mod.conv_op[3], mod.conv_op[4]
        if ctx.s1 is None or ctx.s2 is None:
            raise RuntimeError("ug_pg_unet_b200: backward through eval-mode BatchNorm is not supported; "
                               "call .train() before a training step or run inference under torch.no_grad()")
        dA = dA.contiguous() if dA is not None else None
        dP = dP.contiguous() if dP is not None else None
        head = None
        if ctx.head_w is not None:
            # the second output is the head's map: its gradient arrives in the pooled-gradient slot
            wait_head_grads(dP)
            dl, dP = dP, None
            head = (dl.float().contiguous(), ctx.head_w, []) if dl is not None else None
            if head is None and dA is None:
                raise RuntimeError("DoubleConvFn.backward: no gradient reached the block")
        with nvtx_range(f"bwd {ctx.kind} {conv1.in_channels}->{conv2.out_channels} @{ctx.s2.z.shape[1]}x{ctx.s2.z.shape[2]}"):
            # the input gradient of conv2 IS the activation gradient of bn1: its reduction pass rides in the
            # dgrad's epilogue (raw1), bn1's backward then only runs the apply pass
            d1, g2, raw1 = cbr_backward(ctx.s2, dA, dP, conv2, bn2, mod._packs[1], True, head=head, below=ctx.s1)
            need_in = ctx.x_needs_grad or ctx.skip_needs_grad
            if _WHATIF_SKIP_BN1_REDUCE:
                set_option("bn_skip_reduce", 1)
            d0, g1 = cbr_backward(ctx.s1, d1[0], None, conv1, bn1, mod._packs[0], need_in, raw=raw1)
            if _WHATIF_SKIP_BN1_REDUCE:
                set_option("bn_skip_reduce", 0)
        dx = dskip = None
        if need_in:
            if ctx.kind == "pool":
                dx = ops.maxpool2x2_bwd(ctx.pre, d0[0])
            elif ctx.kind == "up":
                dskip = d0[0]
                dx = ops.upsample2x_bwd(d0[1])
            elif ctx.kind == "cat":
                dx, dskip = d0[0], d0[1]
            else:
                dx = d0[0]
        ctx.s1 = ctx.s2 = ctx.pre = None
        # params order: conv1.weight, conv1.bias, bn1.weight, bn1.bias, conv2.weight, ... (, head.weight, head.bias)
        grads = (None, None, None, None, dx, dskip, g1[0], g1[1], g1[2], g1[3], g2[0], g2[1], g2[2], g2[3])
        if ctx.head_w is not None:
            if head is not None:
                hg = head[2][0]
                nw = ctx.head_w.numel()
                grads = grads + (hg[:nw].view(ctx.head_shape), hg[nw:])
            else:
                grads = grads + (None, None)
        return grads


def double_conv_params(mod) -> Tuple[torch.Tensor, ...]:
    c1, b1, c2, b2 = mod.conv_op[0], mod.conv_op[1], mod.conv_op[3], mod.conv_op[4]
    return (c1.weight, c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, b2.weight, b2.bias)


def head_is_fusable(mod, head_conv) -> bool:
    """The 1x1 head of a decoder block can ride in the block's second BN/ReLU pass (forward) and
    BatchNorm backward: train-mode BatchNorm (in eval mode BN is folded into the conv epilogue and the head
    runs as ugp_outconv_fwd), a plain nn.Conv2d(C, nc, 1) with bias on CUDA, C / nc the kernels cover."""
    bn2 = mod.conv_op[4]
    return (isinstance(head_conv, nn.Conv2d) and tuple(head_conv.kernel_size) == (1, 1) and head_conv.bias is not None
            and bn2.training and head_conv.weight.is_cuda and head_conv.weight.dtype == torch.float32
            and head_conv.in_channels == bn2.num_features
            and ops.head_fusable(bn2.num_features, head_conv.out_channels)
            and os.environ.get("UGP_NO_FUSED_OUTCONV", "0") != "1")


def run_double_conv(mod, kind: str, x, skip=None, emit_pool: bool = False, head=None):
    """`head`: nn.Conv2d 1x1 fused into the block (see head_is_fusable) -> returns (activation, head map)."""
    if head is not None:
        return DoubleConvFn.apply(mod, kind, False, head, x, skip, *double_conv_params(mod), head.weight, head.bias)
    return DoubleConvFn.apply(mod, kind, emit_pool, None, x, skip, *double_conv_params(mod))


# ----------------------------------------------------------------------------- ConvTranspose2d(k=2, s=2)
class TConvPackedWeights:
    """bf16 GEMM operands of one nn.ConvTranspose2d(k=2, s=2) weight (refreshed like PackedWeights) and
    its persistent packed fp32 weight-gradient accumulator."""

    def __init__(self, conv: nn.ConvTranspose2d):
        self.conv = conv
        self.key = None
        self.wf = None
        self.wd = None
        self._dwp = None

    def get(self):
        w = self.conv.weight
        key = (w.data_ptr(), w._version, w.device, PackedWeights.epoch)
        if key != self.key:
            same = self.wf is not None and self.wf.device == w.device and self.wf.numel() == w.numel()
            self.wf, self.wd = ops.pack_weights_tconv(w.detach(), wf=self.wf if same else None,
                                                      wd=self.wd if same else None)
            self.key = key
        return self.wf, self.wd

    def dwp_buffer(self, device) -> torch.Tensor:
        n = self.conv.weight.numel()
        if self._dwp is None or self._dwp.device != device or self._dwp.numel() != n:
            self._dwp = torch.zeros((n,), dtype=torch.float32, device=device)
        return self._dwp


class ConvTranspose2x2Fn(torch.autograd.Function):
    """nn.ConvTranspose2d(Cin, Cout, 2, 2) on an NHWC bf16 activation (UpSample.up,
    UG_unet_parts.py:61,65): forward, input gradient and weight gradient are tcgen05 GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, packs: TConvPackedWeights):
        wf, _ = packs.get()
        out = ops.conv_transpose2x2_fwd(x, wf, weight.shape[1], bias=None if bias is None else bias.detach())
        ctx.packs, ctx.has_bias = packs, bias is not None
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        packs = ctx.packs
        dout = dout.contiguous()
        cin, cout = packs.conv.weight.shape[0], packs.conv.weight.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            _, wd = packs.get()
            dx = ops.conv_transpose2x2_dgrad(dout, wd, cin)
        if ctx.needs_input_grad[1]:
            dwp = packs.dwp_buffer(x.device)
            ops.conv_transpose2x2_wgrad(x, dout, dwp)
            dw = ops.unpack_wgrad_tconv(dwp, cin, cout)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.channel_sum(dout)
        return dx, dw, db, None


# ----------------------------------------------------------------------------- OutConv / head
def wait_head_grads(dl):
    """Orders the current stream after the head-map adjoints HeadCombineFn.backward put on the side stream -
    unless `dl` is the finest map's gradient, which is dlogits itself and was never on the side stream."""
    ev = WgradOverlap.head_event
    if ev is not None and dl is not None and dl.data_ptr() != WgradOverlap.head_direct_ptr:
        torch.cuda.current_stream().wait_event(ev)


class OutConvFn(torch.autograd.Function):
    """1x1 conv C -> num_classes on an NHWC bf16 activation; fp32 NCHW logits out
    (OutConv, UG_unet_parts.py:84-91)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w2 = weight.detach().reshape(weight.shape[0], -1).contiguous()
        out = ops.outconv_fwd(x, w2, bias.detach())
        ctx.save_for_backward(x, weight)
        return out

    @staticmethod
    def backward(ctx, dl):
        wait_head_grads(dl)
        x, weight = ctx.saved_tensors
        nc = weight.shape[0]
        w2 = weight.detach().reshape(nc, -1).contiguous()
        dx = torch.empty_like(x)
        g = torch.zeros((nc * w2.shape[1] + nc,), dtype=torch.float32, device=x.device)
        dw, db = g[: nc * w2.shape[1]].view(nc, -1), g[nc * w2.shape[1]:]
        ops.outconv_bwd(x, w2, dl.contiguous(), dx, False, dw, db)
        return dx, dw.view_as(weight), db


class HeadCombineFn(torch.autograd.Function):
    """logits = sum_s bilinear_up(map_s), coarse -> fine (UG_unet.py:294-303)."""

    @staticmethod
    def forward(ctx, *maps):
        ctx.sizes = [(m.shape[-2], m.shape[-1]) for m in maps]
        return ops.head_combine_fwd([m.contiguous() for m in maps])

    @staticmethod
    def backward(ctx, dlogits):
        dlogits = dlogits.contiguous()
        full = (dlogits.shape[-2], dlogits.shape[-1])
        side = WgradOverlap.stream
        if side is None or not _HEAD_ADJOINT_ON_SIDE_STREAM:
            # the finest map is added as is: its gradient is dlogits itself (no launch, no copy)
            return tuple(dlogits if tuple(s) == full else ops.head_combine_bwd(dlogits, s) for s in ctx.sizes)
        # Inside a trainer step: the coarse maps' adjoints are first needed one decoder block later (the finest
        # block's backward, ~0.6 ms, runs first on dlogits itself), so their launches go to the weight-gradient
        # side stream, idle at this point, instead of delaying the main stream; the consumers wait for
        # WgradOverlap.head_event (wait_head_grads).
        main = torch.cuda.current_stream()
        B, nc = dlogits.shape[0], dlogits.shape[1]
        outs = [None if tuple(s) == full else torch.empty((B, nc, s[0], s[1]), dtype=torch.float32, device=dlogits.device)
                for s in ctx.sizes]
        side.wait_stream(main)
        WgradOverlap.keep.append(dlogits)
        with torch.cuda.stream(side):
            for s, o in zip(ctx.sizes, outs):
                if o is not None:
                    ops.head_combine_bwd(dlogits, s, out=o)
            ev = torch.cuda.Event()
            ev.record(side)
        WgradOverlap.head_event, WgradOverlap.head_direct_ptr = ev, dlogits.data_ptr()
        return tuple(dlogits if o is None else o for o in outs)


class GlobalAvgPoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d(1) + Flatten on an NHWC bf16 feature map -> fp32 [B, C]."""

    @staticmethod
    def forward(ctx, x):
        ctx.hw = (x.shape[1], x.shape[2])
        return ops.global_avg_pool(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.global_avg_pool_bwd(dy.contiguous().float(), *ctx.hw)


class GapMlpHeadFn(torch.autograd.Function):
    """The Herlev classifier (Herlev/train_herlev.py:66-77: AdaptiveAvgPool2d(1), Flatten, Dropout, Linear,
    ReLU, Dropout, Linear, ReLU, Dropout, Linear) on an NHWC bf16 feature map: one forward kernel, one
    backward kernel for the three Linear layers + the pool adjoint (ugp_global_avg_pool_bwd).
    `u`: [B, C+H1+H2] uniforms for the three Dropout layers in train mode (None = inactive)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, u, p0: float, p1: float, p2: float):
        ws = [t.detach().contiguous() for t in (w1, b1, w2, b2, w3, b3)]
        logits, pooled, h1, h2 = ops.gap_mlp_head_fwd(x, *ws, u=u, p=(p0, p1, p2))
        ctx.save_for_backward(ws[0], ws[2], ws[4], pooled, h1, h2)
        ctx.u, ctx.p, ctx.hw = u, (p0, p1, p2), (x.shape[1], x.shape[2])
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        w1, w2, w3, pooled, h1, h2 = ctx.saved_tensors
        dw1, db1, dw2, db2, dw3, db3, dpooled = ops.mlp_head_bwd(dlogits.contiguous().float(), w1, w2, w3, pooled,
                                                                 h1, h2, u=ctx.u, p=ctx.p)
        dx = ops.global_avg_pool_bwd(dpooled, *ctx.hw) if ctx.needs_input_grad[0] else None
        ctx.u = None
        return dx, dw1, db1, dw2, db2, dw3, db3, None, None, None, None


# ----------------------------------------------------------------------------- losses
class BCEUncertaintyLossFn(torch.autograd.Function):
    """mean(BCEWithLogits(pos_weight)(z, t) * (1 + alpha*A)) with dlogits produced in the same pass;
    `metrics` (float64 [5+3B] device tensor, see ugp_bce_uncertainty_loss) rides along."""

    @staticmethod
    def forward(ctx, z, t, A, alpha: float, pos_weight: float):
        zc = z.detach().contiguous()
        n = zc.numel()
        want = ctx.needs_input_grad[0]
        if t.dtype != torch.uint8 and t.dtype != torch.float32:
            t = t.float()
        out, dlogits = ops.bce_uncertainty_loss(zc, t.contiguous(), None if A is None else A.contiguous(),
                                                alpha, pos_weight, want, 1.0 / n)
        ctx.dlogits = dlogits
        loss = (out[0] / n).float()
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, gloss, _gout):
        d = ctx.dlogits
        ctx.dlogits = None
        if d is None:
            return None, None, None, None, None
        return d * gloss, None, None, None, None


class CEWeightedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, u, alpha: float):
        lc = logits.detach().contiguous()
        B = lc.shape[0]
        want = ctx.needs_input_grad[0]
        out, dl = ops.ce_weighted_loss(lc, labels.contiguous(), None if u is None else u.contiguous(), alpha,
                                       want, 1.0 / B)
        ctx.dl = dl
        ctx.mark_non_differentiable(out)
        return (out[0] / B).float(), out

    @staticmethod
    def backward(ctx, gloss, _gout):
        d = ctx.dl
        ctx.dl = None
        return (None if d is None else d * gloss), None, None, None
