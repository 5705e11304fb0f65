"""Tensor-level wrappers over the C ABI (include/ugp_b200.h).

PyTorch is used for device memory and streams only: every function here takes
CUDA tensors, hands raw device pointers to libugp_b200.so on the current
stream and returns torch tensors that own the results.  Activations are NHWC
bf16 tensors of shape [B, H, W, C] (contiguous).  Nothing here falls back to
torch math: a non-CUDA tensor raises.
"""
from __future__ import annotations

from typing import Optional, Sequence

import ctypes as C
import torch

from ._lib import check as _check_rc, lib

bf16 = torch.bfloat16

# ---- instrumentation used by bench.py: kernel-launch counter and per-launch CUDA-event timing ----
_LAUNCHES = 0
_PROFILE = None     # list of [name, flops, bytes, start_event, end_event] while profiling
_PROFILE_EXTERNAL = False


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


def launch_count() -> int:
    return _LAUNCHES


def start_profile(external: bool = False) -> None:
    """Brackets every launch of the instrumented ops with a pair of CUDA events.  `external=True` creates
    the events with cudaEventRecordExternal semantics so that, when the step is being CAPTURED into a CUDA
    graph, they become event-record NODES between the kernel nodes: a replay then timestamps each kernel on
    the device with no host code between the two records (an eager step is host-bound, so events recorded
    from Python around an eager launch also time the host's tensor-map encode + ctypes call)."""
    global _PROFILE, _PROFILE_EXTERNAL
    _PROFILE, _PROFILE_EXTERNAL = [], external


def profile_events():
    """The live list of [name, flops, bytes, start_event, end_event] (read after a synchronise)."""
    return _PROFILE or []


def stop_profile():
    """-> list of (kernel name, algorithmic flops, algorithmic bytes, milliseconds) per instrumented launch."""
    global _PROFILE
    torch.cuda.synchronize()
    out = [(n, f, by, a.elapsed_time(b)) for (n, f, by, a, b) in (_PROFILE or [])]
    _PROFILE = None
    return out


class _timed:
    def __init__(self, name: str, flops: float = 0.0, nbytes: float = 0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if _PROFILE is not None:
            kw = {"external": True} if _PROFILE_EXTERNAL else {}
            self.a = torch.cuda.Event(enable_timing=True, **kw)
            self.b = torch.cuda.Event(enable_timing=True, **kw)
            self.a.record()

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.b.record()
            _PROFILE.append([self.name, self.flops, self.nbytes, self.a, self.b])
        return False


def _nb(*ts) -> float:
    """Bytes of the given tensors (None skipped): the algorithmic traffic of a kernel that touches each once."""
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def check(rc: int, what: str) -> None:
    _check_rc(rc, what)
    if what not in ("ugp_conv_gemm_bf16", "ugp_conv_dgrad_bnstats_bf16", "ugp_conv_wgrad_bf16", "ugp_bn_relu_bwd_head"):
        _count(2 if what == "ugp_bn_relu_bwd" else 1)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the sm_100a path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


def _opt(t, dtype, name):
    return None if t is None else _req(t, dtype, name)


# ----------------------------------------------------------------------------- convolutions
def conv_gemm(src0: torch.Tensor, src1: Optional[torch.Tensor], wpack: torch.Tensor, taps: int,
              cout: int, *, bias=None, scale=None, shift=None, relu: bool = False, stats=None,
              split_n: Optional[int] = None, out0=None, out1=None):
    """D = conv(src0 ++ src1, wpack) (+bias, *scale+shift, relu); see ugp_conv_gemm_bf16."""
    if isinstance(src0, FirstRows):
        assert src1 is None and taps == 1 and split_n is None and out0 is None
        return _conv_first_fwd(src0, wpack, cout, bias, scale, shift, relu, stats)
    _req(src0, bf16, "src0")
    B, H, W, c0 = src0.shape
    c1 = 0
    if src1 is not None:
        _req(src1, bf16, "src1")
        assert src1.shape[:3] == src0.shape[:3]
        c1 = src1.shape[3]
    _req(wpack, bf16, "wpack")
    assert wpack.numel() == taps * cout * (c0 + c1), (wpack.shape, taps, cout, c0, c1)
    if split_n is None or split_n >= cout:
        split_n = cout
        if out0 is None:
            out0 = torch.empty((B, H, W, cout), dtype=bf16, device=src0.device)
    else:
        if out0 is None:
            out0 = torch.empty((B, H, W, split_n), dtype=bf16, device=src0.device)
        if out1 is None:
            out1 = torch.empty((B, H, W, cout - split_n), dtype=bf16, device=src0.device)
    _req(out0, bf16, "out0")
    if out1 is not None:
        _req(out1, bf16, "out1")
    if stats is not None and tuple(stats.shape) != (stat_parts(), 2, cout):
        raise RuntimeError(f"conv_gemm: stats must be a zeroed [{stat_parts()}, 2, {cout}] buffer (ops.new_stats)")
    with _timed("conv_gemm", 2.0 * B * H * W * cout * taps * (c0 + c1), _nb(src0, src1, wpack, out0, out1)):
        rc = lib().ugp_conv_gemm_bf16(_ptr(src0), c0, _ptr(src1), c1, B, H, W, taps, _ptr(wpack), cout,
                                      _ptr(out0), _ptr(out1), split_n,
                                      _ptr(_opt(bias, torch.float32, "bias")),
                                      _ptr(_opt(scale, torch.float32, "scale")),
                                      _ptr(_opt(shift, torch.float32, "shift")), int(relu),
                                      _ptr(_opt(stats, torch.float32, "stats")), _stream())
    check(rc, "ugp_conv_gemm_bf16")
    _count()
    return (out0, out1) if out1 is not None else out0


def dgrad_bnstats_supported(H: int, W: int) -> bool:
    """Spatial sizes whose convolution tiles are runs of consecutive pixels (the fused BN-backward
    statistics address z by pixel index): the power-of-two maps of the model."""
    bw = W if W < 128 else 128
    return W % bw == 0 and 128 % bw == 0 and H % (128 // bw) == 0


def conv_dgrad_bnstats(dz, wd, cin: int, z, scale, shift):
    """dA = conv3x3(dz, wd) plus the reduction pass of the BatchNorm backward that consumes dA; see
    ugp_conv_dgrad_bnstats_bf16.  Returns (dA NHWC bf16 [B,H,W,cin], raw fp32 [2, cin])."""
    _req(dz, bf16, "dz")
    _req(z, bf16, "z")
    B, H, W, c = dz.shape
    assert tuple(z.shape) == (B, H, W, cin) and wd.numel() == 9 * c * cin
    dA = torch.empty((B, H, W, cin), dtype=bf16, device=dz.device)
    raw = torch.empty((2, cin), dtype=torch.float32, device=dz.device)
    with _timed("conv_gemm", 2.0 * B * H * W * cin * 9 * c, _nb(dz, wd, dA, z)):
        rc = lib().ugp_conv_dgrad_bnstats_bf16(_ptr(dz), c, B, H, W, _ptr(_req(wd, bf16, "wd")), cin, _ptr(dA), _ptr(z),
                                               _ptr(_req(scale, torch.float32, "scale")),
                                               _ptr(_req(shift, torch.float32, "shift")), _ptr(raw), _stream())
    check(rc, "ugp_conv_dgrad_bnstats_bf16")
    _count()
    return dA, raw


def bn_relu_bwd_apply(z, dA, raw, scale, shift, mean, rstd, gamma, sums, dbias):
    """The apply pass of bn_relu_bwd with the sums of conv_dgrad_bnstats; see ugp_bn_relu_bwd_apply."""
    _req(z, bf16, "z")
    B, H, W, Cn = z.shape
    dz = torch.empty_like(z)
    with _timed("bn_relu_bwd_apply", 0.0, _nb(z, dA, dz)):
        rc = lib().ugp_bn_relu_bwd_apply(_ptr(z), _ptr(_req(dA, bf16, "dA")), _ptr(_req(raw, torch.float32, "raw")),
                                         _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), _ptr(gamma), B, H, W, Cn,
                                         _ptr(_req(sums, torch.float32, "sums")),
                                         _ptr(_opt(dbias, torch.float32, "dbias")), _ptr(dz), _stream())
    check(rc, "ugp_bn_relu_bwd_apply")
    return dz


def conv_wgrad(src0, src1, dz, taps: int, dw_packed: torch.Tensor):
    """dw_packed[tap][ci][co] += sum_p src[p+shift, ci] * dz[p, co]."""
    if isinstance(src0, FirstRows):
        assert src1 is None and taps == 1
        _req(dz, bf16, "dz")
        _req(dw_packed, torch.float32, "dw_packed")
        B, H, W, _ = src0.shape
        if lib().ugp_conv_first_wgrad_supported(src0.cin, dz.shape[3], B, H, W):
            return _conv_first_wgrad(src0, dz, dw_packed)
        src0 = src0.materialize()          # e.g. the stage-3 network's 128-wide first layer: packed rows
    _req(src0, bf16, "src0")
    _req(dz, bf16, "dz")
    _req(dw_packed, torch.float32, "dw_packed")
    B, H, W, c0 = src0.shape
    c1 = 0
    if src1 is not None:
        _req(src1, bf16, "src1")
        c1 = src1.shape[3]
    cout = dz.shape[3]
    assert dz.shape[:3] == src0.shape[:3]
    assert dw_packed.numel() == taps * (c0 + c1) * cout
    assert dw_packed.data_ptr() % 16 == 0
    with _timed("conv_wgrad", 2.0 * B * H * W * cout * taps * (c0 + c1), _nb(src0, src1, dz, dw_packed)):
        rc = lib().ugp_conv_wgrad_bf16(_ptr(src0), c0, _ptr(src1), c1, _ptr(dz), cout, B, H, W, taps,
                                       _ptr(dw_packed), _stream())
    check(rc, "ugp_conv_wgrad_bf16")
    _count()
    return dw_packed


def pack_weights(w: torch.Tensor, need_dgrad: bool = True, wf=None, wd=None):
    """fp32 OIHW -> (forward pack, input-gradient pack) bf16; see ugp_pack_weights."""
    _req(w, torch.float32, "weight")
    cout, cin = w.shape[0], w.shape[1]
    first = cin % 64 != 0
    if wf is None:
        wf = torch.empty((1, cout, 64) if first else (9, cout, cin), dtype=bf16, device=w.device)
    if first or not need_dgrad:
        wd = None
    elif wd is None:
        wd = torch.empty((9, cin, cout), dtype=bf16, device=w.device)
    check(lib().ugp_pack_weights(_ptr(w), cout, cin, _ptr(wf), _ptr(wd), _stream()), "ugp_pack_weights")
    return wf, wd


def unpack_wgrad(dw_packed, cout: int, cin: int, grad=None, accumulate: bool = False):
    _req(dw_packed, torch.float32, "dw_packed")
    if grad is None:
        grad = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dw_packed.device)
        accumulate = False
    _req(grad, torch.float32, "grad")
    check(lib().ugp_unpack_wgrad(_ptr(dw_packed), cout, cin, _ptr(grad), int(accumulate), _stream()),
          "ugp_unpack_wgrad")
    return grad


# ----------------------------------------------------------------------------- ConvTranspose2d(k=2, s=2)
def pack_weights_tconv(w: torch.Tensor, wf=None, wd=None):
    """fp32 [Cin, Cout, 2, 2] (nn.ConvTranspose2d.weight) -> (wf [4*Cout, Cin], wd [Cin, 4*Cout]) bf16."""
    _req(w, torch.float32, "weight")
    cin, cout = w.shape[0], w.shape[1]
    assert tuple(w.shape[2:]) == (2, 2)
    if wf is None:
        wf = torch.empty((4 * cout, cin), dtype=bf16, device=w.device)
    if wd is None:
        wd = torch.empty((cin, 4 * cout), dtype=bf16, device=w.device)
    check(lib().ugp_pack_weights_tconv(_ptr(w), cin, cout, _ptr(wf), _ptr(wd), _stream()), "ugp_pack_weights_tconv")
    return wf, wd


def conv_transpose2x2_fwd(x, wf, cout: int, bias=None, out=None):
    """x NHWC bf16 [B,H,W,Cin] -> NHWC bf16 [B,2H,2W,Cout]; see ugp_conv_transpose2x2_fwd."""
    _req(x, bf16, "x")
    _req(wf, bf16, "wf")
    B, H, W, cin = x.shape
    assert wf.numel() == 4 * cout * cin
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, cout), dtype=bf16, device=x.device)
    _req(out, bf16, "out")
    with _timed("conv_transpose_fwd", 2.0 * B * H * W * cin * 4 * cout):
        rc = lib().ugp_conv_transpose2x2_fwd(_ptr(x), B, H, W, cin, _ptr(wf), cout,
                                             _ptr(_opt(bias, torch.float32, "bias")), _ptr(out), _stream())
    check(rc, "ugp_conv_transpose2x2_fwd")
    if 4 * cout > 1024:
        _count()
    return out


def conv_transpose2x2_dgrad(dout, wd, cin: int):
    _req(dout, bf16, "dout")
    _req(wd, bf16, "wd")
    B, H2, W2, cout = dout.shape
    assert wd.numel() == 4 * cout * cin
    dx = torch.empty((B, H2 // 2, W2 // 2, cin), dtype=bf16, device=dout.device)
    with _timed("conv_transpose_dgrad", 2.0 * B * (H2 // 2) * (W2 // 2) * cin * 4 * cout):
        rc = lib().ugp_conv_transpose2x2_dgrad(_ptr(dout), B, H2 // 2, W2 // 2, cout, _ptr(wd), cin, _ptr(dx), _stream())
    check(rc, "ugp_conv_transpose2x2_dgrad")
    return dx


def conv_transpose2x2_wgrad(x, dout, dw_packed):
    """dw_packed[i][ci][j*Cout+co] += sum_p x[p, ci] * dout[phase (i,j) of p, co]."""
    _req(x, bf16, "x")
    _req(dout, bf16, "dout")
    _req(dw_packed, torch.float32, "dw_packed")
    B, H, W, cin = x.shape
    cout = dout.shape[3]
    assert tuple(dout.shape) == (B, 2 * H, 2 * W, cout) and dw_packed.numel() == 4 * cin * cout
    with _timed("conv_transpose_wgrad", 2.0 * B * H * W * cin * 4 * cout):
        rc = lib().ugp_conv_transpose2x2_wgrad(_ptr(x), _ptr(dout), B, H, W, cin, cout, _ptr(dw_packed), _stream())
    check(rc, "ugp_conv_transpose2x2_wgrad")
    _count()
    return dw_packed


def unpack_wgrad_tconv(dw_packed, cin: int, cout: int, grad=None, accumulate: bool = False):
    _req(dw_packed, torch.float32, "dw_packed")
    if grad is None:
        grad = torch.empty((cin, cout, 2, 2), dtype=torch.float32, device=dw_packed.device)
        accumulate = False
    check(lib().ugp_unpack_wgrad_tconv(_ptr(dw_packed), cin, cout, _ptr(_req(grad, torch.float32, "grad")),
                                       int(accumulate), _stream()), "ugp_unpack_wgrad_tconv")
    return grad


def channel_sum(x, out=None):
    """Per-channel sum over all pixels of an NHWC bf16 tensor -> fp32 [C] (added to `out` when given)."""
    _req(x, bf16, "x")
    Cn = x.shape[-1]
    if out is None:
        out = torch.zeros((Cn,), dtype=torch.float32, device=x.device)
    check(lib().ugp_channel_sum(_ptr(x), x.numel() // Cn, Cn, _ptr(_req(out, torch.float32, "out")), _stream()),
          "ugp_channel_sum")
    return out


class FirstRows:
    """The first layer's GEMM rows [B,H,W,64] (3x3 neighbourhoods of the - optionally resized - image) NOT
    written to HBM: `conv_gemm` / `conv_wgrad` given this object as their source run the fused first-layer
    kernels (ugp_conv_first_fwd / ugp_conv_first_wgrad), which build the rows in shared memory.  Quacks
    like the packed tensor where the engine looks at it (`shape`, `device`); `materialize()` packs for real."""
    __slots__ = ("x", "u8", "cin", "Hin", "Win", "H", "W", "mean", "std", "shape", "device", "_mean_c", "_std_c")

    def __init__(self, x, H, W, mean=None, std=None):
        self.x = x
        self.u8 = x.dtype == torch.uint8
        if self.u8:
            B, self.Hin, self.Win, self.cin = x.shape
        else:
            B, self.cin, self.Hin, self.Win = x.shape
        self.H, self.W = H, W
        self.mean = None if mean is None else tuple(float(v) for v in mean)
        self.std = None if std is None else tuple(float(v) for v in std)
        self._mean_c = None if self.mean is None else (C.c_float * self.cin)(*self.mean)
        self._std_c = None if self.std is None else (C.c_float * self.cin)(*self.std)
        self.shape = (B, H, W, 64)
        self.device = x.device

    def materialize(self) -> torch.Tensor:
        if self.u8:
            return im2col_first_u8(self.x, self.H, self.W, self.mean, self.std)
        return im2col_first(self.x, self.H, self.W)


def conv_first_supported(cin: int, cout: int, B: int, H: int, W: int) -> bool:
    return bool(lib().ugp_conv_first_supported(cin, cout, B, H, W))


def first_rows(x: torch.Tensor, H: Optional[int] = None, W: Optional[int] = None, mean=None, std=None, cout: int = 64):
    """The first layer's source: a `FirstRows` handle when the fused first-layer kernels cover the shape
    (Cin = 3, 64 or 128 output channels, B*H*W % 128 == 0), else the packed bf16 [B,H,W,64] tensor.
    x: fp32 NCHW image, or uint8 NHWC loader bytes (mean / std: ToTensor + Normalize folded in)."""
    if x.dtype == torch.uint8:
        B, Hin, Win, cin = x.shape
    else:
        _req(x, torch.float32, "x")
        B, cin, Hin, Win = x.shape
    H = Hin if H is None else H
    W = Win if W is None else W
    rows = FirstRows(x, H, W, mean, std)
    if conv_first_supported(cin, cout, B, H, W):
        return rows
    return rows.materialize()


def _conv_first_fwd(rows: "FirstRows", wpack, cout, bias, scale, shift, relu, stats):
    B, H, W, _ = rows.shape
    assert wpack.numel() == cout * 64, (wpack.shape, cout)
    out = torch.empty((B, H, W, cout), dtype=bf16, device=rows.device)
    if stats is not None and tuple(stats.shape) != (stat_parts(), 2, cout):
        raise RuntimeError(f"conv_first: stats must be a zeroed [{stat_parts()}, 2, {cout}] buffer (ops.new_stats)")
    with _timed("conv_first", 2.0 * B * H * W * cout * 9 * rows.cin, _nb(rows.x, wpack, out)):
        rc = lib().ugp_conv_first_fwd(_ptr(rows.x), int(rows.u8), B, rows.cin, rows.Hin, rows.Win, H, W, rows._mean_c,
                                      rows._std_c, _ptr(_req(wpack, bf16, "wpack")), cout, _ptr(out),
                                      _ptr(_opt(bias, torch.float32, "bias")), _ptr(_opt(scale, torch.float32, "scale")),
                                      _ptr(_opt(shift, torch.float32, "shift")), int(relu),
                                      _ptr(_opt(stats, torch.float32, "stats")), _stream())
    check(rc, "ugp_conv_first_fwd")
    return out


def _conv_first_wgrad(rows: "FirstRows", dz, dw_packed):
    B, H, W, _ = rows.shape
    cout = dz.shape[3]
    assert tuple(dz.shape[:3]) == (B, H, W) and dw_packed.numel() == 64 * cout
    with _timed("conv_first_wgrad", 2.0 * B * H * W * cout * 9 * rows.cin, _nb(rows.x, dz, dw_packed)):
        rc = lib().ugp_conv_first_wgrad(_ptr(rows.x), int(rows.u8), B, rows.cin, rows.Hin, rows.Win, H, W, rows._mean_c,
                                        rows._std_c, _ptr(dz), cout, _ptr(dw_packed), _stream())
    check(rc, "ugp_conv_first_wgrad")
    return dw_packed


def im2col_first(x: torch.Tensor, H: Optional[int] = None, W: Optional[int] = None):
    """fp32 NCHW image -> bf16 [B,H,W,64] 3x3 neighbourhoods (optional bilinear resize)."""
    _req(x, torch.float32, "x")
    B, cin, Hin, Win = x.shape
    H = Hin if H is None else H
    W = Win if W is None else W
    out = torch.empty((B, H, W, 64), dtype=bf16, device=x.device)
    with _timed("im2col_first", 0.0, _nb(x, out)):
        rc = lib().ugp_im2col_first(_ptr(x), B, cin, Hin, Win, H, W, _ptr(out), _stream())
    check(rc, "ugp_im2col_first")
    return out


def im2col_first_u8(x: torch.Tensor, H: Optional[int] = None, W: Optional[int] = None, mean=None, std=None):
    """uint8 HWC image bytes [B,Hin,Win,Cin] -> bf16 [B,H,W,64] first-layer rows with ToTensor's /255, an
    optional per-channel Normalize(mean, std) (python sequences) and the bilinear resize folded in."""
    _req(x, torch.uint8, "x")
    B, Hin, Win, cin = x.shape
    H = Hin if H is None else H
    W = Win if W is None else W
    m = None if mean is None else (C.c_float * cin)(*[float(v) for v in mean])
    sd = None if std is None else (C.c_float * cin)(*[float(v) for v in std])
    out = torch.empty((B, H, W, 64), dtype=bf16, device=x.device)
    check(lib().ugp_im2col_first_u8(_ptr(x), B, cin, Hin, Win, H, W, m, sd, _ptr(out), _stream()),
          "ugp_im2col_first_u8")
    return out


def flip_u8(x: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """Per-sample horizontal (bit 0) / vertical (bit 1) flips of uint8 [B,H,W,C] loader bytes (or [B,1,H,W]
    masks, viewed as [B,H,W,1]); see ugp_flip_u8.  `flags`: uint8 [B] on the device."""
    _req(x, torch.uint8, "x")
    _req(flags, torch.uint8, "flags")
    B, H, W, Cn = x.shape
    assert flags.numel() == B
    out = torch.empty_like(x)
    check(lib().ugp_flip_u8(_ptr(x), _ptr(flags), B, H, W, Cn, _ptr(out), _stream()), "ugp_flip_u8")
    return out


def joint_flip_u8(images: torch.Tensor, masks: torch.Tensor, generator=None):
    """The flip part of the reference's train-time augmentation (monuseg_dataset.py:150-158) on a device
    batch: one (hflip, vflip) decision per sample with probability 1/2 each, applied to the image bytes
    [B,H,W,3] and the mask bytes [B,1,H,W] alike.  Returns (images, masks, flags)."""
    B = images.shape[0]
    flags = torch.randint(0, 4, (B,), device=images.device, dtype=torch.uint8, generator=generator)
    m = masks.reshape(B, masks.shape[-2], masks.shape[-1], 1)
    return flip_u8(images, flags), flip_u8(m.contiguous(), flags).reshape(masks.shape), flags


# ----------------------------------------------------------------------------- BN / pool / upsample
_STAT_PARTS = None


def stat_parts() -> int:
    """Slots of a partial-statistics buffer = SM count (one per CTA of a convolution launch)."""
    global _STAT_PARTS
    if _STAT_PARTS is None:
        _STAT_PARTS = int(lib().ugp_device_sm_count())
    return _STAT_PARTS


def new_stats(cout: int, device) -> torch.Tensor:
    """Zeroed [P][2][cout] buffer for conv_gemm(..., stats=); `stats.sum(0)` gives [sum, sum of squares]."""
    return torch.zeros((stat_parts(), 2, cout), dtype=torch.float32, device=device)


def bn_finalize(stats, count: int, gamma, beta, eps: float, momentum: float, running_mean,
                running_var, num_batches_tracked, scale, shift, save_mean, save_rstd):
    Cn = gamma.numel()
    assert stats.dim() == 3 and stats.shape[1] == 2 and stats.shape[2] == Cn, stats.shape
    with _timed("bn_finalize", 0.0, _nb(stats)):
        rc = lib().ugp_bn_finalize(_ptr(_req(stats, torch.float32, "stats")), stats.shape[0], Cn, count, _ptr(gamma),
                                   _ptr(beta), eps, momentum, _ptr(running_mean), _ptr(running_var),
                                   _ptr(num_batches_tracked), _ptr(scale), _ptr(shift), _ptr(save_mean),
                                   _ptr(save_rstd), _stream())
    check(rc, "ugp_bn_finalize")


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps: float, scale, shift):
    check(lib().ugp_bn_eval_coeffs(gamma.numel(), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                   _ptr(running_var), eps, _ptr(scale), _ptr(shift), _stream()),
          "ugp_bn_eval_coeffs")


def bn_relu_apply(z, scale, shift, pool: bool = False, out=None, pooled=None):
    _req(z, bf16, "z")
    B, H, W, Cn = z.shape
    if out is None:
        out = torch.empty_like(z)
    if pool and pooled is None:
        pooled = torch.empty((B, H // 2, W // 2, Cn), dtype=bf16, device=z.device)
    with _timed("bn_relu_apply_pool" if pool else "bn_relu_apply", 0.0, _nb(z, out, pooled if pool else None)):
        rc = lib().ugp_bn_relu_apply(_ptr(z), _ptr(scale), _ptr(shift), B, H, W, Cn, _ptr(out),
                                     _ptr(pooled) if pool else None, _stream())
    check(rc, "ugp_bn_relu_apply")
    return (out, pooled) if pool else out


def bn_relu_apply_head(z, scale, shift, head_w, head_b):
    """bn_relu_apply + the 1x1 head on the activation in one pass; see ugp_bn_relu_apply_head.
    head_w fp32 [nc, C], head_b fp32 [nc] -> (a NHWC bf16, lmap fp32 NCHW [B, nc, H, W])."""
    _req(z, bf16, "z")
    B, H, W, Cn = z.shape
    nc = head_w.shape[0]
    out = torch.empty_like(z)
    lmap = torch.empty((B, nc, H, W), dtype=torch.float32, device=z.device)
    with _timed("bn_relu_apply_head", 0.0, _nb(z, out, lmap)):
        rc = lib().ugp_bn_relu_apply_head(_ptr(z), _ptr(scale), _ptr(shift), B, H, W, Cn,
                                          _ptr(_req(head_w, torch.float32, "head_w")),
                                          _ptr(_req(head_b, torch.float32, "head_b")), nc, _ptr(out), _ptr(lmap),
                                          _stream())
    check(rc, "ugp_bn_relu_apply_head")
    return out, lmap


def head_fusable(C: int, nc: int) -> bool:
    """Shapes ugp_bn_relu_apply_head / ugp_bn_relu_bwd_head cover."""
    return C in (64, 128, 256) and nc in (1, 2)


def bn_relu_bwd_head(z, dA, scale, shift, mean, rstd, gamma, sums, dbias, head_dl, head_w):
    """bn_relu_bwd with the fused head gradient; see ugp_bn_relu_bwd_head.
    Returns (dz, head_grads fp32 [nc*C + nc]: weight gradient then bias gradient)."""
    _req(z, bf16, "z")
    B, H, W, Cn = z.shape
    nc = head_w.shape[0]
    _req(head_dl, torch.float32, "head_dl")
    assert tuple(head_dl.shape) == (B, nc, H, W), head_dl.shape
    dz = torch.empty_like(z)
    hg = torch.empty((nc * Cn + nc,), dtype=torch.float32, device=z.device)
    with _timed("bn_relu_bwd_head", 0.0, 2 * _nb(z, dA, head_dl) + _nb(dz)):
        rc = lib().ugp_bn_relu_bwd_head(_ptr(z), _ptr(_opt(dA, bf16, "dA")), _ptr(scale), _ptr(shift), _ptr(mean),
                                        _ptr(rstd), _ptr(gamma), B, H, W, Cn, _ptr(head_dl),
                                        _ptr(_req(head_w, torch.float32, "head_w")), nc,
                                        _ptr(_req(sums, torch.float32, "sums")),
                                        _ptr(_opt(dbias, torch.float32, "dbias")), _ptr(dz), _ptr(hg), _stream())
    check(rc, "ugp_bn_relu_bwd_head")
    _count(2)
    return dz, hg


def bn_relu_bwd(z, dA, dP, scale, shift, mean, rstd, gamma, sums, dbias, dz=None):
    _req(z, bf16, "z")
    B, H, W, Cn = z.shape
    if dz is None:
        dz = torch.empty_like(z)
    with _timed("bn_relu_bwd_pool" if dP is not None else "bn_relu_bwd", 0.0, 2 * _nb(z, dA, dP) + _nb(dz)):
        rc = lib().ugp_bn_relu_bwd(_ptr(z), _ptr(_opt(dA, bf16, "dA")), _ptr(_opt(dP, bf16, "dP")),
                                   _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), _ptr(gamma), B, H,
                                   W, Cn, _ptr(_req(sums, torch.float32, "sums")),
                                   _ptr(_opt(dbias, torch.float32, "dbias")), _ptr(dz), _stream())
    check(rc, "ugp_bn_relu_bwd")
    return dz


def maxpool2x2(x, return_indices: bool = False):
    _req(x, bf16, "x")
    B, H, W, Cn = x.shape
    out = torch.empty((B, H // 2, W // 2, Cn), dtype=bf16, device=x.device)
    idx = torch.empty((B, H // 2, W // 2, Cn), dtype=torch.int32, device=x.device) if return_indices else None
    with _timed("maxpool2x2", 0.0, _nb(x, out, idx)):
        rc = lib().ugp_maxpool2x2(_ptr(x), B, H, W, Cn, _ptr(out), _ptr(idx), _stream())
    check(rc, "ugp_maxpool2x2")
    return (out, idx) if return_indices else out


def maxpool2x2_bwd(x, dout):
    _req(x, bf16, "x")
    _req(dout, bf16, "dout")
    B, H, W, Cn = x.shape
    dx = torch.empty_like(x)
    check(lib().ugp_maxpool2x2_bwd(_ptr(x), _ptr(dout), B, H, W, Cn, _ptr(dx), _stream()), "ugp_maxpool2x2_bwd")
    return dx


def upsample2x(x):
    _req(x, bf16, "x")
    B, h, w, Cn = x.shape
    out = torch.empty((B, 2 * h, 2 * w, Cn), dtype=bf16, device=x.device)
    with _timed("upsample2x", 0.0, _nb(x, out)):
        rc = lib().ugp_upsample2x(_ptr(x), B, h, w, Cn, _ptr(out), _stream())
    check(rc, "ugp_upsample2x")
    return out


def upsample2x_bwd(dout):
    _req(dout, bf16, "dout")
    B, H, W, Cn = dout.shape
    dx = torch.empty((B, H // 2, W // 2, Cn), dtype=bf16, device=dout.device)
    with _timed("upsample2x_bwd", 0.0, _nb(dout, dx)):
        rc = lib().ugp_upsample2x_bwd(_ptr(dout), B, H // 2, W // 2, Cn, _ptr(dx), _stream())
    check(rc, "ugp_upsample2x_bwd")
    return dx


def global_avg_pool(x):
    """NHWC bf16 [B,H,W,C] -> fp32 [B,C] (AdaptiveAvgPool2d(1) + Flatten)."""
    _req(x, bf16, "x")
    B, H, W, Cn = x.shape
    out = torch.empty((B, Cn), dtype=torch.float32, device=x.device)
    check(lib().ugp_global_avg_pool(_ptr(x), B, H * W, Cn, _ptr(out), _stream()), "ugp_global_avg_pool")
    return out


def global_avg_pool_bwd(dy, H: int, W: int):
    _req(dy, torch.float32, "dy")
    B, Cn = dy.shape
    dx = torch.empty((B, H, W, Cn), dtype=bf16, device=dy.device)
    check(lib().ugp_global_avg_pool_bwd(_ptr(dy), B, H * W, Cn, _ptr(dx), _stream()), "ugp_global_avg_pool_bwd")
    return dx


# ----------------------------------------------------------------------------- Herlev classification head
def gap_mlp_head_fwd(x, w1, b1, w2, b2, w3, b3, u=None, p=(0.0, 0.0, 0.0), tickets=None):
    """AdaptiveAvgPool2d(1) + Flatten + [Dropout, Linear, ReLU] x 2 + Dropout + Linear on an NHWC bf16
    feature map in ONE launch; see ugp_gap_mlp_head_fwd.  Returns (logits, pooled, h1, h2)."""
    _req(x, bf16, "x")
    B, H, W, Cn = x.shape
    H1, H2, K = w1.shape[0], w2.shape[0], w3.shape[0]
    for t, shape, name in ((w1, (H1, Cn), "w1"), (b1, (H1,), "b1"), (w2, (H2, H1), "w2"), (b2, (H2,), "b2"),
                           (w3, (K, H2), "w3"), (b3, (K,), "b3")):
        _req(t, torch.float32, name)
        if tuple(t.shape) != shape:
            raise RuntimeError(f"gap_mlp_head: {name} has shape {tuple(t.shape)}, expected {shape}")
    if u is not None:
        _req(u, torch.float32, "u")
        assert tuple(u.shape) == (B, Cn + H1 + H2), u.shape
    dev = x.device
    if tickets is None:
        tickets = torch.zeros((B,), dtype=torch.int32, device=dev)
    pooled = torch.empty((B, Cn), dtype=torch.float32, device=dev)
    h1 = torch.empty((B, H1), dtype=torch.float32, device=dev)
    h2 = torch.empty((B, H2), dtype=torch.float32, device=dev)
    logits = torch.empty((B, K), dtype=torch.float32, device=dev)
    with _timed("gap_mlp_head_fwd", 0.0, _nb(x)):
        rc = lib().ugp_gap_mlp_head_fwd(_ptr(x), B, H * W, Cn, _ptr(w1), _ptr(b1), H1, _ptr(w2), _ptr(b2), H2,
                                        _ptr(w3), _ptr(b3), K, _ptr(u), float(p[0]), float(p[1]), float(p[2]),
                                        _ptr(pooled), _ptr(h1), _ptr(h2), _ptr(logits), _ptr(tickets), _stream())
    check(rc, "ugp_gap_mlp_head_fwd")
    return logits, pooled, h1, h2


def mlp_head_bwd(dlogits, w1, w2, w3, pooled, h1, h2, u=None, p=(0.0, 0.0, 0.0)):
    """Gradients of the three Linear layers and of the pooled vector in one launch; see ugp_mlp_head_bwd.
    Returns (dw1, db1, dw2, db2, dw3, db3, dpooled)."""
    _req(dlogits, torch.float32, "dlogits")
    B, K = dlogits.shape
    Cn, H1, H2 = w1.shape[1], w1.shape[0], w2.shape[0]
    dev = dlogits.device
    flat = torch.empty((w1.numel() + H1 + w2.numel() + H2 + w3.numel() + K,), dtype=torch.float32, device=dev)
    outs, off = [], 0
    for shape in (w1.shape, (H1,), w2.shape, (H2,), w3.shape, (K,)):
        n = 1
        for d in shape:
            n *= d
        outs.append(flat[off:off + n].view(shape))
        off += n
    dpooled = torch.empty((B, Cn), dtype=torch.float32, device=dev)
    scratch = torch.empty((B * (H1 + H2),), dtype=torch.float32, device=dev)
    with _timed("mlp_head_bwd", 0.0, _nb(w1, w2, w3, flat)):
        rc = lib().ugp_mlp_head_bwd(_ptr(dlogits), B, Cn, _ptr(w1), H1, _ptr(w2), H2, _ptr(w3), K, _ptr(u),
                                    float(p[0]), float(p[1]), float(p[2]), _ptr(pooled), _ptr(h1), _ptr(h2),
                                    *[_ptr(t) for t in outs], _ptr(dpooled), _ptr(scratch), _stream())
    check(rc, "ugp_mlp_head_bwd")
    return (*outs, dpooled)


# ----------------------------------------------------------------------------- head / loss
def outconv_fwd(x, w, bias):
    """x NHWC bf16 [B,h,w,C], w fp32 [nc,C(,1,1)] -> fp32 NCHW [B,nc,h,w]."""
    _req(x, bf16, "x")
    B, h, wd, Cn = x.shape
    nc = w.shape[0]
    out = torch.empty((B, nc, h, wd), dtype=torch.float32, device=x.device)
    with _timed("outconv_fwd", 0.0, _nb(x, out)):
        rc = lib().ugp_outconv_fwd(_ptr(x), _ptr(_req(w, torch.float32, "w")), _ptr(bias), B, h * wd, Cn, nc,
                                   _ptr(out), _stream())
    check(rc, "ugp_outconv_fwd")
    return out


def outconv_bwd(x, w, dl, dx, accumulate: bool, dw, db):
    _req(x, bf16, "x")
    _req(dl, torch.float32, "dl")
    B, h, wd, Cn = x.shape
    nc = w.shape[0]
    with _timed("outconv_bwd", 0.0, _nb(x, dl, dx) + (_nb(dx) if accumulate else 0.0)):
        rc = lib().ugp_outconv_bwd(_ptr(x), _ptr(w), _ptr(dl), B, h * wd, Cn, nc, int(accumulate),
                                   _ptr(_req(dx, bf16, "dx")), _ptr(_req(dw, torch.float32, "dw")),
                                   _ptr(_req(db, torch.float32, "db")), _stream())
    check(rc, "ugp_outconv_bwd")
    return dx


def head_combine_fwd(maps: Sequence[torch.Tensor]):
    """logits = sum of the maps, each bilinearly upsampled (align_corners) to the last one's size."""
    n = len(maps)
    B, nc, RH, RW = maps[-1].shape
    logits = torch.empty_like(maps[-1])
    ptrs = (C.c_void_p * n)(*[_req(m, torch.float32, "map").data_ptr() for m in maps])
    hs = (C.c_int * n)(*[m.shape[-2] for m in maps])
    ws = (C.c_int * n)(*[m.shape[-1] for m in maps])
    with _timed("head_combine_fwd", 0.0, _nb(*maps) + _nb(logits)):
        rc = lib().ugp_head_combine_fwd_hw(ptrs, hs, ws, n, B * nc, RH, RW, _ptr(logits), _stream())
    check(rc, "ugp_head_combine_fwd")
    return logits


def head_combine_bwd(dlogits, size, out=None):
    """`size`: side of a square map or (h, w); `out`: preallocated fp32 [B, nc, h, w] result."""
    _req(dlogits, torch.float32, "dlogits")
    B, nc, RH, RW = dlogits.shape
    h, w = (size, size) if isinstance(size, int) else size
    dl = torch.empty((B, nc, h, w), dtype=torch.float32, device=dlogits.device) if out is None else out
    assert tuple(dl.shape) == (B, nc, h, w) and dl.dtype == torch.float32 and dl.is_contiguous()
    with _timed("head_combine_bwd", 0.0, _nb(dlogits, dl)):
        rc = lib().ugp_head_combine_bwd_hw(_ptr(dlogits), B * nc, RH, RW, h, w, _ptr(dl), _stream())
    check(rc, "ugp_head_combine_bwd")
    return dl


def uncertainty_margin(zprev, H: int, W: int):
    _req(zprev, torch.float32, "zprev")
    B, nc, h, w = zprev.shape
    A = torch.empty((B, nc, H, W), dtype=torch.float32, device=zprev.device)
    with _timed("uncertainty_margin", 0.0, _nb(zprev, A)):
        rc = lib().ugp_uncertainty_margin(_ptr(zprev), B * nc, h, w, H, W, _ptr(A), _stream())
    check(rc, "ugp_uncertainty_margin")
    return A


def softmax_entropy(logits):
    _req(logits, torch.float32, "logits")
    B, K = logits.shape
    u = torch.empty((B,), dtype=torch.float32, device=logits.device)
    check(lib().ugp_softmax_entropy(_ptr(logits), B, K, _ptr(u), _stream()), "ugp_softmax_entropy")
    return u


def bce_uncertainty_loss(z, t, A, alpha: float, pos_weight: float, want_grad: bool, grad_scale: float):
    """Returns (out double[5+3B], dlogits or None); see ugp_bce_uncertainty_loss."""
    _req(z, torch.float32, "logits")
    u8 = t.dtype == torch.uint8          # mask bytes straight from the loader
    _req(t, torch.uint8 if u8 else torch.float32, "target")
    if t.shape != z.shape:
        raise ValueError(f"Target size ({tuple(t.shape)}) must be the same as input size ({tuple(z.shape)})")
    B = z.shape[0]
    per_image = z.numel() // B
    out = torch.empty((5 + 3 * B,), dtype=torch.float64, device=z.device)
    dlogits = torch.empty_like(z) if want_grad else None
    fn = lib().ugp_bce_uncertainty_loss_u8 if u8 else lib().ugp_bce_uncertainty_loss
    with _timed("bce_uncertainty_loss", 0.0, _nb(z, t, A, dlogits)):
        rc = fn(_ptr(z), _ptr(t), _ptr(_opt(A, torch.float32, "uncertainty")), alpha, pos_weight, B, per_image,
                grad_scale, _ptr(dlogits), _ptr(out), _stream())
    check(rc, "ugp_bce_uncertainty_loss")
    return out, dlogits


def seg_confusion(z, gt=None, want_mask: bool = False):
    """Per-image {TP, sum pred, sum gt, sum sigmoid} (float64 [B,4]) of pred = z > 0 against gt (fp32 or
    uint8 0/1, optional) and, when asked, the uint8 prediction mask; see ugp_seg_confusion."""
    _req(z, torch.float32, "logits")
    B = z.shape[0]
    per_image = z.numel() // B
    u8 = gt is not None and gt.dtype == torch.uint8
    if gt is not None:
        _req(gt, torch.uint8 if u8 else torch.float32, "gt")
        if gt.numel() != z.numel():
            raise ValueError(f"ground truth size {tuple(gt.shape)} does not match logits {tuple(z.shape)}")
    out = torch.empty((B, 4), dtype=torch.float64, device=z.device)
    mask = torch.empty(z.shape, dtype=torch.uint8, device=z.device) if want_mask else None
    check(lib().ugp_seg_confusion(_ptr(z), _ptr(gt), int(u8), B, per_image, _ptr(mask), _ptr(out), _stream()),
          "ugp_seg_confusion")
    return out, mask


def cls_eval(logits, labels=None, confusion=None):
    """softmax probabilities, argmax predictions and (with labels) the confusion matrix update."""
    _req(logits, torch.float32, "logits")
    B, K = logits.shape
    probs = torch.empty_like(logits)
    preds = torch.empty((B,), dtype=torch.int64, device=logits.device)
    if labels is not None:
        _req(labels, torch.int64, "labels")
        if confusion is None:
            confusion = torch.zeros((K, K), dtype=torch.int64, device=logits.device)
        _req(confusion, torch.int64, "confusion")
    check(lib().ugp_cls_eval(_ptr(logits), _ptr(labels), B, K, _ptr(probs), _ptr(preds),
                             _ptr(confusion) if labels is not None else None, _stream()), "ugp_cls_eval")
    return probs, preds, confusion


def ce_weighted_loss(logits, labels, u, alpha: float, want_grad: bool, grad_scale: float):
    _req(logits, torch.float32, "logits")
    _req(labels, torch.int64, "labels")
    B, K = logits.shape
    out = torch.empty((3,), dtype=torch.float64, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    check(lib().ugp_ce_weighted_loss(_ptr(logits), _ptr(labels), _ptr(_opt(u, torch.float32, "u")), alpha, B, K,
                                     grad_scale, _ptr(dlogits), _ptr(out), _stream()), "ugp_ce_weighted_loss")
    return out, dlogits


def rmsprop_step(param, grad, square_avg, lr: float, alpha: float = 0.99, eps: float = 1e-8,
                 weight_decay: float = 0.0, grad_scale: float = 1.0):
    _req(param, torch.float32, "param")
    _req(grad, torch.float32, "grad")
    _req(square_avg, torch.float32, "square_avg")
    check(lib().ugp_rmsprop_step(_ptr(param), _ptr(grad), _ptr(square_avg), param.numel(), lr, alpha, eps,
                                 weight_decay, grad_scale, _stream()), "ugp_rmsprop_step")


def rmsprop_multi(params, grads, square_avgs, lr: float, alpha: float, eps: float, weight_decay: float,
                  grad_scale: float = 1.0):
    """One launch of the RMSprop update over a list of tensors (see ugp_rmsprop_multi)."""
    n = len(params)
    for t in list(params) + list(grads) + list(square_avgs):
        _req(t, torch.float32, "rmsprop tensor")
    P = (C.c_void_p * n)(*[t.data_ptr() for t in params])
    G = (C.c_void_p * n)(*[t.data_ptr() for t in grads])
    V = (C.c_void_p * n)(*[t.data_ptr() for t in square_avgs])
    S = (C.c_longlong * n)(*[t.numel() for t in params])
    check(lib().ugp_rmsprop_multi(P, G, V, S, n, lr, alpha, eps, weight_decay, grad_scale, _stream()),
          "ugp_rmsprop_multi")


def counter_add(counter: torch.Tensor, inc: int = 1):
    _req(counter, torch.int64, "counter")
    check(lib().ugp_counter_add(_ptr(counter), int(inc), _stream()), "ugp_counter_add")


def adam_multi(params, grads, exp_avgs, exp_avg_sqs, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, grad_scale: float = 1.0, step_dev: Optional[torch.Tensor] = None):
    """One launch (per 64 tensors) of torch.optim.Adam's update with L2 weight decay; see ugp_adam_multi.
    `step_dev`: int64 device scalar holding the step count (graph-safe), overrides `step`."""
    n = len(params)
    for t in list(params) + list(grads) + list(exp_avgs) + list(exp_avg_sqs):
        _req(t, torch.float32, "adam tensor")
    P = (C.c_void_p * n)(*[t.data_ptr() for t in params])
    G = (C.c_void_p * n)(*[t.data_ptr() for t in grads])
    M = (C.c_void_p * n)(*[t.data_ptr() for t in exp_avgs])
    V = (C.c_void_p * n)(*[t.data_ptr() for t in exp_avg_sqs])
    S = (C.c_longlong * n)(*[t.numel() for t in params])
    check(lib().ugp_adam_multi(P, G, M, V, S, n, lr, beta1, beta2, eps, weight_decay, int(step),
                               _ptr(_opt(step_dev, torch.int64, "step_dev")), grad_scale, _stream()), "ugp_adam_multi")


def rmsprop_conv_packed(weights, square_avgs, dw_packed, wfs, wds, lr: float, alpha: float, eps: float,
                        weight_decay: float, grad_scale: float = 1.0):
    """RMSprop on 3x3 conv weights straight from the packed gradients; rewrites the bf16 operand
    packs and zeroes the gradients (see ugp_rmsprop_conv_packed).  `wds` entries may be None."""
    n = len(weights)
    for w, v, g, wf in zip(weights, square_avgs, dw_packed, wfs):
        _req(w, torch.float32, "weight")
        _req(v, torch.float32, "square_avg")
        _req(g, torch.float32, "dw_packed")
        _req(wf, bf16, "wf")
        assert w.dim() == 4 and w.shape[2:] == (3, 3) and v.shape == w.shape
        assert g.numel() == w.numel() and wf.numel() == w.numel()
    Wp = (C.c_void_p * n)(*[t.data_ptr() for t in weights])
    Vp = (C.c_void_p * n)(*[t.data_ptr() for t in square_avgs])
    Gp = (C.c_void_p * n)(*[t.data_ptr() for t in dw_packed])
    Fp = (C.c_void_p * n)(*[t.data_ptr() for t in wfs])
    Dp = (C.c_void_p * n)(*[None if t is None else _req(t, bf16, "wd").data_ptr() for t in wds])
    co = (C.c_int * n)(*[w.shape[0] for w in weights])
    ci = (C.c_int * n)(*[w.shape[1] for w in weights])
    # 28 B / parameter: read w, v, g (12), write w, v, zeroed g (12), write both bf16 packs (4)
    with _timed("rmsprop_conv_packed", 0.0, 28.0 * sum(w.numel() for w in weights)):
        rc = lib().ugp_rmsprop_conv_packed(Wp, Vp, Gp, Fp, Dp, co, ci, n, lr, alpha, eps, weight_decay, grad_scale,
                                           _stream())
    check(rc, "ugp_rmsprop_conv_packed")
