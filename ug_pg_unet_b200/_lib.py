"""ctypes binding of libugp_b200.so (the C ABI in include/ugp_b200.h).

There is no fallback: if the shared library is missing or a call fails, the
caller gets an exception.  `lib()` builds the library on first use when nvcc is
available (the GPU box receives the prebuilt .so with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# UGP_LIB_PATH: load another build of the same library (A/B timing of kernel changes on one box)
LIB_PATH = Path(os.environ["UGP_LIB_PATH"]).resolve() if os.environ.get("UGP_LIB_PATH") else _PKG / "libugp_b200.so"

_p, _i, _f, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double

# name -> argtypes; every function returns int (0 = ok) unless noted
SIGNATURES = {
    "ugp_conv_gemm_bf16": [_p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p, _p, _p, _i, _p, _p],
    "ugp_conv_wgrad_bf16": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _p],
    "ugp_conv_transpose2x2_fwd": [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p],
    "ugp_conv_transpose2x2_dgrad": [_p, _i, _i, _i, _i, _p, _i, _p, _p],
    "ugp_conv_transpose2x2_wgrad": [_p, _p, _i, _i, _i, _i, _i, _p, _p],
    "ugp_pack_weights_tconv": [_p, _i, _i, _p, _p, _p],
    "ugp_unpack_wgrad_tconv": [_p, _i, _i, _p, _i, _p],
    "ugp_channel_sum": [_p, _ll, _i, _p, _p],
    "ugp_im2col_first_u8": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "ugp_bce_uncertainty_loss_u8": [_p, _p, _p, _f, _f, _i, _ll, _f, _p, _p, _p],
    "ugp_seg_confusion": [_p, _p, _i, _i, _ll, _p, _p, _p],
    "ugp_cls_eval": [_p, _p, _i, _i, _p, _p, _p, _p],
    "ugp_pack_weights": [_p, _i, _i, _p, _p, _p],
    "ugp_unpack_wgrad": [_p, _i, _i, _p, _i, _p],
    "ugp_conv_first_supported": [_i, _i, _i, _i, _i],
    "ugp_conv_first_wgrad_supported": [_i, _i, _i, _i, _i],
    "ugp_conv_first_fwd": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _p],
    "ugp_conv_first_wgrad": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p],
    "ugp_im2col_first": [_p, _i, _i, _i, _i, _i, _i, _p, _p],
    "ugp_bn_finalize": [_p, _i, _i, _ll, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p],
    "ugp_bn_eval_coeffs": [_i, _p, _p, _p, _p, _f, _p, _p, _p],
    "ugp_bn_relu_apply": [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p],
    "ugp_bn_relu_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p],
    "ugp_bn_relu_apply_head": [_p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p],
    "ugp_bn_relu_bwd_head": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p],
    "ugp_conv_dgrad_bnstats_bf16": [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p],
    "ugp_bn_relu_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p],
    "ugp_flip_u8": [_p, _p, _i, _i, _i, _i, _p, _p],
    "ugp_maxpool2x2": [_p, _i, _i, _i, _i, _p, _p, _p],
    "ugp_maxpool2x2_bwd": [_p, _p, _i, _i, _i, _i, _p, _p],
    "ugp_upsample2x": [_p, _i, _i, _i, _i, _p, _p],
    "ugp_upsample2x_bwd": [_p, _i, _i, _i, _i, _p, _p],
    "ugp_outconv_fwd": [_p, _p, _p, _i, _i, _i, _i, _p, _p],
    "ugp_outconv_bwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "ugp_head_combine_fwd": [_p, _p, _i, _i, _i, _p, _p],
    "ugp_head_combine_bwd": [_p, _i, _i, _i, _p, _p],
    "ugp_head_combine_fwd_hw": [_p, _p, _p, _i, _i, _i, _i, _p, _p],
    "ugp_head_combine_bwd_hw": [_p, _i, _i, _i, _i, _i, _p, _p],
    "ugp_uncertainty_margin": [_p, _i, _i, _i, _i, _i, _p, _p],
    "ugp_softmax_entropy": [_p, _i, _i, _p, _p],
    "ugp_bce_uncertainty_loss": [_p, _p, _p, _f, _f, _i, _ll, _f, _p, _p, _p],
    "ugp_ce_weighted_loss": [_p, _p, _p, _f, _i, _i, _f, _p, _p, _p],
    "ugp_global_avg_pool": [_p, _i, _i, _i, _p, _p],
    "ugp_global_avg_pool_bwd": [_p, _i, _i, _i, _p, _p],
    "ugp_gap_mlp_head_fwd": [_p, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _p, _i, _p, _f, _f, _f, _p, _p, _p, _p, _p, _p],
    "ugp_mlp_head_bwd": [_p, _i, _i, _p, _i, _p, _i, _p, _i, _p, _f, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "ugp_rmsprop_step": [_p, _p, _p, _ll, _f, _f, _f, _f, _f, _p],
    "ugp_rmsprop_multi": [_p, _p, _p, _p, _i, _f, _f, _f, _f, _f, _p],
    "ugp_adam_multi": [_p, _p, _p, _p, _p, _i, _d, _d, _d, _d, _d, _ll, _p, _f, _p],
    "ugp_counter_add": [_p, _ll, _p],
    "ugp_rmsprop_conv_packed": [_p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _f, _f, _p],
}
INFO_FUNCS = ("ugp_last_error", "ugp_version", "ugp_device_sm_count", "ugp_set_option")
EXPORTED = tuple(SIGNATURES) + INFO_FUNCS

_lib = None


class UgpError(RuntimeError):
    """A C-ABI call returned a negative status."""


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from . import build as _build

        try:
            _build.build()
        except Exception as e:  # no nvcc and no prebuilt library: fail loudly
            raise UgpError(
                f"{LIB_PATH} is missing and could not be built ({e}). "
                "Run `python -m ug_pg_unet_b200.build` where nvcc is available."
            ) from e
    handle = C.CDLL(str(LIB_PATH))
    for name, args in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.argtypes = args
        fn.restype = C.c_int
    handle.ugp_last_error.restype = C.c_char_p
    handle.ugp_last_error.argtypes = []
    handle.ugp_version.restype = C.c_int
    handle.ugp_device_sm_count.restype = C.c_int
    handle.ugp_set_option.restype = C.c_int
    handle.ugp_set_option.argtypes = [C.c_char_p, C.c_int]
    _lib = handle
    # UGP_OPTIONS="name=value,name=value": tuning / what-if switches of ugp_set_option applied at load
    for item in filter(None, os.environ.get("UGP_OPTIONS", "").split(",")):
        name, _, value = item.partition("=")
        check(handle.ugp_set_option(name.strip().encode(), int(value)), f"ugp_set_option({name})")
    return handle


def set_option(name: str, value: int) -> None:
    check(lib().ugp_set_option(name.encode(), int(value)), f"ugp_set_option({name})")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ugp_last_error().decode("utf-8", "replace")
        raise UgpError(f"{what} failed (code {rc}): {msg}")
