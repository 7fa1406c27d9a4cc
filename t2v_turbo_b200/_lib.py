"""ctypes binding of libt2v_b200.so (include/t2v_b200.h).

The library is the product; this file only marshals torch tensors (device pointers, strides, the
current CUDA stream) into the POD descriptors.  There is NO fallback: if the shared object is
missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libt2v_b200.so"

MAX_DIMS = 4
MAX_TAPS = 9
EPI_GEGLU = 1
EPI_OUT_F32 = 2
EPI_GELU = 4
WS_CLEAN = 8      # the split-K workspace is all-zero on entry and is left zeroed: no zero kernel before a split-K GEMM


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p * 2),
        ("a_ch", C.c_int32 * 2),
        ("a_ch_total", C.c_int32 * 2),
        ("a_size", C.c_int64 * MAX_DIMS),
        ("a_stride", (C.c_int64 * MAX_DIMS) * 2),
        ("box", C.c_int32 * MAX_DIMS),
        ("n_taps", C.c_int32),
        ("tap_off", (C.c_int32 * MAX_DIMS) * MAX_TAPS),
        ("tap_ch_off", C.c_int32 * MAX_TAPS),
        ("b", C.c_void_p),
        ("b_rows", C.c_int64),
        ("b_batches", C.c_int64),
        ("b_batch_stride", C.c_int64),
        ("b_row_stride", C.c_int64),
        ("b_batch_dim", C.c_int32),
        ("out", C.c_void_p),
        ("o_size", C.c_int64 * MAX_DIMS),
        ("o_stride", C.c_int64 * MAX_DIMS),
        ("n_out", C.c_int32),
        ("bias", C.c_void_p),
        ("bias_row_stride", C.c_int64),
        ("bias_dim", C.c_int32),
        ("bias_div", C.c_int32),
        ("residual", C.c_void_p),
        ("r_stride", C.c_int64 * MAX_DIMS),
        ("alpha", C.c_float),
        ("flags", C.c_uint32),
        ("block_n", C.c_int32),
        ("split_k", C.c_int32),
        ("tune", C.c_int32),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("row_stats", C.c_void_p),
        ("col_sum", C.c_void_p),
        ("ln_raw", C.c_int32), ("ln_channels", C.c_int32), ("ln_eps", C.c_float),
        ("row_accum", C.c_void_p),
        ("col_accum", C.c_void_p),
        ("cs_mult", C.c_int32 * 4),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("len_q", C.c_int32), ("len_k", C.c_int32),
        ("q_stride_b", C.c_int64), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_b", C.c_int64), ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_b", C.c_int64), ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_b", C.c_int64), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("kv_batch_div", C.c_int32), ("scale", C.c_float), ("causal", C.c_int32),
        ("lse2", C.c_void_p),
    ]


class ShortAttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("n_seq_outer", C.c_int32), ("n_seq_inner", C.c_int32), ("heads", C.c_int32), ("len", C.c_int32),
        ("q_stride_outer", C.c_int64), ("q_stride_inner", C.c_int64), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_outer", C.c_int64), ("k_stride_inner", C.c_int64), ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_outer", C.c_int64), ("v_stride_inner", C.c_int64), ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_outer", C.c_int64), ("o_stride_inner", C.c_int64), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("scale", C.c_float),
        ("probs", C.c_void_p), ("probs_dtype", C.c_int32),
    ]


class GroupNormDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p * 2), ("ch", C.c_int32 * 2),
        ("x_row_stride", C.c_int64 * 2),
        ("out", C.c_void_p), ("out_row_stride", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("rows", C.c_int64), ("rows_per_sample", C.c_int64),
        ("groups", C.c_int32), ("eps", C.c_float), ("silu", C.c_int32),
        ("workspace", C.c_void_p),
        ("mode", C.c_int32),
        ("chan_sums", C.c_void_p * 2),
        ("chan_group", C.c_int32),
    ]


class LayerNormDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_row_stride", C.c_int64),
        ("out", C.c_void_p), ("out_row_stride", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("rows", C.c_int64), ("channels", C.c_int32), ("eps", C.c_float),
    ]


class SmallLinearDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_row_stride", C.c_int64),
        ("w", C.c_void_p), ("bias", C.c_void_p),
        ("add", C.c_void_p), ("add_row_stride", C.c_int64),
        ("out", C.c_void_p), ("out_row_stride", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("silu_in", C.c_int32), ("silu_out", C.c_int32), ("round_bf16", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_ch", C.c_int32),
        ("a_size", C.c_int64 * MAX_DIMS), ("a_stride", C.c_int64 * MAX_DIMS),
        ("b", C.c_void_p), ("b_cols", C.c_int32),
        ("o_size", C.c_int64 * MAX_DIMS), ("b_stride", C.c_int64 * MAX_DIMS),
        ("box", C.c_int32 * MAX_DIMS),
        ("n_taps", C.c_int32),
        ("tap_off", (C.c_int32 * MAX_DIMS) * MAX_TAPS),
        ("out", C.c_void_p),
        ("out_j_stride", C.c_int64), ("out_c_stride", C.c_int64), ("out_tap_stride", C.c_int64),
        ("alpha", C.c_float),
    ]


class GroupNormBwdDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_row_stride", C.c_int64),
        ("dy", C.c_void_p), ("dy_row_stride", C.c_int64),
        ("dx_add", C.c_void_p), ("dx_add_row_stride", C.c_int64),
        ("dx", C.c_void_p), ("dx_row_stride", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("rows", C.c_int64), ("rows_per_sample", C.c_int64),
        ("channels", C.c_int32), ("groups", C.c_int32), ("eps", C.c_float), ("silu", C.c_int32),
        ("workspace", C.c_void_p),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("d_o", C.c_void_p),
        ("lse2", C.c_void_p), ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("len_q", C.c_int32), ("len_k", C.c_int32),
    ] + [(f"{n}_stride_{a}", C.c_int64) for n in ("q", "k", "v", "do", "dq", "dk", "dv") for a in ("b", "t", "h")] + [
        ("kv_batch_div", C.c_int32), ("scale", C.c_float),
    ]


class ShortAttnBwdDesc(C.Structure):
    _fields_ = [("fwd", ShortAttnDesc), ("d_o", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p)]


# every symbol include/t2v_b200.h declares: (name, restype, argtypes)
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = {
    "t2v_version": (C.c_int, []),
    "t2v_last_error": (C.c_char_p, []),
    "t2v_gemm": (C.c_int, [C.POINTER(GemmDesc), _vp]),
    "t2v_attn_fwd": (C.c_int, [C.POINTER(AttnDesc), _vp]),
    "t2v_attn_short_fwd": (C.c_int, [C.POINTER(ShortAttnDesc), _vp]),
    "t2v_groupnorm": (C.c_int, [C.POINTER(GroupNormDesc), _vp]),
    "t2v_layernorm": (C.c_int, [C.POINTER(LayerNormDesc), _vp]),
    "t2v_layernorm_stats": (C.c_int, [C.POINTER(LayerNormDesc), _vp, _vp]),
    "t2v_small_linear": (C.c_int, [C.POINTER(SmallLinearDesc), _vp]),
    "t2v_sinusoidal_embedding": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "t2v_conv3x3_small_cin": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "t2v_bcthw_to_frames": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "t2v_bcthw_to_frames_pad": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "t2v_bcthw_to_frames_mix": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp]),
    "t2v_frames_to_bcthw": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "t2v_upsample_nearest2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "t2v_concat_channels": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    "t2v_softmax_rows": (C.c_int, [_vp, _i64, _i32, _i64, _f32, _vp]),
    "t2v_gaussian_sample": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "t2v_wgrad": (C.c_int, [C.POINTER(WgradDesc), _vp]),
    "t2v_scale_mask": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _vp]),
    "t2v_dropout_scale": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, C.c_uint32, _vp]),
    "t2v_adamw_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    "t2v_sum_squares": (C.c_int, [_vp, _i64, _vp, _vp]),
    "t2v_mse_loss_grad": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "t2v_groupnorm_bwd": (C.c_int, [C.POINTER(GroupNormBwdDesc), _vp]),
    "t2v_layernorm_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _f32, _vp]),
    "t2v_geglu": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "t2v_ew2d": (C.c_int, [_i32, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "t2v_colsum_samples": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "t2v_resample2x": (C.c_int, [_i32, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "t2v_attn_delta": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i32, _i32, _i32, _vp]),
    "t2v_attn_bwd": (C.c_int, [C.POINTER(AttnBwdDesc), _vp]),
    "t2v_attn_short_bwd": (C.c_int, [C.POINTER(ShortAttnBwdDesc), _vp]),
    "t2v_huber_loss_grad": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _vp]),
    "t2v_embedding_gather": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "t2v_video_to_uint8": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "t2v_scale_add_rows": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    "t2v_lcm_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "t2v_gemm_workspace_bytes": (C.c_int64, [C.POINTER(GemmDesc)]),
    "t2v_groupnorm_workspace_bytes": (C.c_int64, [_i64, _i32, _i32]),
    "t2v_groupnorm_affine_grad": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _f32, _i32, _vp]),
    "t2v_layernorm_affine_grad": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _f32, _vp]),
    "t2v_ema_update": (C.c_int, [_vp, _vp, _i64, _f32, _vp]),
    "t2v_attn_short_probs_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "t2v_softmax_bwd_rows": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _vp]),
    "t2v_pack_conv_weight": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _vp]),
    "t2v_pack_geglu_rows": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library; a fresh clone builds it on first use (nvcc, ~2 min).  Fails loudly when it can be
    neither found nor built: there is no CPU / PyTorch fallback for the hot path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            try:
                from . import build as _build
                _build.build()
            except Exception as e:   # no nvcc, compile error, read-only tree ...
                raise RuntimeError(
                    f"{LIB_PATH} is missing and could not be built ({e}); build it with "
                    "`python -m t2v_turbo_b200.build` (there is no CPU / PyTorch fallback for the hot path)") from e
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().t2v_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


DTYPE_CODE = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
