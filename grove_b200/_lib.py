"""ctypes binding of libgrove_b200.so (the C ABI declared in include/grove_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrove_b200.so")


class GemmEpilogue(C.Structure):
    """mirror of `struct grove_gemm_epilogue`"""
    _fields_ = [("bias", C.c_void_p), ("resid", C.c_void_p), ("resid_row_mod", C.c_int), ("gate_alpha", C.c_void_p),
                ("act", C.c_int), ("out_f32", C.c_int), ("out2_bf16", C.c_void_p), ("max_ctas", C.c_int), ("force_ctas", C.c_int),
                ("out2_pre_act", C.c_int), ("dact_pre", C.c_void_p), ("dact", C.c_int), ("splits", C.c_int),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong), ("resid_bf16", C.c_void_p),
                ("ln_stats_out", C.c_void_p), ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_parts", C.c_int), ("ln_eps", C.c_float)]


class TwoWayAParams(C.Structure):
    """mirror of `struct grove_twoway_a_params`"""
    _fields_ = [(n, C.c_void_p) for n in ("wq_t", "bq", "wk_t", "bk", "wv_t", "bv", "wo_t", "bo", "ln_g", "ln_b")] + [("ln_eps", C.c_float)] + \
               [("wq2_t", C.c_void_p), ("bq2", C.c_void_p), ("skip_pe", C.c_int)]


class TwoWayBParams(C.Structure):
    """mirror of `struct grove_twoway_b_params`"""
    _fields_ = [("wo_t", C.c_void_p), ("bo", C.c_void_p), ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p), ("ln2_eps", C.c_float),
                ("w1_t", C.c_void_p), ("b1", C.c_void_p), ("w2_t", C.c_void_p), ("b2", C.c_void_p), ("mlp_dim", C.c_int),
                ("ln3_g", C.c_void_p), ("ln3_b", C.c_void_p), ("ln3_eps", C.c_float),
                ("wk_t", C.c_void_p), ("bk", C.c_void_p), ("wv_t", C.c_void_p), ("bv", C.c_void_p), ("wqf_t", C.c_void_p), ("bqf", C.c_void_p)]


_P, _I, _F, _LL, _D = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double

# name -> argtypes (every function returns int unless listed in _RESTYPES); the stream is always last
SIGNATURES = {
    "grove_abi_version": [],
    "grove_last_error": [],
    "grove_launch_count": [],
    "grove_reset_launch_count": [],
    "grove_add_launch_count": [C.c_longlong],
    "grove_gemm_bf16": [_P, _P, _P, _I, _I, _I, C.POINTER(GemmEpilogue), _P],
    "grove_conv_gemm_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(GemmEpilogue), _P],
    "grove_im2col_patch16": [_P, _P, _I, _I, _I, _I, _P],
    "grove_layernorm": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "grove_layernorm_bf16in": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "grove_attn_window_relpos_tc_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_attn_window_relpos_tc_fwd_lse": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_attn_global_relpos_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_attn_global_relpos_fwd_lse": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_cast_f32_bf16": [_P, _P, _LL, _P],
    "grove_tokens_to_nchw_bf16": [_P, _P, _I, _I, _I, _P],
    "grove_adaptive_avgpool3d_tokens": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "grove_nchw_to_tokens_bf16": [_P, _P, _I, _I, _I, _P],
    "grove_gather_rows_bf16": [_P, _I, _P, _P, _I, _I, _P],
    "grove_dense_pe": [_P, _P, _I, _I, _P],
    "grove_add_rowvec_bf16": [_P, _P, _P, _LL, _I, _P],
    "grove_decoder_t2i_attention": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_decoder_i2t_attention": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_decoder_keys_add_ln": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "grove_small_linear_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_twoway_block_tokens_a_fwd": [_P, _P, C.POINTER(TwoWayAParams), _P, _P, _I, _I, _I, _P],
    "grove_twoway_block_tokens_b_fwd": [_P, _P, _P, C.POINTER(TwoWayBParams), _P, _P, _P, _P, _I, _I, _I, _P],
    "grove_decoder_t2i_attention_wide": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_decoder_heads_fwd": [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_token_self_attention": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_add_layernorm_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "grove_box_postprocess": [_P, _P, _P, _F, _P, _P, _I, _P],
    "grove_box_losses_fwd": [_P, _P, _P, _P, _P, _P, _I, _P],
    "grove_box_iou": [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "grove_greedy_match": [_P, _P, _D, _D, _P, _P, _I, _I, _P],
    "grove_center_in_box": [_P, _P, _P, _I, _P],
    "grove_viou_decisions": [_P, _P, _P, _I, _I, _P, _P, _P, _P],
    "grove_val_metrics": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "grove_resize_rows_u8": [_P, _P, _P, _P, _I, _LL, _I, _I, _P],
    "grove_frames_to_patches_u8": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P],
    # training step (backward pass)
    "grove_conv_wgrad_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "grove_transpose_to_bf16": [_P, _I, _P, _I, _I, _P],
    "grove_transpose_shift3_to_bf16": [_P, _I, _P, _I, _I, _I, _P],
    "grove_reduce_partials_f32": [_P, _I, _LL, _P, _I, _F, _P],
    "grove_layernorm_bwd": [_P, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _LL, _I, _F, _P],
    "grove_adapter_gate_bwd": [_P, _P, _P, _P, _P, _P, _LL, _I, _P],
    "grove_colsum": [_P, _I, _P, _LL, _I, _P],
    "grove_segment_sum_f32": [_P, _P, _P, _I, _LL, _I, _P],
    "grove_small_wgrad_f32": [_P, _P, _P, _I, _I, _I, _P],
    "grove_act_bwd_f32": [_P, _P, _P, _LL, _I, _P],
    "grove_token_self_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_decoder_t2i_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_decoder_i2t_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_batch_sum_bf16": [_P, _P, _I, _LL, _P],
    "grove_attn_relpos_bwd_workspace_bytes": [_I, _I, _I, _I, _I],
    "grove_attn_relpos_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_attn_relpos_bwd_lse": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_box_losses_bwd": [_P, _P, _P, _P, _P, _F, _F, _P, _P, _I, _P],
}
_RESTYPES = {"grove_last_error": C.c_char_p, "grove_launch_count": C.c_longlong, "grove_reset_launch_count": None, "grove_add_launch_count": None,
             "grove_attn_relpos_bwd_workspace_bytes": C.c_longlong}

# test-only cross-check kernels (libgrove_b200_legacy.so, include/grove_b200_legacy.h)
LEGACY_LIB_PATH = os.path.join(_HERE, "libgrove_b200_legacy.so")
LEGACY_SIGNATURES = {
    "grove_attn_window_relpos_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "grove_attn_global_relpos_fwd_mma": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "grove_last_error": [],
}

_lib = None
_legacy = None


def lib_legacy() -> C.CDLL:
    """the test-only library with the round-1 mma.sync forward attention kernels"""
    global _legacy
    if _legacy is None:
        if not os.path.exists(LEGACY_LIB_PATH):
            raise RuntimeError(f"{LEGACY_LIB_PATH} is missing: build it with `python -m grove_b200.build`")
        l = C.CDLL(LEGACY_LIB_PATH)
        for name, args in LEGACY_SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_char_p if name == "grove_last_error" else C.c_int
        _legacy = l
    return _legacy


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m grove_b200.build` "
                               "(or __graft_entry__.build()); grove_b200 has no CPU / PyTorch fallback")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        if l.grove_abi_version() != 5:
            raise RuntimeError("libgrove_b200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().grove_last_error().decode(errors="replace")
        raise RuntimeError(f"grove_b200 {what} failed (code {rc}): {msg}")
