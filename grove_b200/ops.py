"""Tensor-level wrappers around the C ABI (torch here is plumbing: device memory + streams only)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import GemmEpilogue, check, lib

BF16, F32 = torch.bfloat16, torch.float32
ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "sigmoid": 3}


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"grove_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"grove_b200: `{name}` must be contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


def _epilogue(bias, resid, resid_row_mod, gate_alpha, act, out, out2, max_ctas, force_ctas=0):
    e = GemmEpilogue()
    e.bias = None if bias is None else _req(bias, F32, "bias").data_ptr()
    e.resid = None if resid is None else _req(resid, F32, "resid").data_ptr()
    e.resid_row_mod = int(resid_row_mod)
    e.gate_alpha = None if gate_alpha is None else _req(gate_alpha, F32, "gate_alpha").data_ptr()
    e.act = ACT[act]
    e.out_f32 = 1 if out.dtype == F32 else 0
    e.out2_bf16 = None if out2 is None else _req(out2, BF16, "out2").data_ptr()
    e.max_ctas = int(max_ctas)
    e.force_ctas = int(force_ctas)
    return e


def gemm(a, w, out, *, bias=None, resid=None, resid_row_mod=0, gate_alpha=None, act=None, out2=None, max_ctas=0, force_ctas=0):
    """out[M,N] = resid + tanh(gate_alpha) * act(a[M,K] @ w[N,K]^T + bias)   (tcgen05 GEMM)"""
    _req(a, BF16, "a"); _req(w, BF16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.shape == (M, N) and out.is_contiguous() and out.dtype in (BF16, F32)
    e = _epilogue(bias, resid, resid_row_mod, gate_alpha, act, out, out2, max_ctas, force_ctas)
    check(lib().grove_gemm_bf16(_p(a), _p(w), _p(out), M, N, K, C.byref(e), _stream(a)), "grove_gemm_bf16")
    return out


def conv_gemm(x, wp, out, *, V, T, G, kt, bias=None, resid=None, gate_alpha=None, act=None, out2=None, max_ctas=0, force_ctas=0):
    """implicit-GEMM 'same' conv over token-major x[V,T,G,G,C]; wp[N, taps*C] tap-major"""
    _req(x, BF16, "x"); _req(wp, BF16, "wp")
    Cc = x.shape[-1]
    N = wp.shape[0]
    assert x.numel() == V * T * G * G * Cc and wp.shape[1] == 9 * kt * Cc
    assert out.shape == (V * T * G * G, N) and out.is_contiguous()
    e = _epilogue(bias, resid, 0, gate_alpha, act, out, out2, max_ctas, force_ctas)
    check(lib().grove_conv_gemm_bf16(_p(x), _p(wp), _p(out), V, T, G, Cc, N, kt, C.byref(e), _stream(x)), "grove_conv_gemm_bf16")
    return out


def im2col_patch16(images, out):
    _req(images, BF16, "images")
    V, c, T, H, W = images.shape
    assert c == 3 and out.shape == (V * T * (H // 16) * (W // 16), 768)
    check(lib().grove_im2col_patch16(_p(images), _p(_req(out, BF16, "out")), V, T, H, W, _stream(images)), "grove_im2col_patch16")
    return out


def layernorm(x, gamma, beta, out, eps):
    _req(x, F32, "x"); _req(gamma, F32, "gamma"); _req(beta, F32, "beta")
    rows, D = x.shape
    assert out.shape == x.shape and out.is_contiguous() and out.dtype in (BF16, F32)
    check(lib().grove_layernorm(_p(x), _p(gamma), _p(beta), _p(out), 1 if out.dtype == F32 else 0, rows, D, float(eps), _stream(x)),
          "grove_layernorm")
    return out


def attn_window(qkv, qkv_bias_bf16, rel_h, rel_w, out, *, F, G, heads, hd, ws=14):
    for t, n in ((qkv, "qkv"), (qkv_bias_bf16, "qkv_bias"), (rel_h, "rel_pos_h"), (rel_w, "rel_pos_w"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_h.shape == (2 * ws - 1, hd)
    check(lib().grove_attn_window_relpos_fwd(_p(qkv), _p(qkv_bias_bf16), _p(rel_h), _p(rel_w), _p(out), F, G, heads, hd, ws, _stream(qkv)),
          "grove_attn_window_relpos_fwd")
    return out


def window_rel_table(rel_h, rel_w):
    """[64, hd] bf16 table for the tcgen05 window kernel: rows 0..26 rel_pos_h, rows 32..58 rel_pos_w"""
    assert rel_h.shape[0] == 27 and rel_w.shape[0] == 27
    t = torch.zeros(64, rel_h.shape[1], device=rel_h.device, dtype=BF16)
    t[0:27] = rel_h.to(BF16)
    t[32:59] = rel_w.to(BF16)
    return t


def attn_window_tc(qkv, qkv_bias_bf16, rel_table, out, *, F, G, heads, hd, ws=14):
    for t, n in ((qkv, "qkv"), (qkv_bias_bf16, "qkv_bias"), (rel_table, "rel_table"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_table.shape == (64, hd)
    check(lib().grove_attn_window_relpos_tc_fwd(_p(qkv), _p(qkv_bias_bf16), _p(rel_table), _p(out), F, G, heads, hd, ws, _stream(qkv)),
          "grove_attn_window_relpos_tc_fwd")
    return out


def attn_global(qkv, rel_h, rel_w, out, *, F, G, heads, hd, legacy_mma=False):
    for t, n in ((qkv, "qkv"), (rel_h, "rel_pos_h"), (rel_w, "rel_pos_w"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_h.shape == (2 * G - 1, hd)
    fn = lib().grove_attn_global_relpos_fwd_mma if legacy_mma else lib().grove_attn_global_relpos_fwd
    check(fn(_p(qkv), _p(rel_h), _p(rel_w), _p(out), F, G, heads, hd, _stream(qkv)), "grove_attn_global_relpos_fwd")
    return out


def cast_f32_bf16(x, out):
    check(lib().grove_cast_f32_bf16(_p(_req(x, F32, "x")), _p(_req(out, BF16, "out")), x.numel(), _stream(x)), "grove_cast_f32_bf16")
    return out


def tokens_to_nchw(tok, out, F, N, Cc):
    check(lib().grove_tokens_to_nchw_bf16(_p(_req(tok, BF16, "tok")), _p(_req(out, BF16, "out")), F, N, Cc, _stream(tok)), "tokens_to_nchw")
    return out


def nchw_to_tokens(x, out, F, N, Cc):
    check(lib().grove_nchw_to_tokens_bf16(_p(_req(x, BF16, "x")), _p(_req(out, BF16, "out")), F, N, Cc, _stream(x)), "nchw_to_tokens")
    return out


def gather_rows_bf16(src, idx, out):
    assert src.is_cuda and src.is_contiguous() and src.dtype in (BF16, F32) and idx.dtype == torch.int32
    n, D = idx.numel(), src.shape[-1]
    if n:
        check(lib().grove_gather_rows_bf16(_p(src), 1 if src.dtype == F32 else 0, _p(idx), _p(_req(out, BF16, "out")), n, D, _stream(src)),
              "grove_gather_rows_bf16")
    return out


def dense_pe(gauss, G):
    _req(gauss, F32, "gauss")
    F2 = gauss.shape[1]
    pe = torch.empty(G * G, 2 * F2, device=gauss.device, dtype=F32)
    check(lib().grove_dense_pe(_p(gauss), _p(pe), G, F2, _stream(gauss)), "grove_dense_pe")
    return pe


def add_rowvec_bf16(x, vec, out):
    rows, Cc = x.shape
    check(lib().grove_add_rowvec_bf16(_p(_req(x, BF16, "x")), _p(_req(vec, F32, "vec")), _p(_req(out, BF16, "out")), rows, Cc, _stream(x)),
          "grove_add_rowvec_bf16")
    return out


def t2i_attention(q, k, v, src_of, B, T, N, heads, dh):
    out = torch.empty(B, T, heads * dh, device=q.device, dtype=F32)
    check(lib().grove_decoder_t2i_attention(_p(_req(q, F32, "q")), _p(_req(k, BF16, "k")), _p(_req(v, BF16, "v")), _p(src_of), _p(out),
                                            B, T, N, heads, dh, _stream(q)), "grove_decoder_t2i_attention")
    return out


def i2t_attention(qi, kt, vt, src_of, out, B, T, N, heads, dh):
    check(lib().grove_decoder_i2t_attention(_p(_req(qi, BF16, "qi")), _p(_req(kt, F32, "kt")), _p(_req(vt, F32, "vt")), _p(src_of),
                                            _p(_req(out, BF16, "out")), B, T, N, heads, dh, _stream(qi)), "grove_decoder_i2t_attention")
    return out


def keys_add_ln(keys_in, src_of, delta, g, b, out, B, N, Cc, eps=1e-5):
    check(lib().grove_decoder_keys_add_ln(_p(_req(keys_in, BF16, "keys")), _p(src_of), _p(_req(delta, F32, "delta")), _p(_req(g, F32, "g")),
                                          _p(_req(b, F32, "b")), _p(_req(out, BF16, "out")), B, N, Cc, float(eps), _stream(delta)),
          "grove_decoder_keys_add_ln")
    return out


def small_linear(x, w, b=None, *, act=None, resid=None):
    _req(x, F32, "x"); _req(w, F32, "w")
    R, K = x.shape
    N = w.shape[0]
    y = torch.empty(R, N, device=x.device, dtype=F32)
    if R:
        check(lib().grove_small_linear_f32(_p(x), _p(w), _p(b), _p(resid), _p(y), R, N, K, ACT[act], _stream(x)), "grove_small_linear_f32")
    return y


def token_self_attention(q, k, v, B, T, heads, dh):
    out = torch.empty_like(q)
    check(lib().grove_token_self_attention(_p(_req(q, F32, "q")), _p(_req(k, F32, "k")), _p(_req(v, F32, "v")), _p(out), B, T, heads, dh,
                                           _stream(q)), "grove_token_self_attention")
    return out


def add_layernorm(x, r, g, b, *, eps=1e-5, add2=None):
    _req(x, F32, "x")
    R, Cc = x.shape
    y = torch.empty_like(x)
    y2 = torch.empty_like(x) if add2 is not None else None
    check(lib().grove_add_layernorm_f32(_p(x), _p(r), _p(_req(g, F32, "g")), _p(_req(b, F32, "b")), _p(y), _p(add2), _p(y2), R, Cc, float(eps),
                                        _stream(x)), "grove_add_layernorm_f32")
    return (y, y2) if add2 is not None else y


def box_postprocess(boxes, logits, size_wh, thr):
    B = boxes.shape[0]
    xyxy = torch.empty(B, 4, device=boxes.device, dtype=F32)
    keep = torch.empty(B, device=boxes.device, dtype=torch.uint8)
    if B:
        check(lib().grove_box_postprocess(_p(_req(boxes, F32, "boxes")), _p(_req(logits, F32, "logits")), _p(_req(size_wh, F32, "size_wh")),
                                          float(thr), _p(xyxy), _p(keep), B, _stream(boxes)), "grove_box_postprocess")
    return xyxy, keep


def box_losses(boxes, logits, gt, sel, labels):
    B = boxes.shape[0]
    sums = torch.empty(3, device=boxes.device, dtype=F32)
    check(lib().grove_box_losses_fwd(_p(_req(boxes, F32, "boxes")), _p(_req(logits, F32, "logits")), _p(_req(gt, F32, "gt")),
                                     _p(_req(sel, torch.uint8, "sel")), _p(_req(labels, F32, "labels")), _p(sums), B, _stream(boxes)),
          "grove_box_losses_fwd")
    return sums


def box_iou(a, b, mode, frm_mask=None):
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype and a.dtype in (F32, torch.float64) and a.is_contiguous() and b.is_contiguous()
    n, m = a.shape[0], b.shape[0]
    out = torch.empty(n, m, device=a.device, dtype=a.dtype)
    if frm_mask is not None:
        _req(frm_mask, torch.uint8, "frm_mask")
    check(lib().grove_box_iou(_p(a), a.shape[1], _p(b), b.shape[1], _p(frm_mask), _p(out), n, m, mode, 1 if a.dtype == torch.float64 else 0,
                              _stream(a)), "grove_box_iou")
    return out


def greedy_match(iou, sim, iou_thr, sim_thr):
    n, m = iou.shape
    iou = iou.clone().contiguous(); sim = sim.clone().contiguous()
    pairs = torch.zeros(max(min(n, m), 1), 2, device=iou.device, dtype=torch.int32)
    count = torch.zeros(1, device=iou.device, dtype=torch.int32)
    check(lib().grove_greedy_match(_p(_req(iou, torch.float64, "iou")), _p(_req(sim, torch.float64, "sim")), float(iou_thr), float(sim_thr),
                                   _p(pairs), _p(count), n, m, _stream(iou)), "grove_greedy_match")
    c = int(count.item())
    return [tuple(int(v) for v in p) for p in pairs[:c].tolist()]


def launch_count() -> int:
    return int(lib().grove_launch_count())


def reset_launch_count() -> None:
    lib().grove_reset_launch_count()
