"""Tensor-level wrappers around the C ABI (torch here is plumbing: device memory + streams only)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import GemmEpilogue, TwoWayAParams, TwoWayBParams, check, lib, lib_legacy

BF16, F32 = torch.bfloat16, torch.float32
ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "sigmoid": 3}


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    """the launching stream = torch's current stream on the tensor's device.  The C ABI launches on the CURRENT device, so a tensor that
    lives elsewhere is a caller error (the nn.Module entry points switch devices themselves, see `device_of`)."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"grove_b200: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           "wrap the call in `with torch.cuda.device(tensor.device):`")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def device_of(t: torch.Tensor):
    """context manager making `t`'s device current for the kernels launched inside (multi-GPU processes)"""
    if not t.is_cuda:
        raise RuntimeError("grove_b200 runs on CUDA tensors only (there is no CPU path)")
    return torch.cuda.device(t.device)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"grove_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"grove_b200: `{name}` must be contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


def _epilogue(bias, resid, resid_row_mod, gate_alpha, act, out, out2, max_ctas, force_ctas=0, out2_pre_act=0, dact_pre=None, dact=None, splits=1,
              ln_stats_out=None, ln_fold=None):
    e = GemmEpilogue()
    e.ln_stats_out = None if ln_stats_out is None else _req(ln_stats_out, F32, "ln_stats_out").data_ptr()
    if ln_fold is not None:        # (row statistics [M, parts, 2], column sums of gamma*W [N], eps)
        st, cs, eps = ln_fold
        e.ln_stats, e.ln_colsum = _req(st, F32, "ln_stats").data_ptr(), _req(cs, F32, "ln_colsum").data_ptr()
        e.ln_parts, e.ln_eps = int(st.shape[1]), float(eps)
    else:
        e.ln_stats, e.ln_colsum, e.ln_parts, e.ln_eps = None, None, 0, 0.0
    e.out2_pre_act = int(out2_pre_act)
    e.dact_pre = None if dact_pre is None else _req(dact_pre, BF16, "dact_pre").data_ptr()
    e.dact = ACT[dact] if dact_pre is not None else 0
    e.splits = int(splits)
    e.bias = None if bias is None else _req(bias, F32, "bias").data_ptr()
    if resid is not None and resid.dtype == BF16:      # bf16 residual stream (bf16 `out`, may alias it)
        assert out.dtype == BF16 and resid_row_mod == 0 and out2 is None
        e.resid, e.resid_bf16 = None, _req(resid, BF16, "resid").data_ptr()
    else:
        e.resid = None if resid is None else _req(resid, F32, "resid").data_ptr()
        e.resid_bf16 = None
    e.resid_row_mod = int(resid_row_mod)
    e.gate_alpha = None if gate_alpha is None else _req(gate_alpha, F32, "gate_alpha").data_ptr()
    e.act = ACT[act]
    e.out_f32 = 1 if out.dtype == F32 else 0
    e.out2_bf16 = None if out2 is None else _req(out2, BF16, "out2").data_ptr()
    e.max_ctas = int(max_ctas)
    e.force_ctas = int(force_ctas)
    if splits == 1:      # scratch for the wave-quantisation tail split (residual-stream GEMMs) and the few-tiles / long-K split
        ws = _gemm_workspace(out.device)
        e.workspace, e.workspace_bytes = ws.data_ptr(), ws.numel()
    return e


_WORKSPACES = {}


def _gemm_workspace(device):
    ws = _WORKSPACES.get(device)
    if ws is None:
        ws = _WORKSPACES[device] = torch.empty(32 << 20, device=device, dtype=torch.uint8)
    return ws


def gemm(a, w, out, *, bias=None, resid=None, resid_row_mod=0, gate_alpha=None, act=None, out2=None, max_ctas=0, force_ctas=0,
         out2_pre_act=0, dact_pre=None, dact=None, splits=1, ln_stats_out=None, ln_fold=None):
    """out[M,N] = resid + tanh(gate_alpha) * act(a[M,K] @ w[N,K]^T + bias) * dact'(dact_pre)   (tcgen05 GEMM).
    splits > 1: out is fp32 [splits, M, N] raw partial sums (finish with reduce_partials).
    ln_stats_out (with a bf16 resid): fp32 [M, N/128, 2] receives per-row partial (sum, sum of squares) of the output.
    ln_fold = (stats, colsum, eps): nn.LayerNorm of the rows of `a` folded into this GEMM (see grove_gemm_epilogue.ln_stats)."""
    _req(a, BF16, "a"); _req(w, BF16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.is_contiguous() and out.dtype in (BF16, F32)
    assert out.shape == ((M, N) if splits == 1 else (splits, M, N))
    if ln_stats_out is not None:
        assert ln_stats_out.shape == (M, N // 128, 2)
    if ln_fold is not None:
        assert ln_fold[0].shape[0] == M and ln_fold[0].shape[2] == 2 and ln_fold[1].shape == (N,)
    e = _epilogue(bias, resid, resid_row_mod, gate_alpha, act, out, out2, max_ctas, force_ctas, out2_pre_act, dact_pre, dact, splits,
                  ln_stats_out, ln_fold)
    check(lib().grove_gemm_bf16(_p(a), _p(w), _p(out), M, N, K, C.byref(e), _stream(a)), "grove_gemm_bf16")
    return out


def conv_gemm(x, wp, out, *, V, T, G, kt, bias=None, resid=None, gate_alpha=None, act=None, out2=None, max_ctas=0, force_ctas=0,
              out2_pre_act=0):
    """implicit-GEMM 'same' conv over token-major x[V,T,G,G,C]; wp[N, taps*C] tap-major"""
    _req(x, BF16, "x"); _req(wp, BF16, "wp")
    Cc = x.shape[-1]
    N = wp.shape[0]
    assert x.numel() == V * T * G * G * Cc and wp.shape[1] == 9 * kt * Cc
    assert out.shape == (V * T * G * G, N) and out.is_contiguous()
    e = _epilogue(bias, resid, 0, gate_alpha, act, out, out2, max_ctas, force_ctas, out2_pre_act)
    check(lib().grove_conv_gemm_bf16(_p(x), _p(wp), _p(out), V, T, G, Cc, N, kt, C.byref(e), _stream(x)), "grove_conv_gemm_bf16")
    return out


def im2col_patch16(images, out):
    _req(images, BF16, "images")
    V, c, T, H, W = images.shape
    assert c == 3 and out.shape == (V * T * (H // 16) * (W // 16), 768)
    check(lib().grove_im2col_patch16(_p(images), _p(_req(out, BF16, "out")), V, T, H, W, _stream(images)), "grove_im2col_patch16")
    return out


def layernorm(x, gamma, beta, out, eps):
    """out = LayerNorm(x) over the last dim; x fp32 or bf16 (bf16 residual stream), statistics in fp32"""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (BF16, F32)
    _req(gamma, F32, "gamma"); _req(beta, F32, "beta")
    rows, D = x.shape
    assert out.shape == x.shape and out.is_contiguous() and out.dtype in (BF16, F32)
    fn = lib().grove_layernorm if x.dtype == F32 else lib().grove_layernorm_bf16in
    check(fn(_p(x), _p(gamma), _p(beta), _p(out), 1 if out.dtype == F32 else 0, rows, D, float(eps), _stream(x)), "grove_layernorm")
    return out


def attn_window(qkv, qkv_bias_bf16, rel_h, rel_w, out, *, F, G, heads, hd, ws=14):
    """TEST-ONLY: the round-1 mma.sync window attention (libgrove_b200_legacy.so), cross-check of attn_window_tc"""
    for t, n in ((qkv, "qkv"), (qkv_bias_bf16, "qkv_bias"), (rel_h, "rel_pos_h"), (rel_w, "rel_pos_w"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_h.shape == (2 * ws - 1, hd)
    rc = lib_legacy().grove_attn_window_relpos_fwd(_p(qkv), _p(qkv_bias_bf16), _p(rel_h), _p(rel_w), _p(out), F, G, heads, hd, ws, _stream(qkv))
    if rc:
        raise RuntimeError(f"grove_attn_window_relpos_fwd (legacy) failed: {lib_legacy().grove_last_error().decode(errors='replace')}")
    return out


def window_rel_table(rel_h, rel_w):
    """[64, hd] bf16 table for the tcgen05 window kernel: rows 0..26 rel_pos_h, rows 32..58 rel_pos_w"""
    assert rel_h.shape[0] == 27 and rel_w.shape[0] == 27
    t = torch.zeros(64, rel_h.shape[1], device=rel_h.device, dtype=BF16)
    t[0:27] = rel_h.to(BF16)
    t[32:59] = rel_w.to(BF16)
    return t


def attn_window_tc(qkv, qkv_bias_bf16, rel_table, out, *, F, G, heads, hd, ws=14, lse=None):
    for t, n in ((qkv, "qkv"), (qkv_bias_bf16, "qkv_bias"), (rel_table, "rel_table"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_table.shape == (64, hd)
    if lse is not None:
        assert lse.numel() == F * G * G * heads
        check(lib().grove_attn_window_relpos_tc_fwd_lse(_p(qkv), _p(qkv_bias_bf16), _p(rel_table), _p(out), _p(_req(lse, F32, "lse")), F, G, heads, hd,
                                                        ws, _stream(qkv)), "grove_attn_window_relpos_tc_fwd_lse")
        return out
    check(lib().grove_attn_window_relpos_tc_fwd(_p(qkv), _p(qkv_bias_bf16), _p(rel_table), _p(out), F, G, heads, hd, ws, _stream(qkv)),
          "grove_attn_window_relpos_tc_fwd")
    return out


def attn_global(qkv, rel_h, rel_w, out, *, F, G, heads, hd, legacy_mma=False, lse=None):
    for t, n in ((qkv, "qkv"), (rel_h, "rel_pos_h"), (rel_w, "rel_pos_w"), (out, "out")):
        _req(t, BF16, n)
    assert qkv.numel() == F * G * G * 3 * heads * hd and rel_h.shape == (2 * G - 1, hd)
    if lse is not None:
        assert not legacy_mma and lse.numel() == F * G * G * heads
        check(lib().grove_attn_global_relpos_fwd_lse(_p(qkv), _p(rel_h), _p(rel_w), _p(out), _p(_req(lse, F32, "lse")), F, G, heads, hd,
                                                     _stream(qkv)), "grove_attn_global_relpos_fwd_lse")
        return out
    if legacy_mma:      # test-only cross-check kernel
        rc = lib_legacy().grove_attn_global_relpos_fwd_mma(_p(qkv), _p(rel_h), _p(rel_w), _p(out), F, G, heads, hd, _stream(qkv))
        if rc:
            raise RuntimeError(f"grove_attn_global_relpos_fwd_mma (legacy) failed: {lib_legacy().grove_last_error().decode(errors='replace')}")
        return out
    check(lib().grove_attn_global_relpos_fwd(_p(qkv), _p(rel_h), _p(rel_w), _p(out), F, G, heads, hd, _stream(qkv)), "grove_attn_global_relpos_fwd")
    return out


def cast_f32_bf16(x, out):
    check(lib().grove_cast_f32_bf16(_p(_req(x, F32, "x")), _p(_req(out, BF16, "out")), x.numel(), _stream(x)), "grove_cast_f32_bf16")
    return out


def adaptive_avgpool3d_tokens(x, out, *, B, T, H, W, OT, OH, OW):
    """x [(B*T), H*W, C] -> out [B, OT*OH*OW, C] (bf16 or fp32, same dtype): pooling.py:6-25 with the rearranges folded in"""
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype and x.dtype in (BF16, F32)
    Cc = x.shape[-1]
    assert x.numel() == B * T * H * W * Cc and out.numel() == B * OT * OH * OW * Cc
    check(lib().grove_adaptive_avgpool3d_tokens(_p(x), _p(out), 1 if x.dtype == F32 else 0, B, T, H, W, Cc, OT, OH, OW, _stream(x)),
          "grove_adaptive_avgpool3d_tokens")
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


def t2i_attention(q, k, v, src_of, B, T, N, heads, dh, lse=None):
    out = torch.empty(B, T, heads * dh, device=q.device, dtype=F32)
    check(lib().grove_decoder_t2i_attention(_p(_req(q, F32, "q")), _p(_req(k, BF16, "k")), _p(_req(v, BF16, "v")), _p(src_of), _p(out),
                                            _p(lse), B, T, N, heads, dh, _stream(q)), "grove_decoder_t2i_attention")
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


def _struct_of(cls, fields: dict):
    """ctypes parameter block from {field: fp32 CUDA tensor | number | None}; returns (struct, tensors kept alive)"""
    st, keep = cls(), []
    for k, v in fields.items():
        if isinstance(v, torch.Tensor):
            _req(v, F32, k)
            keep.append(v)
            setattr(st, k, v.data_ptr())
        elif v is None:
            setattr(st, k, None)
        else:
            setattr(st, k, v)
    return st, keep


def twoway_tokens_a(queries, tokens, params: dict):
    """part A of a two-way block on the token side (see grove_twoway_block_tokens_a_fwd): -> (queries_out [B,6,256], qt [B,6,128])"""
    B, T, Cc = queries.shape
    _req(queries, F32, "queries"); _req(tokens, F32, "tokens")
    st, keep = _struct_of(TwoWayAParams, params)
    q_out = torch.empty_like(queries)
    qt = torch.empty(B, T, params["wq2_t"].shape[1], device=queries.device, dtype=F32)
    check(lib().grove_twoway_block_tokens_a_fwd(_p(queries), _p(tokens), C.byref(st), _p(q_out), _p(qt), B, T, Cc, _stream(queries)),
          "grove_twoway_block_tokens_a_fwd")
    return q_out, qt


def twoway_tokens_b(queries, att, tokens, params: dict, want_qf: bool):
    """part B (see grove_twoway_block_tokens_b_fwd): -> (queries_out [B,6,256], kt, vt [B,6,128], qf [B,6,128] | None)"""
    B, T, Cc = queries.shape
    _req(queries, F32, "queries"); _req(att, F32, "att"); _req(tokens, F32, "tokens")
    st, keep = _struct_of(TwoWayBParams, params)
    CI = params["wk_t"].shape[1]
    q_out = torch.empty_like(queries)
    kt = torch.empty(B, T, CI, device=queries.device, dtype=F32)
    vt = torch.empty_like(kt)
    qf = torch.empty_like(kt) if want_qf else None
    check(lib().grove_twoway_block_tokens_b_fwd(_p(queries), _p(att), _p(tokens), C.byref(st), _p(q_out), _p(kt), _p(vt), _p(qf), B, T, Cc,
                                                _stream(queries)), "grove_twoway_block_tokens_b_fwd")
    return q_out, kt, vt, qf


def t2i_attention_wide(q, k, v, src_of, B, T, N, heads, dh, lse=None):
    out = torch.empty(B, T, heads * dh, device=q.device, dtype=F32)
    check(lib().grove_decoder_t2i_attention_wide(_p(_req(q, F32, "q")), _p(_req(k, BF16, "k")), _p(_req(v, BF16, "v")), _p(src_of), _p(out),
                                                 _p(lse), B, T, N, heads, dh, _stream(q)), "grove_decoder_t2i_attention_wide")
    return out


def decoder_heads(queries, att, wo, bo, ln_g, ln_b, eps, w0, b0, w2, b2, wt, bt, records, *, tok, hs_out=None):
    """records[B,5] = (box cxcywh, objectness logit) of the prompt token: final out-proj + LayerNorm + both heads in one launch"""
    B, T, Cc = queries.shape
    CI = att.shape[-1]
    assert att.shape[:2] == (B, T) and records.shape == (B, 5)
    for t, n in ((queries, "queries"), (att, "att"), (wo, "wo"), (bo, "bo"), (ln_g, "ln_g"), (ln_b, "ln_b"), (w0, "w0"), (b0, "b0"), (w2, "w2"),
                 (b2, "b2"), (records, "records")):
        _req(t, F32, n)
    check(lib().grove_decoder_heads_fwd(_p(queries), _p(att), _p(wo), _p(bo), _p(ln_g), _p(ln_b), float(eps), _p(w0), _p(b0), _p(w2), _p(b2),
                                        _p(wt), _p(bt), _p(records), _p(hs_out), B, T, int(tok), Cc, CI, _stream(queries)), "grove_decoder_heads_fwd")
    return records


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


def center_in_box(pred, gt):
    n = pred.shape[0]
    out = torch.empty(n, device=pred.device, dtype=torch.uint8)
    check(lib().grove_center_in_box(_p(_req(pred, torch.float64, "pred")), _p(_req(gt, torch.float64, "gt")), _p(out), n, _stream(pred)),
          "grove_center_in_box")
    return out


def viou_decisions(pred, gt, thr):
    n, k = pred.shape[0], thr.numel()
    ious = torch.empty(max(n, 1), device=pred.device, dtype=torch.float64)
    viou = torch.empty(1, device=pred.device, dtype=torch.float64)
    over = torch.empty(max(k, 1), device=pred.device, dtype=torch.uint8)
    check(lib().grove_viou_decisions(_p(_req(pred, torch.float64, "pred")), _p(_req(gt, torch.float64, "gt")), _p(_req(thr, torch.float64, "thr")),
                                     n, k, _p(ious), _p(viou), _p(over), _stream(thr)), "grove_viou_decisions")
    return ious[:n], viou, over[:k]


def val_metrics(boxes, logits, gt, sel, labels):
    B = boxes.shape[0]
    giou = torch.empty(1, device=boxes.device, dtype=torch.float64)
    acc = torch.empty(1, device=boxes.device, dtype=torch.int32)
    check(lib().grove_val_metrics(_p(_req(boxes, F32, "boxes")), _p(_req(logits, F32, "logits")), _p(_req(gt, F32, "gt")),
                                  _p(_req(sel, torch.uint8, "sel")), _p(_req(labels, torch.int32, "labels")), _p(giou), _p(acc), None, B,
                                  _stream(boxes)), "grove_val_metrics")
    return float(giou.item()), int(acc.item())


# ------------------------------------------------------------------ training step (backward pass)
def transpose_to_bf16(x, out=None):
    """[R,C] fp32|bf16 -> [C,R] bf16"""
    assert x.is_cuda and x.is_contiguous() and x.dim() == 2 and x.dtype in (BF16, F32)
    R, Cc = x.shape
    if out is None:
        out = torch.empty(Cc, R, device=x.device, dtype=BF16)
    check(lib().grove_transpose_to_bf16(_p(x), 1 if x.dtype == F32 else 0, _p(_req(out, BF16, "out")), R, Cc, _stream(x)), "grove_transpose_to_bf16")
    return out


def reduce_partials(partials, out, accumulate=False, scale=1.0):
    S = partials.shape[0]
    n = partials[0].numel()
    assert out.numel() == n and out.is_contiguous()
    check(lib().grove_reduce_partials_f32(_p(_req(partials, F32, "partials")), S, n, _p(_req(out, F32, "out")), 1 if accumulate else 0, float(scale),
                                          _stream(out)), "grove_reduce_partials_f32")
    return out


def _even_splits(splits, kblocks):
    """the number of non-empty split-K planes the kernel writes for this request (gemm_tcgen05.cu: dispatch_gemm)"""
    per = (kblocks + splits - 1) // splits
    return (kblocks + per - 1) // per


def wgrad(dy, x, out, accumulate=True):
    """out[N,K] (fp32) (+)= dy[M,N]^T @ x[M,K] on the tcgen05 GEMM (K runs over the M rows; split-K when the output has few tiles).
    dy, x: bf16 (or fp32, cast while transposing).  Requires K % 128 == 0, M % 8 == 0."""
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and out.shape == (N, K) and out.dtype == F32
    dyt = transpose_to_bf16(dy)
    xt = transpose_to_bf16(x)
    tiles = ((N + 255) // 256) * max(K // 256, 1)
    kblocks = (M + 63) // 64
    splits = _even_splits(max(1, min(64, (74 + tiles - 1) // tiles, kblocks // 4)), kblocks)
    part = torch.empty(splits, N, K, device=dy.device, dtype=F32)
    gemm(dyt, xt, part if splits > 1 else part[0], splits=splits)
    return reduce_partials(part, out, accumulate=accumulate)


def conv_wgrad(dy, x, out, *, V, T, G, kt, accumulate=True):
    """out[N, taps*C] fp32 (+)= weight gradient of conv_gemm; dy [tokens, N], x [tokens, C] token-major bf16/fp32"""
    tokens, N = dy.shape
    Cc = x.shape[1]
    taps = 9 * kt
    assert x.shape[0] == tokens == V * T * G * G and out.shape == (N, taps * Cc) and out.dtype == F32
    dyt = transpose_to_bf16(dy)
    assert x.is_contiguous() and x.dtype in (BF16, F32)
    xt = torch.empty(3, Cc, tokens, device=x.device, dtype=BF16)
    check(lib().grove_transpose_shift3_to_bf16(_p(x), 1 if x.dtype == F32 else 0, _p(xt), tokens, Cc, G, _stream(x)), "grove_transpose_shift3_to_bf16")
    tiles = ((N + 255) // 256) * (taps * Cc // 256)
    splits = _even_splits(max(1, min(16, 74 // max(tiles, 1), tokens // 64 // 8)), tokens // 64)
    part = torch.empty(splits, N, taps * Cc, device=dy.device, dtype=F32)
    check(lib().grove_conv_wgrad_bf16(_p(dyt), _p(xt), _p(part), V, T, G, Cc, N, kt, splits, _stream(dy)), "grove_conv_wgrad_bf16")
    return reduce_partials(part, out, accumulate=accumulate)


def layernorm_bwd(x, gamma, dy, *, eps, r=None, dx_in=None, dx_out=None, dx_bf16=None, dgamma=None, dbeta=None, keys_src_of=None, keys_N=0):
    """dx_out = (dx_in or 0) + dLN(dy) for y = LN(x (+ r)); x fp32 [rows,D], or bf16 keys gathered through keys_src_of (then r = fp32 delta)."""
    keys = x.dtype == BF16
    rows, D = dy.shape
    assert dy.is_contiguous() and dy.dtype in (BF16, F32)
    check(lib().grove_layernorm_bwd(_p(x), _p(r), _p(keys_src_of), int(keys_N), 1 if keys else 0, _p(_req(gamma, F32, "gamma")), _p(dy),
                                    1 if dy.dtype == F32 else 0, _p(dx_in), _p(dx_out), _p(dx_bf16), _p(dgamma), _p(dbeta), rows, D, float(eps),
                                    _stream(dy)), "grove_layernorm_bwd")


def adapter_gate_bwd(dy, relu_out, alpha, dyc, dbias, dalpha):
    rows, D = dy.shape
    check(lib().grove_adapter_gate_bwd(_p(_req(dy, F32, "dy")), _p(_req(relu_out, BF16, "relu_out")), _p(_req(alpha, F32, "alpha")),
                                       _p(_req(dyc, BF16, "dyc")), _p(_req(dbias, F32, "dbias")), _p(_req(dalpha, F32, "dalpha")), rows, D,
                                       _stream(dy)), "grove_adapter_gate_bwd")


def colsum(x, out):
    R, Cc = x.shape
    assert x.is_contiguous() and x.dtype in (BF16, F32) and out.numel() == Cc
    check(lib().grove_colsum(_p(x), 1 if x.dtype == F32 else 0, _p(_req(out, F32, "out")), R, Cc, _stream(x)), "grove_colsum")
    return out


def segment_sum(x, offsets, out, accumulate=False):
    segs = offsets.numel() - 1
    n = x[0].numel()
    check(lib().grove_segment_sum_f32(_p(_req(x, F32, "x")), _p(offsets), _p(_req(out, F32, "out")), segs, n, 1 if accumulate else 0, _stream(x)),
          "grove_segment_sum_f32")
    return out


def small_wgrad(dy, x, dw):
    R, N = dy.shape
    K = x.shape[1]
    assert dw.shape == (N, K)
    check(lib().grove_small_wgrad_f32(_p(_req(dy, F32, "dy")), _p(_req(x, F32, "x")), _p(_req(dw, F32, "dw")), R, N, K, _stream(dy)),
          "grove_small_wgrad_f32")
    return dw


def act_bwd(dy, y, kind):
    dx = torch.empty_like(dy)
    check(lib().grove_act_bwd_f32(_p(_req(dy, F32, "dy")), _p(_req(y, F32, "y")), _p(dx), dy.numel(), ACT[kind], _stream(dy)), "grove_act_bwd_f32")
    return dx


def token_self_attention_bwd(q, k, v, dout, B, T, heads, dh):
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    check(lib().grove_token_self_attention_bwd(_p(_req(q, F32, "q")), _p(_req(k, F32, "k")), _p(_req(v, F32, "v")), _p(_req(dout, F32, "dout")),
                                               _p(dq), _p(dk), _p(dv), B, T, heads, dh, _stream(q)), "grove_token_self_attention_bwd")
    return dq, dk, dv


def t2i_attention_bwd(q, k, v, src_of, att, datt, lse, B, T, N, heads, dh):
    dq = torch.empty(B, T, heads * dh, device=q.device, dtype=F32)
    dk = torch.empty(B * N, heads * dh, device=q.device, dtype=BF16)
    dv = torch.empty(B * N, heads * dh, device=q.device, dtype=BF16)
    check(lib().grove_decoder_t2i_attention_bwd(_p(_req(q, F32, "q")), _p(_req(k, BF16, "k")), _p(_req(v, BF16, "v")), _p(src_of),
                                                _p(_req(att, F32, "att")), _p(_req(datt, F32, "datt")), _p(_req(lse, F32, "lse")), _p(dq), _p(dk),
                                                _p(dv), B, T, N, heads, dh, _stream(q)), "grove_decoder_t2i_attention_bwd")
    return dq, dk, dv


def i2t_attention_bwd(qi, kt, vt, src_of, dout, B, T, N, heads, dh):
    dqi = torch.empty(B * N, heads * dh, device=qi.device, dtype=BF16)
    dkt = torch.zeros(B, T, heads * dh, device=qi.device, dtype=F32)
    dvt = torch.zeros(B, T, heads * dh, device=qi.device, dtype=F32)
    check(lib().grove_decoder_i2t_attention_bwd(_p(_req(qi, BF16, "qi")), _p(_req(kt, F32, "kt")), _p(_req(vt, F32, "vt")), _p(src_of),
                                                _p(_req(dout, BF16, "dout")), _p(dqi), _p(dkt), _p(dvt), B, T, N, heads, dh, _stream(qi)),
          "grove_decoder_i2t_attention_bwd")
    return dqi, dkt, dvt


def batch_sum_bf16(x, B):
    n = x.numel() // B
    out = torch.empty(n, device=x.device, dtype=F32)
    check(lib().grove_batch_sum_bf16(_p(_req(x, BF16, "x")), _p(out), B, n, _stream(x)), "grove_batch_sum_bf16")
    return out


def attn_relpos_bwd(qkv, qkv_bias_bf16, rel_h, rel_w, att, datt, dqkv, *, F, G, heads, hd, ws=0, lse=None):
    """d(qkv) of the rel-pos attention (ws = 14 windowed on the unpartitioned tensors, ws = 0 global)"""
    for t, n in ((qkv, "qkv"), (rel_h, "rel_pos_h"), (rel_w, "rel_pos_w"), (att, "att"), (datt, "datt"), (dqkv, "dqkv")):
        _req(t, BF16, n)
    S = ws if ws > 0 else G
    assert rel_h.shape == (2 * S - 1, hd) and rel_w.shape == (2 * S - 1, hd)
    nbytes = lib().grove_attn_relpos_bwd_workspace_bytes(F, G, heads, hd, ws)
    wsb = torch.empty(nbytes // 4, device=qkv.device, dtype=F32)
    if lse is not None:
        _req(lse, F32, "lse")
    check(lib().grove_attn_relpos_bwd_lse(_p(qkv), _p(qkv_bias_bf16), _p(rel_h), _p(rel_w), _p(att), _p(datt), _p(dqkv), _p(wsb), _p(lse), F, G, heads,
                                          hd, ws, _stream(qkv)), "grove_attn_relpos_bwd")
    return dqkv


def box_losses_bwd(boxes, logits, gt, sel, labels, cg, co):
    B = boxes.shape[0]
    db = torch.empty(B, 4, device=boxes.device, dtype=F32)
    dl = torch.empty(B, device=boxes.device, dtype=F32)
    check(lib().grove_box_losses_bwd(_p(_req(boxes, F32, "boxes")), _p(_req(logits, F32, "logits")), _p(_req(gt, F32, "gt")),
                                     _p(_req(sel, torch.uint8, "sel")), _p(_req(labels, F32, "labels")), float(cg), float(co), _p(db), _p(dl), B,
                                     _stream(boxes)), "grove_box_losses_bwd")
    return db, dl


def launch_count() -> int:
    return int(lib().grove_launch_count())


def reset_launch_count() -> None:
    lib().grove_reset_launch_count()


def add_launch_count(n: int) -> None:
    lib().grove_add_launch_count(int(n))
