"""Training step of the box decoder (BASELINE config 4): forward with saved activations + hand-scheduled backward.

The reference trains the whole mask decoder through autograd (train.py:281-289; forward mask_decoder.py:155-205 over
transformer.py:62-182).  Here the backward pass is an explicit reverse schedule over the same CUDA library as the forward:

  token side (6 x 256 per instance, fp32)   small_linear / small_wgrad / colsum / LayerNorm-bwd / self-attention-bwd kernels
  image side (N x 256 per instance, bf16)    input gradients on the tcgen05 GEMM with transposed weights; weight gradients on the
                                             same GEMM over transposed activations (K = instances*N, split-K);
                                             (keys + pe) W is differentiated as keys.W + pe.W (the pe part reduces over instances
                                             first); layer 0's projections are shared by the phrases of a frame, so their
                                             cotangents are summed per frame before the GEMMs.

Gradients are accumulated in fp32 in a `GradStore` keyed by parameter.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from .common import bf16, f32


class GradStore:
    """fp32 gradient accumulators keyed by parameter (zero-initialised on first use).  `arena_elems` (the `elems` of the previous step's
    store) makes them views of ONE zero-filled buffer: a single fill launch per step instead of one per parameter (117 for ViT-B)."""

    def __init__(self, arena_elems: int = 0, device=None):
        self.g: Dict[nn.Parameter, torch.Tensor] = {}
        self.unpack = {}
        self.elems = 0                                    # 64-element-aligned total carved so far (what the next step's arena needs)
        self._arena = torch.zeros(arena_elems, device=device, dtype=torch.float32) if arena_elems > 0 and device is not None else None

    def buf(self, p: torch.Tensor, shape=None, unpack=None) -> torch.Tensor:
        """accumulator for parameter p; `shape` / `unpack` when the kernels produce the gradient in a packed layout"""
        t = self.g.get(p)
        if t is None:
            shape = tuple(p.shape) if shape is None else tuple(shape)
            n = 1
            for d in shape:
                n *= int(d)
            a = self._arena
            if a is not None and a.device == p.device and self.elems + n <= a.numel():
                t = a[self.elems:self.elems + n].view(shape)
            else:
                t = torch.zeros(shape, device=p.device, dtype=torch.float32)
            self.elems += (n + 63) // 64 * 64             # 256-byte aligned slices
            self.g[p] = t
            if unpack is not None:
                self.unpack[p] = unpack
        return t

    def grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        g = self.g.get(p)
        if g is None:
            return None
        return self.unpack[p](g) if p in self.unpack else g.reshape(p.shape)

    def apply(self, scale: float = 1.0) -> None:
        """p.grad (+)= scale * accumulated gradient, in the parameter's dtype and shape (trainable parameters only)"""
        for p in self.g:
            if not p.requires_grad:
                continue
            g = self.grad_of(p)
            g = (g * scale if scale != 1.0 else g).to(p.dtype)
            p.grad = g if p.grad is None else p.grad + g


def _wants(p: Optional[torch.Tensor]) -> bool:
    return p is not None and p.requires_grad


# ------------------------------------------------------------------ token-side helpers (fp32)
def lin_bwd(dec, key: str, lin: nn.Linear, dy: torch.Tensor, x: torch.Tensor, grads: GradStore, need_dx: bool = True):
    """y = x W^T + b on the token side: accumulates dW, db; returns dx."""
    if _wants(lin.weight):
        ops.small_wgrad(dy, x, grads.buf(lin.weight))
    if _wants(lin.bias):
        ops.colsum(dy, grads.buf(lin.bias))
    if not need_dx:
        return None
    N = dy.shape[1]
    if N % 4:   # objectness head: one output column — pad the cotangent and W^T to four columns
        pad = 4 - N % 4
        wt = dec._pack.get(key + ".wT32p", [lin.weight], lambda w: torch.cat([f32(w.t()), torch.zeros(w.shape[1], pad, device=w.device)], 1).contiguous())
        dy = torch.cat([dy, torch.zeros(dy.shape[0], pad, device=dy.device)], 1).contiguous()
    else:
        wt = dec._pack.get(key + ".wT32", [lin.weight], lambda w: f32(w.t()))
    return ops.small_linear(dy, wt)


def ln_bwd(ln: nn.LayerNorm, x, r, dy, grads: GradStore):
    """y = LN(x + r): returns d(x + r); accumulates d gamma / d beta"""
    g = f32(ln.weight)
    dx = torch.empty_like(dy)
    want = _wants(ln.weight) or _wants(ln.bias)
    ops.layernorm_bwd(x, g, dy, eps=ln.eps, r=r, dx_out=dx, dgamma=grads.buf(ln.weight) if want else None,
                      dbeta=grads.buf(ln.bias) if want else None)
    return dx


# ------------------------------------------------------------------ image-side helpers
def img_proj_bwd(dec, key: str, lin: nn.Linear, dy: torch.Tensor, keys: torch.Tensor, pe: Optional[torch.Tensor], N: int, grads: GradStore,
                 dkeys: torch.Tensor, first: bool):
    """y = (keys (+ pe)) W^T + b with y bf16 [rows, internal], keys bf16 [rows, C]:  dkeys (fp32 [rows, C]) (+)= dy W; dW, db accumulated."""
    wt = dec._pack.get(key + ".wT16", [lin.weight], lambda w: bf16(w.t()))
    ops.gemm(dy, wt, dkeys, resid=None if first else dkeys)
    if _wants(lin.weight):
        ops.wgrad(dy, keys, grads.buf(lin.weight))
        if pe is not None:
            reps = dy.shape[0] // N
            dys = ops.batch_sum_bf16(dy, reps).view(N, dy.shape[1]) if reps > 1 else dy
            ops.wgrad(dys, pe, grads.buf(lin.weight))
    if _wants(lin.bias):
        ops.colsum(dy, grads.buf(lin.bias))


def frame_sum_bf16(x_bf16: torch.Tensor, offsets: torch.Tensor, Fr: int, N: int) -> torch.Tensor:
    """per-instance bf16 [B*N, c] -> per-frame bf16 [Fr*N, c] (phrases of a frame share layer-0 keys)"""
    c = x_bf16.shape[1]
    xf = x_bf16.float()   # dtype plumbing; the reduction runs in grove_segment_sum_f32
    out = torch.empty(Fr, N * c, device=x_bf16.device, dtype=torch.float32)
    ops.segment_sum(xf.view(-1, N * c), offsets, out)
    o16 = torch.empty(Fr * N, c, device=x_bf16.device, dtype=torch.bfloat16)
    ops.cast_f32_bf16(out.view(-1), o16.view(-1))
    return o16


# ------------------------------------------------------------------ forward with tape
def decode_train(dec, tokens, keys0, shared, pe, frame_of, N, C):
    """MaskDecoder._decode with every intermediate the backward needs kept on a tape (same kernels, same arithmetic)."""
    tr = dec.transformer
    B, T, _ = tokens.shape
    H = tr.num_heads
    R = B * T
    dev = tokens.device
    tok = tokens.reshape(R, C)
    tape = {"tok": tok, "B": B, "T": T, "N": N, "C": C, "frame_of": frame_of, "keys0": keys0, "pe": pe, "layers": []}
    queries, q_in = tok, tok
    keys, src_of = keys0, frame_of
    for li, layer in enumerate(tr.layers):
        k = f"l{li}"
        L = {"queries_in": queries, "q_in_in": q_in, "keys_in": keys, "src_of": src_of}
        sa = layer.self_attn
        wq, bq = dec._w32(k + ".sa.q", sa.q_proj); wk, bk = dec._w32(k + ".sa.k", sa.k_proj); wv, bv = dec._w32(k + ".sa.v", sa.v_proj)
        L["sa_q"], L["sa_k"], L["sa_v"] = ops.small_linear(q_in, wq, bq), ops.small_linear(q_in, wk, bk), ops.small_linear(queries, wv, bv)
        L["sa_att"] = ops.token_self_attention(L["sa_q"], L["sa_k"], L["sa_v"], B, T, H, sa.internal_dim // H)
        L["sa_o"] = dec._token_attn_out(k + ".sa", sa, L["sa_att"])
        g, b = dec._ln(k + ".n1", layer.norm1)
        queries, q_in = ops.add_layernorm(L["sa_o"], None if layer.skip_first_layer_pe else queries, g, b, eps=layer.norm1.eps, add2=tok)
        L["queries1"], L["q_in1"] = queries, q_in
        ca = layer.cross_attn_token_to_image
        dh = ca.internal_dim // H
        wq, bq = dec._w32(k + ".t2i.q", ca.q_proj)
        L["t_q"] = ops.small_linear(q_in, wq, bq)
        if li == 0:
            kp, vp = shared["k"], shared["v"]
        else:
            kp = dec._image_proj(k + ".t2i.k", ca.k_proj, keys, pe, N)
            vp = dec._image_proj(k + ".t2i.v", ca.v_proj, keys, None, N)
        L["t_kp"], L["t_vp"] = kp, vp
        L["t_lse"] = torch.empty(B, T, H, device=dev, dtype=torch.float32)
        L["t_att"] = ops.t2i_attention(L["t_q"], kp, vp, src_of, B, T, N, H, dh, lse=L["t_lse"]).reshape(R, ca.internal_dim)
        L["t_o"] = dec._token_attn_out(k + ".t2i", ca, L["t_att"])
        g, b = dec._ln(k + ".n2", layer.norm2)
        queries = ops.add_layernorm(queries, L["t_o"], g, b, eps=layer.norm2.eps)
        L["queries2"] = queries
        w1, b1 = dec._w32(k + ".m1", layer.mlp.lin1); w2, b2 = dec._w32(k + ".m2", layer.mlp.lin2)
        L["h1"] = ops.small_linear(queries, w1, b1, act="relu")
        L["m"] = ops.small_linear(L["h1"], w2, b2)
        g, b = dec._ln(k + ".n3", layer.norm3)
        queries, q_in = ops.add_layernorm(queries, L["m"], g, b, eps=layer.norm3.eps, add2=tok)
        L["queries3"], L["q_in3"] = queries, q_in
        ia = layer.cross_attn_image_to_token
        wk, bk = dec._w32(k + ".i2t.k", ia.k_proj); wv, bv = dec._w32(k + ".i2t.v", ia.v_proj)
        L["i_kt"], L["i_vt"] = ops.small_linear(q_in, wk, bk), ops.small_linear(queries, wv, bv)
        L["i_qi"] = shared["qi"] if li == 0 else dec._image_proj(k + ".i2t.q", ia.q_proj, keys, pe, N)
        ai = torch.empty(B * N, ia.internal_dim, device=dev, dtype=torch.bfloat16)
        ops.i2t_attention(L["i_qi"], L["i_kt"], L["i_vt"], src_of, ai, B, T, N, H, ia.internal_dim // H)
        L["i_ai"] = ai
        wo, bo = dec._w16(k + ".i2t.o", ia.out_proj)
        delta = torch.empty(B * N, C, device=dev, dtype=torch.float32)
        ops.gemm(ai, wo, delta, bias=bo)
        L["delta"] = delta
        g, b = dec._ln(k + ".n4", layer.norm4)
        new_keys = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
        ops.keys_add_ln(keys, src_of, delta, g, b, new_keys, B, N, C, eps=layer.norm4.eps)
        keys, src_of = new_keys, None
        tape["layers"].append(L)
    fa = tr.final_attn_token_to_image
    wq, bq = dec._w32("f.q", fa.q_proj)
    tape["f_queries"], tape["f_q_in"], tape["f_keys"] = queries, q_in, keys
    tape["f_q"] = ops.small_linear(q_in, wq, bq)
    tape["f_kp"] = dec._image_proj("f.k", fa.k_proj, keys, pe, N)
    tape["f_vp"] = dec._image_proj("f.v", fa.v_proj, keys, None, N)
    tape["f_lse"] = torch.empty(B, T, H, device=dev, dtype=torch.float32)
    tape["f_att"] = ops.t2i_attention(tape["f_q"], tape["f_kp"], tape["f_vp"], None, B, T, N, H, fa.internal_dim // H,
                                      lse=tape["f_lse"]).reshape(R, fa.internal_dim)
    tape["f_o"] = dec._token_attn_out("f", fa, tape["f_att"])
    g, b = dec._ln("f.n", tr.norm_final_attn)
    hs = ops.add_layernorm(queries, tape["f_o"], g, b, eps=tr.norm_final_attn.eps)
    qo = hs.view(B, T, C)[:, 1 + dec.num_mask_tokens, :].contiguous()
    tape["qo"] = qo
    w0, b0 = dec._w32("h.0", dec.bbox_prediction_head[0]); w2, b2 = dec._w32("h.2", dec.bbox_prediction_head[2])
    tape["hb"] = ops.small_linear(qo, w0, b0, act="relu")
    boxes = ops.small_linear(tape["hb"], w2, b2, act="sigmoid")
    tape["boxes"] = boxes
    if dec.use_temp_objectness:
        wt, bt = dec._w32("h.t", dec.temporal_objectness_head)
        logits = ops.small_linear(qo, wt, bt).reshape(B)
    else:
        logits = torch.zeros(B, device=dev, dtype=torch.float32)
    return boxes, logits, tape


# ------------------------------------------------------------------ backward
def _t2i_bwd(dec, key, attn, L_q, kp, vp, src_of, att, lse, q_in, d_o_in, keys, pe, tape, grads, dkeys, first, frame_off, Fr):
    """backward of out_proj(T2I(q_proj(q_in), k_proj(keys+pe), v_proj(keys))) given the cotangent of its output; returns d q_in"""
    B, T, N, H = tape["B"], tape["T"], tape["N"], dec.transformer.num_heads
    d_att = lin_bwd(dec, key + ".o", attn.out_proj, d_o_in, att, grads)
    dq, dkp, dvp = ops.t2i_attention_bwd(L_q, kp, vp, src_of, att.view(B, T, -1), d_att.view(B, T, -1), lse, B, T, N, H, attn.internal_dim // H)
    d_q_in = lin_bwd(dec, key + ".q", attn.q_proj, dq.view(B * T, -1), q_in, grads)
    if src_of is not None:   # layer 0: the projections belong to the frame — sum the phrases' cotangents first
        dkp, dvp = frame_sum_bf16(dkp, frame_off, Fr, N), frame_sum_bf16(dvp, frame_off, Fr, N)
    img_proj_bwd(dec, key + ".k", attn.k_proj, dkp, keys, pe, N, grads, dkeys, first)
    img_proj_bwd(dec, key + ".v", attn.v_proj, dvp, keys, None, N, grads, dkeys, False)
    return d_q_in


def decode_backward(dec, tape, dboxes: torch.Tensor, dlogits: Optional[torch.Tensor], grads: GradStore, Fr: int):
    """Reverse schedule of decode_train.  Returns (d keys0 fp32 [Fr*N, C], d tokens fp32 [B, T, C])."""
    tr = dec.transformer
    B, T, N, C = tape["B"], tape["T"], tape["N"], tape["C"]
    H = tr.num_heads
    R = B * T
    dev = dboxes.device
    tok, pe, frame_of = tape["tok"], tape["pe"], tape["frame_of"]
    counts = torch.bincount(frame_of.long(), minlength=Fr)
    frame_off = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).to(torch.int32)

    # ---- heads (mask_decoder.py:191-203)
    h0, h2 = dec.bbox_prediction_head[0], dec.bbox_prediction_head[2]
    d2 = ops.act_bwd(dboxes.contiguous(), tape["boxes"], "sigmoid")
    d_hb = lin_bwd(dec, "h.2", h2, d2, tape["hb"], grads)
    d0 = ops.act_bwd(d_hb, tape["hb"], "relu")
    d_qo = lin_bwd(dec, "h.0", h0, d0, tape["qo"], grads)
    if dec.use_temp_objectness and dlogits is not None:
        d_qo = d_qo + lin_bwd(dec, "h.t", dec.temporal_objectness_head, dlogits.reshape(B, 1).contiguous(), tape["qo"], grads)
    d_hs = torch.zeros(B, T, C, device=dev, dtype=torch.float32)
    d_hs[:, 1 + dec.num_mask_tokens, :] = d_qo
    d_hs = d_hs.view(R, C)

    # ---- final token -> image attention + norm_final_attn (transformer.py:99-104)
    fa = tr.final_attn_token_to_image
    d_sum = ln_bwd(tr.norm_final_attn, tape["f_queries"], tape["f_o"], d_hs, grads)
    d_queries = d_sum
    d_keys = torch.empty(B * N, C, device=dev, dtype=torch.float32)
    d_q_in = _t2i_bwd(dec, "f", fa, tape["f_q"], tape["f_kp"], tape["f_vp"], None, tape["f_att"], tape["f_lse"], tape["f_q_in"], d_sum,
                      tape["f_keys"], pe, tape, grads, d_keys, True, frame_off, Fr)
    d_queries = d_queries + d_q_in
    d_tok = d_q_in.clone()

    for li in reversed(range(len(tr.layers))):
        layer, L, k = tr.layers[li], tape["layers"][li], f"l{li}"
        src_of, keys_in = L["src_of"], L["keys_in"]
        shared = src_of is not None
        # (4) keys' = norm4(keys + out_proj(I2T(q_proj(keys+pe), k_proj(q_in3), v_proj(queries3))))   (transformer.py:175-180)
        ia = layer.cross_attn_image_to_token
        n4 = layer.norm4
        want4 = _wants(n4.weight) or _wants(n4.bias)
        d_x = torch.empty(B * N, C, device=dev, dtype=torch.float32)          # cotangent of keys + delta, per instance
        d_delta = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
        ops.layernorm_bwd(keys_in, f32(n4.weight), d_keys, eps=n4.eps, r=L["delta"], dx_out=d_x, dx_bf16=d_delta,
                          dgamma=grads.buf(n4.weight) if want4 else None, dbeta=grads.buf(n4.bias) if want4 else None,
                          keys_src_of=src_of, keys_N=N)
        wot = dec._pack.get(k + ".i2t.o.wT16", [ia.out_proj.weight], lambda w: bf16(w.t()))
        d_ai = torch.empty(B * N, ia.internal_dim, device=dev, dtype=torch.bfloat16)
        ops.gemm(d_delta, wot, d_ai)
        if _wants(ia.out_proj.weight):
            ops.wgrad(d_delta, L["i_ai"], grads.buf(ia.out_proj.weight))
        if _wants(ia.out_proj.bias):
            ops.colsum(d_delta, grads.buf(ia.out_proj.bias))
        d_qi, d_kt, d_vt = ops.i2t_attention_bwd(L["i_qi"], L["i_kt"], L["i_vt"], src_of, d_ai, B, T, N, H, ia.internal_dim // H)
        d_q_in3 = lin_bwd(dec, k + ".i2t.k", ia.k_proj, d_kt.view(R, -1), L["q_in3"], grads)
        d_queries = d_queries + d_q_in3 + lin_bwd(dec, k + ".i2t.v", ia.v_proj, d_vt.view(R, -1), L["queries3"], grads)
        d_tok = d_tok + d_q_in3
        if shared:
            d_keys_in = torch.empty(Fr * N, C, device=dev, dtype=torch.float32)
            ops.segment_sum(d_x.view(B, N * C), frame_off, d_keys_in.view(Fr, N * C))
            d_qi = frame_sum_bf16(d_qi, frame_off, Fr, N)
        else:
            d_keys_in = d_x
        img_proj_bwd(dec, k + ".i2t.q", ia.q_proj, d_qi, keys_in, pe, N, grads, d_keys_in, False)
        # (3) queries3 = norm3(queries2 + mlp(queries2))
        d_sum = ln_bwd(layer.norm3, L["queries2"], L["m"], d_queries, grads)
        d_h1 = lin_bwd(dec, k + ".m2", layer.mlp.lin2, d_sum, L["h1"], grads)
        d_queries = d_sum + lin_bwd(dec, k + ".m1", layer.mlp.lin1, ops.act_bwd(d_h1, L["h1"], "relu"), L["queries2"], grads)
        # (2) queries2 = norm2(queries1 + T2I(...))
        ca = layer.cross_attn_token_to_image
        d_sum = ln_bwd(layer.norm2, L["queries1"], L["t_o"], d_queries, grads)
        d_q_in1 = _t2i_bwd(dec, k + ".t2i", ca, L["t_q"], L["t_kp"], L["t_vp"], src_of, L["t_att"], L["t_lse"], L["q_in1"], d_sum, keys_in, pe,
                           tape, grads, d_keys_in, False, frame_off, Fr)
        d_queries = d_sum + d_q_in1
        d_tok = d_tok + d_q_in1
        # (1) queries1 = norm1([queries +] self_attn(q_in, q_in, queries))
        sa = layer.self_attn
        d_sum = ln_bwd(layer.norm1, L["sa_o"], None if layer.skip_first_layer_pe else L["queries_in"], d_queries, grads)
        d_att = lin_bwd(dec, k + ".sa.o", sa.out_proj, d_sum, L["sa_att"], grads)
        dq, dk, dv = ops.token_self_attention_bwd(L["sa_q"], L["sa_k"], L["sa_v"], d_att, B, T, H, sa.internal_dim // H)
        d_qin = lin_bwd(dec, k + ".sa.q", sa.q_proj, dq, L["q_in_in"], grads) + lin_bwd(dec, k + ".sa.k", sa.k_proj, dk, L["q_in_in"], grads)
        d_v_in = lin_bwd(dec, k + ".sa.v", sa.v_proj, dv, L["queries_in"], grads)
        if layer.skip_first_layer_pe:       # layer 0: queries == q_in == tokens, no residual
            d_tok = d_tok + d_qin + d_v_in
            d_queries = None
        else:
            d_queries = d_sum + d_qin + d_v_in
            d_tok = d_tok + d_qin
        d_keys = d_keys_in
    if d_queries is not None:   # a stack whose first layer keeps the residual: the layer input is the tokens themselves
        d_tok = d_tok + d_queries
    return d_keys, d_tok.view(B, T, C)
