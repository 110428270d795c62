"""Box decoder ("MaskDecoder" with decoding_type="query") — reference: model/SAM/modeling/mask_decoder.py.

Same constructor / forward contract as the reference (mask_decoder.py:19-134); only the query path that GROVE
runs (:191-205) is built.  The two-way transformer (transformer.py:62-182) is lowered as:

  image side  (N x 256 per instance, bf16 in HBM): k/v/q projections and the image->token out_proj on the tcgen05
              GEMM; `keys + pe` is never materialised — (keys+pe)W = keys.W + (pe.W), and pe.W is added by the GEMM
              epilogue as a row-periodic fp32 residual; layer-0 projections are computed once per FRAME because all
              phrases of a frame share src = emb + no_mask until the first image->token update (:178-185);
  token side  (6 x 256 per instance, fp32): small fused fp32 kernels.
"""
from __future__ import annotations

from typing import List, Tuple, Type

import torch
import torch.nn as nn

from .. import ops
from .common import LayerNorm2d, PackCache, _ContainerOnly, bf16, f32, no_grad_entry


class MLP(_ContainerOnly):
    """mask_decoder.py:230-254 (only used by the mask / IoU heads that GROVE never runs)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, sigmoid_output: bool = False) -> None:
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3, activation: Type[nn.Module] = nn.GELU,
                 iou_head_depth: int = 3, iou_head_hidden_dim: int = 256, decoding_type: str = "query", use_temp_objectness: bool = True) -> None:
        super().__init__()
        if decoding_type != "query":
            raise NotImplementedError("grove_b200 builds the box ('query') decoding path only; the mask path never runs in GROVE")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        # unused by the query path; kept so reference checkpoints load key-for-key (mask_decoder.py:57-78)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2), LayerNorm2d(transformer_dim // 4), activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2), activation())
        self.output_hypernetworks_mlps = nn.ModuleList([MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3)
                                                        for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self.decoding_type = decoding_type
        self.bbox_prediction_head = nn.Sequential(nn.Linear(transformer_dim, transformer_dim), nn.ReLU(), nn.Linear(transformer_dim, 4))
        if use_temp_objectness:
            self.temporal_objectness_head = nn.Linear(transformer_dim, 1)
            self.use_temp_objectness = True
        else:
            self.use_temp_objectness = False
        self._pack = PackCache()
        self._pew = {}
        self._idx_cache = {}
        self.max_instances_per_pass = 256  # bounds the per-instance key buffers (N x 256 bf16 + fp32 delta)

    # ------------------------------------------------------------------ helpers
    def _w32(self, key, lin):
        return self._pack.get(key + ".w32", [lin.weight], f32), self._pack.get(key + ".b32", [lin.bias], f32)

    def _w16(self, key, lin):
        return self._pack.get(key + ".w16", [lin.weight], bf16), self._pack.get(key + ".b32", [lin.bias], f32)

    def _ln(self, key, ln):
        return self._pack.get(key + ".g", [ln.weight], f32), self._pack.get(key + ".b", [ln.bias], f32)

    @staticmethod
    def _tokens_of(x: torch.Tensor, dtype) -> torch.Tensor:
        """[F,C,G,G] (NCHW or channels-last view) -> token-major [F, G*G, C] of `dtype`."""
        Fr, C, G, _ = x.shape
        if x.permute(0, 2, 3, 1).is_contiguous():
            t = x.permute(0, 2, 3, 1).reshape(Fr, G * G, C)
            return t if t.dtype == dtype else t.to(dtype)
        if dtype == torch.bfloat16:
            xc = x.to(torch.bfloat16).contiguous()
            out = torch.empty(Fr, G * G, C, device=x.device, dtype=torch.bfloat16)
            return ops.nchw_to_tokens(xc, out, Fr, G * G, C)
        return x.permute(0, 2, 3, 1).reshape(Fr, G * G, C).to(dtype).contiguous()

    # ------------------------------------------------------------------ forward
    def forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings: torch.Tensor,
                dense_prompt_embeddings: torch.Tensor, multimask_output: bool, reps: List[int]):
        """mask_decoder.py:91-134: returns (bbox_pred [B,4], temp_objectness_logits [B]) or bbox_pred."""
        boxes, logits = self.predict_masks(image_embeddings=image_embeddings, image_pe=image_pe,
                                           sparse_prompt_embeddings=sparse_prompt_embeddings,
                                           dense_prompt_embeddings=dense_prompt_embeddings, reps=reps)
        return (boxes, logits) if self.use_temp_objectness else boxes

    @no_grad_entry("MaskDecoder.forward", lambda self: self.parameters())
    def predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, reps: List[int]):
        """mask_decoder.py:155-205 (query path)."""
        if not image_embeddings.is_cuda:
            raise RuntimeError("grove_b200.MaskDecoder runs on CUDA only (no CPU fallback)")
        Fr, C, G, _ = image_embeddings.shape
        B = sparse_prompt_embeddings.shape[0]
        if len(reps) != Fr or sum(reps) != B:
            raise ValueError(f"reps must have one entry per frame summing to the number of prompts (got {len(reps)} entries, sum {sum(reps)}, "
                             f"{Fr} frames, {B} prompts)")
        if sparse_prompt_embeddings.shape[1] != 1:
            raise NotImplementedError("one text token per prompt (GROVE.py:272)")
        d = dense_prompt_embeddings
        # a broadcast view of one embedding: expand() keeps the stride of size-1 dims (B == 1 has stride(0) == C, a G == 1 grid non-zero spatial strides)
        if d.dim() != 4 or (d.shape[0] > 0 and any(d.shape[k] > 1 and d.stride(k) != 0 for k in (0, 2, 3))):
            raise NotImplementedError("dense prompts other than the broadcast no-mask embedding are out of scope (prompt_encoder.py:182-184)")
        out_dtype = sparse_prompt_embeddings.dtype
        dev = image_embeddings.device
        if B == 0:
            return torch.zeros(0, 4, device=dev, dtype=out_dtype), torch.zeros(0, device=dev, dtype=out_dtype)
        no_mask = d[0, :, 0, 0].to(torch.float32).contiguous()
        emb = self._tokens_of(image_embeddings, torch.bfloat16)                       # [F,N,C]
        text = sparse_prompt_embeddings.reshape(B, C).to(torch.float32)
        with ops.device_of(emb):
            rec = self.decode_records(emb, image_pe, text, no_mask, reps)
        return rec[:, :4].to(out_dtype), rec[:, 4].to(out_dtype)

    def _frame_index(self, reps, dev) -> torch.Tensor:
        """int32 [B] on `dev`: the frame every (frame, phrase) instance reads (the reference's index_select, mask_decoder.py:178-185).
        Cached per `reps`, so a steady-state step (and a CUDA-graph capture) does no host->device copy for it."""
        key = (tuple(reps), dev)
        hit = self._idx_cache.get(key)
        if hit is None:
            if len(self._idx_cache) > 64:
                self._idx_cache.clear()
            hit = torch.repeat_interleave(torch.arange(len(reps)), torch.tensor(reps)).to(torch.int32).to(dev)
            self._idx_cache[key] = hit
        return hit

    @torch.no_grad()
    def decode_records(self, emb_tokens: torch.Tensor, image_pe: torch.Tensor, text: torch.Tensor, no_mask: torch.Tensor, reps: List[int],
                       records: torch.Tensor = None) -> torch.Tensor:
        """The box decoder on token-major inputs: emb_tokens bf16 [F, N, C] (encoder output), image_pe [1,C,G,G], text fp32 [B, C] (one prompt
        token per instance), no_mask fp32 [C] -> packed fp32 records [B, 5] = (cx, cy, w, h, objectness logit), written by the heads kernel
        (into `records` when given, e.g. a slice of the all-gather buffer of parallel.ground_sharded_clip).  No host synchronisation."""
        Fr, N, C = emb_tokens.shape
        B = text.shape[0]
        dev = emb_tokens.device
        if records is None:
            records = torch.empty(B, 5, device=dev, dtype=torch.float32)
        if B == 0:
            return records
        pe = self._tokens_of(image_pe, torch.float32).reshape(N, C)                   # [N,C] (a view of the caller's tensor when possible)
        if not pe.is_contiguous():
            pe = pe.contiguous()
        keys0 = torch.empty(Fr * N, C, device=dev, dtype=torch.bfloat16)
        ops.add_rowvec_bf16(emb_tokens.reshape(Fr * N, C), no_mask, keys0)
        out_tok = torch.cat([self.iou_token.weight, self.mask_tokens.weight], 0).to(torch.float32)
        frame_of_all = self._frame_index(reps, dev)
        shared = self._layer0_shared(keys0, pe, N, C)
        step = self.max_instances_per_pass
        for s in range(0, B, step):
            e = min(B, s + step)
            tokens = torch.cat([out_tok.unsqueeze(0).expand(e - s, -1, -1), text[s:e].unsqueeze(1)], 1).contiguous()
            self._decode(tokens, keys0, shared, pe, frame_of_all[s:e], N, C, records[s:e])
        return records

    # ------------------------------------------------------------------ training step (forward with tape, backward)
    def predict_masks_train(self, emb_tokens: torch.Tensor, image_pe: torch.Tensor, text: torch.Tensor, no_mask: torch.Tensor, reps: List[int]):
        """predict_masks for the training step: emb_tokens bf16 [F, N, C] (token-major encoder output), text fp32 [B, C] (one prompt
        token per instance), no_mask fp32 [C].  Returns fp32 boxes [B,4], logits [B] and the tape for `backward`."""
        from .decoder_train import decode_train
        Fr, N, C = emb_tokens.shape
        B = text.shape[0]
        if len(reps) != Fr or sum(reps) != B or B == 0:
            raise ValueError("reps must have one entry per frame summing to the (non-zero) number of prompts")
        dev = emb_tokens.device
        G = int(round(N ** 0.5))
        pe = self._tokens_of(image_pe, torch.float32).reshape(N, C)
        if not pe.is_contiguous():
            pe = pe.contiguous()
        keys0 = torch.empty(Fr * N, C, device=dev, dtype=torch.bfloat16)
        ops.add_rowvec_bf16(emb_tokens.reshape(Fr * N, C), no_mask, keys0)
        out_tok = torch.cat([self.iou_token.weight, self.mask_tokens.weight], 0).to(torch.float32)
        frame_of = torch.repeat_interleave(torch.arange(Fr), torch.tensor(reps)).to(torch.int32).to(dev)
        shared = self._layer0_shared(keys0, pe, N, C)
        tokens = torch.cat([out_tok.unsqueeze(0).expand(B, -1, -1), text.to(torch.float32).unsqueeze(1)], 1).contiguous()
        boxes, logits, tape = decode_train(self, tokens, keys0, shared, pe, frame_of, N, C)
        tape["Fr"], tape["G"] = Fr, G
        return boxes, logits, tape

    def backward(self, tape, dboxes: torch.Tensor, dlogits, grads):
        """Reverse schedule of predict_masks_train.  Returns (d emb_tokens fp32 [F*N, C], d text fp32 [B, C]) and accumulates the
        decoder's parameter gradients (transformer, heads, iou/mask tokens) into `grads`."""
        from .decoder_train import decode_backward
        d_keys0, d_tokens = decode_backward(self, tape, dboxes, dlogits, grads, tape["Fr"])
        B, T, C = d_tokens.shape
        dsum = torch.zeros(T * C, device=d_tokens.device, dtype=torch.float32)
        ops.colsum(d_tokens.reshape(B, T * C).contiguous(), dsum)
        dsum = dsum.view(T, C)
        if self.iou_token.weight.requires_grad:
            grads.buf(self.iou_token.weight).add_(dsum[0:1])
        if self.mask_tokens.weight.requires_grad:
            grads.buf(self.mask_tokens.weight).add_(dsum[1:1 + self.num_mask_tokens])
        return d_keys0, d_tokens[:, 1 + self.num_mask_tokens, :].contiguous()

    # ------------------------------------------------------------------ the two-way transformer
    def _pe_w(self, key, lin, pe):
        """(pe . W^T) [N, internal] fp32 — the positional part of (keys + pe) W.  Cached per (weight, pe tensor); the cache
        entry keeps `pe` alive, so its data pointer cannot be recycled for a different table while the entry exists."""
        w, _ = self._w32(key, lin)
        hit = self._pew.get(key)
        sig = (lin.weight.data_ptr(), lin.weight._version, pe.data_ptr(), pe._version, tuple(pe.shape))
        if hit is None or hit[0] != sig:
            hit = (sig, pe, ops.small_linear(pe, w))
            self._pew[key] = hit
        return hit[2]

    def _image_proj(self, key, lin, keys, pe, N):
        """bf16 [rows, internal] = (keys (+pe)) W^T + b on the tcgen05 GEMM."""
        w, b = self._w16(key, lin)
        out = torch.empty(keys.shape[0], w.shape[0], device=keys.device, dtype=torch.bfloat16)
        if pe is None:
            return ops.gemm(keys, w, out, bias=b)
        return ops.gemm(keys, w, out, bias=b, resid=self._pe_w(key, lin, pe), resid_row_mod=N)

    def _layer0_shared(self, keys0, pe, N, C):
        l0 = self.transformer.layers[0]
        return {"k": self._image_proj("l0.t2i.k", l0.cross_attn_token_to_image.k_proj, keys0, pe, N),
                "v": self._image_proj("l0.t2i.v", l0.cross_attn_token_to_image.v_proj, keys0, None, N),
                "qi": self._image_proj("l0.i2t.q", l0.cross_attn_image_to_token.q_proj, keys0, pe, N)}

    def _token_attn_out(self, key, attn, att, resid=None):
        w, b = self._w32(key + ".o", attn.out_proj)
        return ops.small_linear(att, w, b, resid=resid)

    def _wt32(self, key, lin):
        """(W^T [in, out] fp32, bias fp32): the layout of the fused token kernels"""
        return self._pack.get(key + ".wT32", [lin.weight], lambda w: f32(w.t())), self._pack.get(key + ".b32", [lin.bias], f32)

    def _fused_ok(self) -> bool:
        tr = self.transformer
        return (self.transformer_dim == 256 and tr.num_heads == 8 and tr.mlp_dim == 2048 and self.num_mask_tokens == 4
                and all(l.cross_attn_token_to_image.internal_dim == 128 and l.self_attn.internal_dim == 256 for l in tr.layers))

    def _decode(self, tokens, keys0, shared, pe, frame_of, N, C, records):
        """One pass of the two-way transformer + heads over <= max_instances_per_pass instances.  Token side: two fused kernels per layer
        around the token->image attention (csrc/decoder_fused.cu); image side: tcgen05 GEMMs + the two attention kernels + norm4."""
        if not self._fused_ok():
            return self._decode_unfused(tokens, keys0, shared, pe, frame_of, N, C, records)
        tr = self.transformer
        B, T, _ = tokens.shape
        H = tr.num_heads
        queries = tokens
        keys, src_of = keys0, frame_of       # layer 0 reads the per-frame keys through the index
        qf = None
        last = len(tr.layers) - 1
        for li, layer in enumerate(tr.layers):
            k = f"l{li}"
            if layer.skip_first_layer_pe and li != 0:
                raise NotImplementedError("skip_first_layer_pe is only meaningful on layer 0 (transformer.py:45-55)")
            sa, ca, ia = layer.self_attn, layer.cross_attn_token_to_image, layer.cross_attn_image_to_token
            pa = {}
            for nm, lin in (("q", sa.q_proj), ("k", sa.k_proj), ("v", sa.v_proj), ("o", sa.out_proj)):
                pa[f"w{nm}_t"], pa[f"b{nm}"] = self._wt32(f"{k}.sa.{nm}", lin)
            pa["ln_g"], pa["ln_b"] = self._ln(k + ".n1", layer.norm1)
            pa["ln_eps"] = float(layer.norm1.eps)
            pa["wq2_t"], pa["bq2"] = self._wt32(k + ".t2i.q", ca.q_proj)
            pa["skip_pe"] = 1 if layer.skip_first_layer_pe else 0
            queries, qt = ops.twoway_tokens_a(queries, tokens, pa)
            # token -> image cross attention (:164-169)
            dh = ca.internal_dim // H
            if li == 0:
                kp, vp = shared["k"], shared["v"]
            else:
                kp = self._image_proj(k + ".t2i.k", ca.k_proj, keys, pe, N)
                vp = self._image_proj(k + ".t2i.v", ca.v_proj, keys, None, N)
            att = ops.t2i_attention_wide(qt, kp, vp, src_of, B, T, N, H, dh)
            pb = {}
            pb["wo_t"], pb["bo"] = self._wt32(k + ".t2i.o", ca.out_proj)
            pb["ln2_g"], pb["ln2_b"] = self._ln(k + ".n2", layer.norm2)
            pb["ln2_eps"] = float(layer.norm2.eps)
            pb["w1_t"], pb["b1"] = self._wt32(k + ".m1", layer.mlp.lin1)
            pb["w2_t"], pb["b2"] = self._wt32(k + ".m2", layer.mlp.lin2)
            pb["mlp_dim"] = int(tr.mlp_dim)
            pb["ln3_g"], pb["ln3_b"] = self._ln(k + ".n3", layer.norm3)
            pb["ln3_eps"] = float(layer.norm3.eps)
            pb["wk_t"], pb["bk"] = self._wt32(k + ".i2t.k", ia.k_proj)
            pb["wv_t"], pb["bv"] = self._wt32(k + ".i2t.v", ia.v_proj)
            fa = tr.final_attn_token_to_image
            if li == last:
                pb["wqf_t"], pb["bqf"] = self._wt32("f.q", fa.q_proj)
            else:
                pb["wqf_t"], pb["bqf"] = None, None
            queries, kt, vt, qf = ops.twoway_tokens_b(queries, att, tokens, pb, want_qf=(li == last))
            # image -> token cross attention, updates the keys (:175-180)
            qi = shared["qi"] if li == 0 else self._image_proj(k + ".i2t.q", ia.q_proj, keys, pe, N)
            ai = torch.empty(B * N, ia.internal_dim, device=tokens.device, dtype=torch.bfloat16)
            ops.i2t_attention(qi, kt, vt, src_of, ai, B, T, N, H, ia.internal_dim // H)
            wo, bo = self._w16(k + ".i2t.o", ia.out_proj)
            delta = torch.empty(B * N, C, device=tokens.device, dtype=torch.float32)
            ops.gemm(ai, wo, delta, bias=bo)
            g, b = self._ln(k + ".n4", layer.norm4)
            new_keys = torch.empty(B * N, C, device=tokens.device, dtype=torch.bfloat16)
            ops.keys_add_ln(keys, src_of, delta, g, b, new_keys, B, N, C, eps=layer.norm4.eps)
            keys, src_of = new_keys, None
            del delta, ai
        # final token -> image attention (transformer.py:99-104) + norm_final_attn + heads (mask_decoder.py:191-203) -> packed record
        fa = tr.final_attn_token_to_image
        kp = self._image_proj("f.k", fa.k_proj, keys, pe, N)
        vp = self._image_proj("f.v", fa.v_proj, keys, None, N)
        att = ops.t2i_attention_wide(qf, kp, vp, src_of, B, T, N, H, fa.internal_dim // H)
        wo, bo = self._w32("f.o", fa.out_proj)
        g, b = self._ln("f.n", tr.norm_final_attn)
        w0, b0 = self._w32("h.0", self.bbox_prediction_head[0]); w2, b2 = self._w32("h.2", self.bbox_prediction_head[2])
        wt, bt = self._w32("h.t", self.temporal_objectness_head) if self.use_temp_objectness else (None, None)
        return ops.decoder_heads(queries, att, wo, bo, g, b, tr.norm_final_attn.eps, w0, b0, w2, b2, wt, bt, records, tok=1 + self.num_mask_tokens)

    def _decode_unfused(self, tokens, keys0, shared, pe, frame_of, N, C, records):
        """the same pass on the per-op kernels of decoder_ops.cu (other widths than 256 / 128 / 2048; cross-check of the fused path)"""
        tr = self.transformer
        B, T, _ = tokens.shape
        H = tr.num_heads
        R = B * T
        tok = tokens.reshape(R, C)
        queries = tok
        q_in = tok                           # queries + query_pe; layer 0 skips the PE (queries == tokens there)
        keys, src_of = keys0, frame_of       # layer 0 reads the per-frame keys through the index
        for li, layer in enumerate(tr.layers):
            k = f"l{li}"
            sa = layer.self_attn
            # (1) token self-attention (transformer.py:155-161)
            if layer.skip_first_layer_pe and li != 0:
                raise NotImplementedError("skip_first_layer_pe is only meaningful on layer 0 (transformer.py:45-55)")
            wq, bq = self._w32(k + ".sa.q", sa.q_proj); wk, bk = self._w32(k + ".sa.k", sa.k_proj); wv, bv = self._w32(k + ".sa.v", sa.v_proj)
            q = ops.small_linear(q_in, wq, bq); kk = ops.small_linear(q_in, wk, bk); v = ops.small_linear(queries, wv, bv)
            att = ops.token_self_attention(q, kk, v, B, T, H, sa.internal_dim // H)
            o = self._token_attn_out(k + ".sa", sa, att)
            g, b = self._ln(k + ".n1", layer.norm1)
            queries, q_in = ops.add_layernorm(o, None if layer.skip_first_layer_pe else queries, g, b, eps=layer.norm1.eps, add2=tok)
            # (2) token -> image cross attention (:164-169)
            ca = layer.cross_attn_token_to_image
            dh = ca.internal_dim // H
            wq, bq = self._w32(k + ".t2i.q", ca.q_proj)
            q = ops.small_linear(q_in, wq, bq)
            if li == 0:
                kp, vp = shared["k"], shared["v"]
            else:
                kp = self._image_proj(k + ".t2i.k", ca.k_proj, keys, pe, N)
                vp = self._image_proj(k + ".t2i.v", ca.v_proj, keys, None, N)
            att = ops.t2i_attention(q, kp, vp, src_of, B, T, N, H, dh).reshape(R, ca.internal_dim)
            o = self._token_attn_out(k + ".t2i", ca, att)
            g, b = self._ln(k + ".n2", layer.norm2)
            queries = ops.add_layernorm(queries, o, g, b, eps=layer.norm2.eps)
            # (3) MLP (:171-173)
            w1, b1 = self._w32(k + ".m1", layer.mlp.lin1); w2, b2 = self._w32(k + ".m2", layer.mlp.lin2)
            m = ops.small_linear(ops.small_linear(queries, w1, b1, act="relu"), w2, b2)
            g, b = self._ln(k + ".n3", layer.norm3)
            queries, q_in = ops.add_layernorm(queries, m, g, b, eps=layer.norm3.eps, add2=tok)
            # (4) image -> token cross attention, updates the keys (:175-180)
            ia = layer.cross_attn_image_to_token
            wk, bk = self._w32(k + ".i2t.k", ia.k_proj); wv, bv = self._w32(k + ".i2t.v", ia.v_proj)
            kt = ops.small_linear(q_in, wk, bk); vt = ops.small_linear(queries, wv, bv)
            qi = shared["qi"] if li == 0 else self._image_proj(k + ".i2t.q", ia.q_proj, keys, pe, N)
            ai = torch.empty(B * N, ia.internal_dim, device=tok.device, dtype=torch.bfloat16)
            ops.i2t_attention(qi, kt, vt, src_of, ai, B, T, N, H, ia.internal_dim // H)
            wo, bo = self._w16(k + ".i2t.o", ia.out_proj)
            delta = torch.empty(B * N, C, device=tok.device, dtype=torch.float32)
            ops.gemm(ai, wo, delta, bias=bo)
            g, b = self._ln(k + ".n4", layer.norm4)
            new_keys = torch.empty(B * N, C, device=tok.device, dtype=torch.bfloat16)
            ops.keys_add_ln(keys, src_of, delta, g, b, new_keys, B, N, C, eps=layer.norm4.eps)
            keys, src_of = new_keys, None
            del delta, ai
        # final token -> image attention (transformer.py:99-104)
        fa = tr.final_attn_token_to_image
        wq, bq = self._w32("f.q", fa.q_proj)
        q = ops.small_linear(q_in, wq, bq)
        kp = self._image_proj("f.k", fa.k_proj, keys, pe, N)
        vp = self._image_proj("f.v", fa.v_proj, keys, None, N)
        att = ops.t2i_attention(q, kp, vp, src_of, B, T, N, H, fa.internal_dim // H).reshape(R, fa.internal_dim)
        # final out-proj + norm_final_attn + heads for token 1 + num_mask_tokens (mask_decoder.py:191-203) -> packed record
        wo, bo = self._w32("f.o", fa.out_proj)
        g, b = self._ln("f.n", tr.norm_final_attn)
        w0, b0 = self._w32("h.0", self.bbox_prediction_head[0]); w2, b2 = self._w32("h.2", self.bbox_prediction_head[2])
        wt, bt = self._w32("h.t", self.temporal_objectness_head) if self.use_temp_objectness else (None, None)
        return ops.decoder_heads(queries.view(B, T, C), att.view(B, T, fa.internal_dim), wo, bo, g, b, tr.norm_final_attn.eps, w0, b0, w2, b2,
                                 wt, bt, records, tok=1 + self.num_mask_tokens)
