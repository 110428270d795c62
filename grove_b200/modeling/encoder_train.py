"""Training step of the SAM ViT encoder with trainable Conv3d adapters (BASELINE config 4).

In GROVE the ViT blocks are frozen and the four SpatioTemporalConvAdapters train (train.py:254-255, 279-280), but the encoder is
NOT under no_grad (GROVE.py:134-136): autograd walks back through every block after the first adapter.  `encode_train` is
ImageEncoderViT.forward_tokens with the activations of those blocks kept (fresh buffers instead of the in-place residual stream),
`encode_backward` is the explicit reverse schedule:

  neck           LayerNorm-bwd -> implicit-GEMM 3x3 with flipped taps -> LayerNorm-bwd -> GEMM with W^T
  block          GEMM(W2^T) * GELU'(pre) -> GEMM(W1^T) -> LayerNorm-bwd -> GEMM(Wproj^T) -> attention-bwd -> GEMM(Wqkv^T) -> LayerNorm-bwd
  adapter        gate/ReLU mask + d alpha + d bias -> Conv3d weight gradient (tcgen05 GEMM over tokens) -> Conv3d input gradient
                 (implicit GEMM with flipped, transposed taps), skipped for the first adapter (nothing trainable precedes it)

Frozen parameters get no gradient; the walk stops right after the first adapter's weight gradient.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .common import bf16, f32
from .decoder_train import GradStore, _wants


def encode_train(enc, x: torch.Tensor):
    """[V,3,T,H,W] -> (token-major embeddings [V*T, G*G, out_chans] bf16, tape)"""
    from .image_encoder import _resize_rel_pos, is_conv_adapter
    if x.dim() != 5 or x.shape[1] != 3 or not x.is_cuda:
        raise ValueError("expected CUDA images of shape [V,3,T,H,W]")
    V, _, T, H, W = x.shape
    Fr, G, D, heads = V * T, H // 16, enc.embed_dim, enc.num_heads
    hd, N = D // heads, (H // 16) ** 2
    if Fr % 8:
        raise ValueError("the spatio-temporal adapter groups frames by 8 (image_encoder.py:52)")
    dev = x.device
    M = Fr * N
    gi = list(enc.global_attn_indexes)
    convs = [is_conv_adapter(a) for a in enc.adapters]
    trainable = [c and any(p.requires_grad for p in a.parameters()) for c, a in zip(convs, enc.adapters)]
    first = next((gi[k] for k in range(len(gi)) if trainable[k]), None)      # block index followed by the first trainable adapter
    tape = {"Fr": Fr, "G": G, "N": N, "M": M, "blocks": {}, "adapters": {}, "first": first}

    img = x.to(torch.bfloat16).contiguous()
    patches = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    ops.im2col_patch16(img, patches)
    xs = torch.empty(M, D, device=dev, dtype=torch.float32)
    wpe = enc._pack.get("pe.w", [enc.patch_embed.proj.weight], lambda w: bf16(w.reshape(w.shape[0], -1)))
    bpe = enc._pack.get("pe.b", [enc.patch_embed.proj.bias], f32)
    pos = None
    if enc.pos_embed is not None:
        pos = enc._pack.get("pos", [enc.pos_embed], lambda p: f32(p.reshape(N, D)))
    ops.gemm(patches, wpe, xs, bias=bpe, resid=pos, resid_row_mod=N if pos is not None else 0)
    del patches

    h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mlp_dim = enc.blocks[0].mlp.lin1.out_features
    hid = torch.empty(M, mlp_dim, device=dev, dtype=torch.bfloat16)
    scratch = {"qkv": torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16), "att": torch.empty(M, D, device=dev, dtype=torch.bfloat16)}
    xb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    last = len(enc.blocks) - 1
    for i, blk in enumerate(enc.blocks):
        k = f"b{i}"
        keep = first is not None and i > first          # this block lies on the backward path
        g1, b1 = enc._ln(k + ".n1", blk.norm1)
        ops.layernorm(xs, g1, b1, h, blk.norm1.eps)
        wq, bq = enc._linear(k + ".qkv", blk.attn.qkv)
        qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16) if keep else scratch["qkv"]
        att = torch.empty(M, D, device=dev, dtype=torch.bfloat16) if keep else scratch["att"]
        ops.gemm(h, wq, qkv, bias=bq)
        S = blk.window_size if blk.window_size > 0 else G
        lse = None
        if blk.window_size > 0:
            bqb = enc._pack.get(k + ".qkvb16", [blk.attn.qkv.bias], bf16)
            tab = enc._pack.get(k + ".reltab", [blk.attn.rel_pos_h, blk.attn.rel_pos_w],
                                lambda a, b, S=S: ops.window_rel_table(_resize_rel_pos(a, S), _resize_rel_pos(b, S)))
            lse = torch.empty(M, heads, device=dev, dtype=torch.float32) if keep else None   # row log-sum-exp, reused by the backward pass
            ops.attn_window_tc(qkv, bqb, tab, att, F=Fr, G=G, heads=heads, hd=hd, ws=blk.window_size, lse=lse)
        else:
            rh = enc._pack.get(k + ".rh", [blk.attn.rel_pos_h], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
            rw = enc._pack.get(k + ".rw", [blk.attn.rel_pos_w], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
            lse = torch.empty(M, heads, device=dev, dtype=torch.float32) if keep else None   # row log-sum-exp, reused by the backward pass
            ops.attn_global(qkv, rh, rw, att, F=Fr, G=G, heads=heads, hd=hd, lse=lse)
        wp, bp = enc._linear(k + ".proj", blk.attn.proj)
        x1 = torch.empty(M, D, device=dev, dtype=torch.float32) if keep else xs
        ops.gemm(att, wp, x1, bias=bp, resid=xs)
        g2, b2 = enc._ln(k + ".n2", blk.norm2)
        ops.layernorm(x1, g2, b2, h, blk.norm2.eps)
        w1, bb1 = enc._linear(k + ".l1", blk.mlp.lin1)
        pre = torch.empty(M, mlp_dim, device=dev, dtype=torch.bfloat16) if keep else None
        ops.gemm(h, w1, hid, bias=bb1, act="gelu", out2=pre, out2_pre_act=1 if keep else 0)
        w2, bb2 = enc._linear(k + ".l2", blk.mlp.lin2)
        adapter = enc.adapters[gi.index(i)] if i in gi else None
        conv = is_conv_adapter(adapter)
        x2 = torch.empty(M, D, device=dev, dtype=torch.float32) if keep else x1
        ops.gemm(hid, w2, x2, bias=bb2, resid=x1, out2=xb if (conv or i == last) else None)
        if keep:
            tape["blocks"][i] = {"x0": xs, "x1": x1, "qkv": qkv, "att": att, "pre": pre, "lse": lse}
        xs = x2
        if conv:
            kk = gi.index(i)
            c3 = adapter.conv3d
            wc = enc._pack.get(k + ".c3w", [c3.weight], lambda w: bf16(w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)))
            bc = enc._pack.get(k + ".c3b", [c3.bias], f32)
            al = enc._pack.get(k + ".alpha", [adapter.alpha], f32)
            on_path = first is not None and i >= first
            y = torch.empty(M, D, device=dev, dtype=torch.float32) if on_path else xs
            if on_path:
                r = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
                ops.conv_gemm(xb, wc, y, V=Fr // 8, T=8, G=G, kt=3, bias=bc, act="relu", gate_alpha=al, resid=xs, out2=r, out2_pre_act=2)
                tape["adapters"][kk] = {"x_in": xb, "relu": r, "block": i}
                xb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
                if i == last:
                    ops.cast_f32_bf16(y.view(-1), xb.view(-1))
            else:
                xb2 = torch.empty(M, D, device=dev, dtype=torch.bfloat16) if i == last else None
                ops.conv_gemm(xb, wc, y, V=Fr // 8, T=8, G=G, kt=3, bias=bc, act="relu", gate_alpha=al, resid=xs, out2=xb2)
                if i == last:
                    xb = xb2
            xs = y
    # neck
    C = enc.out_chans
    wn0 = enc._pack.get("n0", [enc.neck[0].weight], lambda w: bf16(w.reshape(w.shape[0], -1)))
    y0a = torch.empty(M, C, device=dev, dtype=torch.float32)
    ops.gemm(xb, wn0, y0a)
    g, b = enc._ln("n1", enc.neck[1])
    y1 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.layernorm(y0a, g, b, y1, enc.neck[1].eps)
    wn2 = enc._pack.get("n2", [enc.neck[2].weight], lambda w: bf16(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)))
    y0b = torch.empty(M, C, device=dev, dtype=torch.float32)
    ops.conv_gemm(y1, wn2, y0b, V=Fr, T=1, G=G, kt=1)
    g, b = enc._ln("n3", enc.neck[3])
    emb = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.layernorm(y0b, g, b, emb, enc.neck[3].eps)
    tape["y0a"], tape["y0b"] = y0a, y0b
    return emb.view(Fr, N, C), tape


def encode_backward(enc, tape, d_emb: torch.Tensor, grads: GradStore, on_ready=None) -> None:
    """d_emb: fp32 [M, out_chans] cotangent of the token-major embeddings.  Accumulates the adapters' gradients into `grads`.
    `on_ready(list of fp32 accumulators)` is called as soon as an adapter's gradients are final (its weight gradient is the last thing the
    walk adds to them), so a data-parallel reducer can overlap their all-reduce with the rest of the walk."""
    from .image_encoder import _resize_rel_pos
    first = tape["first"]
    if first is None:
        return
    Fr, G, N, M = tape["Fr"], tape["G"], tape["N"], tape["M"]
    D, heads, C = enc.embed_dim, enc.num_heads, enc.out_chans
    hd = D // heads
    dev = d_emb.device
    gi = list(enc.global_attn_indexes)
    # ---- neck (image_encoder.py:152-168), frozen: input gradients only
    d16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.layernorm_bwd(tape["y0b"], f32(enc.neck[3].weight), d_emb, eps=enc.neck[3].eps, dx_bf16=d16)
    wn2f = enc._pack.get("n2.flipT", [enc.neck[2].weight], lambda w: bf16(w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1)))
    d_y1 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.conv_gemm(d16, wn2f, d_y1, V=Fr, T=1, G=G, kt=1)
    ops.layernorm_bwd(tape["y0a"], f32(enc.neck[1].weight), d_y1, eps=enc.neck[1].eps, dx_bf16=d16)
    wn0t = enc._pack.get("n0.T", [enc.neck[0].weight], lambda w: bf16(w.reshape(w.shape[0], -1).t()))
    dxs = torch.empty(M, D, device=dev, dtype=torch.float32)       # cotangent of the residual stream
    g16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)      # its bf16 copy (GEMM operand)
    ops.gemm(d16, wn0t, dxs, out2=g16)
    del d16, d_y1
    mlp_dim = enc.blocks[0].mlp.lin1.out_features
    dhid = torch.empty(M, mlp_dim, device=dev, dtype=torch.bfloat16)
    dh = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    dyc = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    for i in range(len(enc.blocks) - 1, first - 1, -1):
        blk = enc.blocks[i]
        k = f"b{i}"
        if i in gi and gi.index(i) in tape["adapters"]:
            kk = gi.index(i)
            ad, A = enc.adapters[kk], tape["adapters"][kk]
            c3 = ad.conv3d
            al = enc._pack.get(k + ".alpha", [ad.alpha], f32)
            ops.adapter_gate_bwd(dxs, A["relu"], al, dyc, grads.buf(c3.bias), grads.buf(ad.alpha))
            if _wants(c3.weight):
                gw = grads.buf(c3.weight, shape=(D, 27 * D), unpack=lambda g, w=c3.weight: unpack_conv3d_grad(g, w))   # tap-major packed layout
                ops.conv_wgrad(dyc, A["x_in"], gw, V=Fr // 8, T=8, G=G, kt=3)
            if on_ready is not None:
                on_ready([grads.g[p] for p in (c3.weight, c3.bias, ad.alpha) if p in grads.g])
            if i == first:
                break                                               # nothing trainable precedes the first adapter
            wcf = enc._pack.get(k + ".c3w.flipT", [c3.weight], lambda w: bf16(w.flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(w.shape[1], -1)))
            ops.conv_gemm(dyc, wcf, dxs, V=Fr // 8, T=8, G=G, kt=3, resid=dxs, out2=g16)
        B_ = tape["blocks"][i]
        # MLP: x2 = x1 + lin2(gelu(lin1(norm2(x1))))
        w2t = enc._pack.get(k + ".l2.T", [blk.mlp.lin2.weight], lambda w: bf16(w.t()))
        ops.gemm(g16, w2t, dhid, dact_pre=B_["pre"], dact="gelu")
        w1t = enc._pack.get(k + ".l1.T", [blk.mlp.lin1.weight], lambda w: bf16(w.t()))
        ops.gemm(dhid, w1t, dh)
        ops.layernorm_bwd(B_["x1"], f32(blk.norm2.weight), dh, eps=blk.norm2.eps, dx_in=dxs, dx_out=dxs, dx_bf16=g16)
        # attention: x1 = x0 + proj(attn(qkv(norm1(x0))))
        wpt = enc._pack.get(k + ".proj.T", [blk.attn.proj.weight], lambda w: bf16(w.t()))
        ops.gemm(g16, wpt, dh)                                       # d att
        S = blk.window_size if blk.window_size > 0 else G
        rh = enc._pack.get(k + ".rh", [blk.attn.rel_pos_h], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
        rw = enc._pack.get(k + ".rw", [blk.attn.rel_pos_w], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
        bqb = enc._pack.get(k + ".qkvb16", [blk.attn.qkv.bias], bf16)
        ops.attn_relpos_bwd(B_["qkv"], bqb if blk.window_size > 0 else None, rh, rw, B_["att"], dh, dqkv, F=Fr, G=G, heads=heads, hd=hd,
                            ws=blk.window_size, lse=B_["lse"])
        wqt = enc._pack.get(k + ".qkv.T", [blk.attn.qkv.weight], lambda w: bf16(w.t()))
        ops.gemm(dqkv, wqt, dh)
        ops.layernorm_bwd(B_["x0"], f32(blk.norm1.weight), dh, eps=blk.norm1.eps, dx_in=dxs, dx_out=dxs, dx_bf16=g16)
        tape["blocks"][i] = None                                      # release the activations as soon as they are consumed


def unpack_conv3d_grad(g: torch.Tensor, weight: nn.Parameter) -> torch.Tensor:
    """tap-major [N, (kd,kh,kw,C)] -> Conv3d layout [N, C, kd, kh, kw]"""
    N, C = weight.shape[0], weight.shape[1]
    return g.view(N, 3, 3, 3, C).permute(0, 4, 1, 2, 3).contiguous()
