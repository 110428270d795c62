"""TEST INFRASTRUCTURE ONLY — fp32 functional restatement of GROVE's grounding path.

Every function takes a flat ``state_dict`` that uses the *reference's parameter
names* (so the key-name contract of SURVEY.md §5 is exercised as well) and plain
tensors; it is device-agnostic (CPU here, ``cuda`` on the GPU box for the
full-size parity runs) and computes in whatever dtype the inputs have (fp32 in
every test).  Citations are ``file:line`` under ``/root/reference``.

Pinned by ``tests/test_oracle_golden.py`` against outputs of the reference's own
modules (``oracle/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- #
# Stage 1: SAM ViT image encoder with spatio-temporal adapters
# --------------------------------------------------------------------------- #
def layer_norm(x, w, b, eps):
    """nn.LayerNorm over the last dim (model/SAM/build_sam.py:76 sets eps=1e-6)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def layer_norm_2d(x, w, b, eps=1e-6):
    """model/SAM/modeling/common.py:31-43 — channel-first LN on [B,C,H,W]."""
    u = x.mean(1, keepdim=True)
    s = ((x - u) ** 2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def window_partition(x, ws):
    """model/SAM/modeling/image_encoder.py:329-354."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(w, ws, pad_hw, hw):
    """model/SAM/modeling/image_encoder.py:357-384."""
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // (Hp * Wp // ws // ws)
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def rel_pos_table(size, rel_pos):
    """get_rel_pos for q_size == k_size (image_encoder.py:387-417): rows (q-k)+(S-1).
    Linear interpolation of the table when its length != 2S-1 (:399-408)."""
    L = 2 * size - 1
    if rel_pos.shape[0] != L:
        rel_pos = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
        rel_pos = rel_pos.reshape(-1, L).permute(1, 0)
    idx = torch.arange(size, device=rel_pos.device)
    rel = idx[:, None] - idx[None, :] + (size - 1)
    return rel_pos[rel]  # [S(q), S(k), hd]


def vit_attention(x, sd: SD, pre: str, heads: int):
    """Attention.forward, image_encoder.py:301-326 (+ add_decomposed_rel_pos :420-458).
    The bias uses the UNSCALED q (:313-315)."""
    B, H, W, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
    qkv = qkv.reshape(B, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, H * W, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh = rel_pos_table(H, sd[pre + "rel_pos_h"])
    Rw = rel_pos_table(W, sd[pre + "rel_pos_w"])
    rq = q.reshape(B * heads, H, W, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * heads, H, W, H, W) + rel_h[..., :, None] + rel_w[..., None, :]).view(B * heads, H * W, H * W)
    attn = attn.softmax(-1)
    o = (attn @ v).view(B, heads, H, W, hd).permute(0, 2, 3, 1, 4).reshape(B, H, W, D)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def vit_block(x, sd: SD, pre: str, heads: int, window: int, eps=1e-6):
    """Block.forward, image_encoder.py:243-259."""
    sc = x
    x = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    if window > 0:
        H, W = x.shape[1], x.shape[2]
        x, pad_hw = window_partition(x, window)
    x = vit_attention(x, sd, pre + "attn.", heads)
    if window > 0:
        x = window_unpartition(x, window, pad_hw, (H, W))
    x = sc + x
    y = layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    y = F.linear(y, sd[pre + "mlp.lin1.weight"], sd[pre + "mlp.lin1.bias"])
    y = F.gelu(y)  # exact erf GELU (common.py:18)
    y = F.linear(y, sd[pre + "mlp.lin2.weight"], sd[pre + "mlp.lin2.bias"])
    return x + y


def conv_adapter(x, sd: SD, pre: str, t: int = 8):
    """SpatioTemporalConvAdapter.forward, image_encoder.py:48-59, with the grid side
    inferred from the tensor instead of the hard-coded h=32 (same arithmetic;
    SURVEY.md §0.5).  x: [(b t), G, G, D]."""
    BT, G, _, D = x.shape
    xv = x.view(BT // t, t, G, G, D).permute(0, 4, 1, 2, 3)  # b c t h w
    y = F.conv3d(xv, sd[pre + "conv3d.weight"], sd[pre + "conv3d.bias"], padding=1)
    y = torch.tanh(sd[pre + "alpha"]) * F.relu(y) + xv
    return y.permute(0, 2, 3, 4, 1).reshape(BT, G, G, D)


def image_encoder(images, sd: SD, *, depth: int, heads: int, global_idx: Sequence[int], window: int = 14,
                  pre: str = "", adapters: bool = True, return_tokens: bool = False):
    """ImageEncoderViT.forward, image_encoder.py:172-191.  images: [V,3,T,H,W] -> [V*T,256,G,G]."""
    V, C, T, H, W = images.shape
    x = images.permute(0, 2, 1, 3, 4).reshape(V * T, C, H, W)
    x = F.conv2d(x, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=16)
    x = x.permute(0, 2, 3, 1)
    if pre + "pos_embed" in sd:
        x = x + sd[pre + "pos_embed"]
    for i in range(depth):
        x = vit_block(x, sd, f"{pre}blocks.{i}.", heads, 0 if i in global_idx else window)
        if adapters and i in global_idx:
            x = conv_adapter(x, sd, f"{pre}adapters.{list(global_idx).index(i)}.")
    tokens = x
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[pre + "neck.0.weight"])
    x = layer_norm_2d(x, sd[pre + "neck.1.weight"], sd[pre + "neck.1.bias"])
    x = F.conv2d(x, sd[pre + "neck.2.weight"], padding=1)
    x = layer_norm_2d(x, sd[pre + "neck.3.weight"], sd[pre + "neck.3.bias"])
    return (x, tokens) if return_tokens else x


# --------------------------------------------------------------------------- #
# Stage 2: text_hidden_fcs projection + [DET] gather
# --------------------------------------------------------------------------- #
def create_det_token_mask(input_ids, det_token_idx: int, right_pad: int = 1):
    """GROVE.py:200-205 (training, right_pad=1) / :427-430 (generate, right_pad=0)."""
    m = input_ids[:, 1:] == det_token_idx
    z = lambda n: torch.zeros((m.shape[0], n), dtype=torch.bool, device=m.device)
    return torch.cat([z(575), m] + ([z(right_pad)] if right_pad else []), dim=1)


def process_hidden_states(hidden, det_mask, sd: SD, num_frames: int, pre: str = "text_hidden_fcs.0."):
    """GROVE.py:248-268: Linear-ReLU-Linear on all L tokens, repeat per frame, gather [DET] rows."""
    h = F.linear(hidden, sd[pre + "0.weight"], sd[pre + "0.bias"])
    h = F.linear(F.relu(h), sd[pre + "2.weight"], sd[pre + "2.bias"])
    h = h.repeat_interleave(num_frames, dim=0)
    m = det_mask.repeat_interleave(num_frames, dim=0)
    pred = h[m]
    counts = m.int().sum(-1)
    off = torch.cat([counts.new_zeros(1), counts.cumsum(-1)]).tolist()
    return [pred[off[i]:off[i + 1]] for i in range(len(off) - 1)]


# --------------------------------------------------------------------------- #
# Stage 3: prompt encoder (text path) + two-way transformer
# --------------------------------------------------------------------------- #
def dense_pe(gauss, G: int):
    """PromptEncoder.get_dense_pe / PositionEmbeddingRandom.forward, prompt_encoder.py:67-76,203-229."""
    ones = torch.ones((G, G), device=gauss.device, dtype=gauss.dtype)
    y = (ones.cumsum(0) - 0.5) / G
    x = (ones.cumsum(1) - 0.5) / G
    c = 2 * torch.stack([x, y], -1) - 1
    c = 2 * math.pi * (c @ gauss)
    return torch.cat([torch.sin(c), torch.cos(c)], -1).permute(2, 0, 1).unsqueeze(0)  # [1,256,G,G]


def dec_attention(q, k, v, sd: SD, pre: str, heads: int = 8):
    """transformer.py:220-242 (scores scaled AFTER QK^T by sqrt(c_per_head))."""
    q = F.linear(q, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"])
    k = F.linear(k, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"])
    v = F.linear(v, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"])
    sep = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)
    q, k, v = sep(q), sep(k), sep(v)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), -1)
    o = (a @ v).transpose(1, 2).reshape(q.shape[0], q.shape[2], -1)
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def _ln(x, sd, pre, eps=1e-5):
    return layer_norm(x, sd[pre + "weight"], sd[pre + "bias"], eps)


def two_way_block(queries, keys, qpe, kpe, sd: SD, pre: str, skip_first_pe: bool):
    """TwoWayAttentionBlock.forward, transformer.py:151-182."""
    if skip_first_pe:
        queries = dec_attention(queries, queries, queries, sd, pre + "self_attn.")
    else:
        q = queries + qpe
        queries = queries + dec_attention(q, q, queries, sd, pre + "self_attn.")
    queries = _ln(queries, sd, pre + "norm1.")
    q, k = queries + qpe, keys + kpe
    queries = _ln(queries + dec_attention(q, k, keys, sd, pre + "cross_attn_token_to_image."), sd, pre + "norm2.")
    m = F.linear(F.relu(F.linear(queries, sd[pre + "mlp.lin1.weight"], sd[pre + "mlp.lin1.bias"])),
                 sd[pre + "mlp.lin2.weight"], sd[pre + "mlp.lin2.bias"])
    queries = _ln(queries + m, sd, pre + "norm3.")
    q, k = queries + qpe, keys + kpe
    keys = _ln(keys + dec_attention(k, q, queries, sd, pre + "cross_attn_image_to_token."), sd, pre + "norm4.")
    return queries, keys


def two_way_transformer(src, pos, tokens, sd: SD, pre: str, depth: int = 2):
    """TwoWayTransformer.forward, transformer.py:62-106.  src,pos: [B,256,G,G]; tokens [B,6,256]."""
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos.flatten(2).permute(0, 2, 1)
    queries = tokens
    for i in range(depth):
        queries, keys = two_way_block(queries, keys, tokens, kpe, sd, f"{pre}layers.{i}.", i == 0)
    q, k = queries + tokens, keys + kpe
    queries = queries + dec_attention(q, k, keys, sd, pre + "final_attn_token_to_image.")
    return _ln(queries, sd, pre + "norm_final_attn."), keys


def box_decoder(image_embeddings, image_pe, text_embeds, reps: List[int], sd: SD,
                pe_pre: str = "prompt_encoder.", md_pre: str = "mask_decoder."):
    """PromptEncoder.forward text path (prompt_encoder.py:164-186) + MaskDecoder.predict_masks
    query path (mask_decoder.py:155-205).  text_embeds [B,1,256]; returns boxes [B,4] (cxcywh in
    (0,1)) and objectness logits [B]."""
    B = text_embeds.shape[0]
    G = image_embeddings.shape[-1]
    dense = sd[pe_pre + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, G, G)
    out_tok = torch.cat([sd[md_pre + "iou_token.weight"], sd[md_pre + "mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tok.unsqueeze(0).expand(B, -1, -1), text_embeds], 1)
    idx = torch.repeat_interleave(torch.arange(image_embeddings.shape[0]), torch.tensor(reps)).to(image_embeddings.device)
    src = image_embeddings.index_select(0, idx) + dense
    pos = image_pe.expand(B, -1, -1, -1)
    hs, _ = two_way_transformer(src, pos, tokens, sd, md_pre + "transformer.")
    qo = hs[:, 5:, :]
    h = F.relu(F.linear(qo, sd[md_pre + "bbox_prediction_head.0.weight"], sd[md_pre + "bbox_prediction_head.0.bias"]))
    boxes = torch.sigmoid(F.linear(h, sd[md_pre + "bbox_prediction_head.2.weight"], sd[md_pre + "bbox_prediction_head.2.bias"])).squeeze(1)
    logits = F.linear(qo, sd[md_pre + "temporal_objectness_head.weight"], sd[md_pre + "temporal_objectness_head.bias"]).squeeze((1, 2))
    return boxes, logits


# --------------------------------------------------------------------------- #
# Stage 4: post-process, losses
# --------------------------------------------------------------------------- #
def unnormalize_bboxes(b, w, h):
    """utils/bbox_utils.py:25-44."""
    o = torch.zeros_like(b)
    o[:, 0], o[:, 1], o[:, 2], o[:, 3] = b[:, 0] * w, b[:, 1] * h, b[:, 2] * w, b[:, 3] * h
    return o


def box_cxcywh_to_xyxy(b):
    """utils/bbox_utils.py:46-62."""
    cx, cy, w, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), -1)


def box_xyxy_to_cxcywh(b):
    """utils/bbox_utils.py:64-80."""
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), -1)


def postprocess(boxes, logits, reps: List[int], num_frames: int, orig_sizes, infer: bool, thr: float = 0.5):
    """GROVE.py:297-331 slicing loop.  Returns (nested boxes, nested logits)."""
    out_b, out_l, s = [], [], 0
    for i in range(0, len(reps), num_frames):
        fb, fl = [], []
        for j in range(num_frames):
            n = reps[i + j]
            b, l = boxes[s:s + n], logits[s:s + n]
            if infer:
                w, h = orig_sizes[i // num_frames]
                b = box_cxcywh_to_xyxy(unnormalize_bboxes(b, w, h))
                b = b[torch.sigmoid(l) > thr]
            fb.append(b)
            fl.append(l)
            s += n
        out_b.append(fb)
        out_l.append(fl)
    return out_b, out_l


def giou_loss_sum(b1, b2, eps=1e-7):
    """torchvision.ops.generalized_box_iou_loss(reduction='sum') (third-party, pinned 0.20.1 in
    README.md:53; arithmetic: iou = I/(U+eps); giou = iou - (C-U)/(C+eps); I is zero unless the
    intersection has strictly positive extent).  Call sites GROVE.py:361-363."""
    b1, b2 = b1.float(), b2.float()
    x1, y1, x2, y2 = b1.unbind(-1)
    x1g, y1g, x2g, y2g = b2.unbind(-1)
    xk1, yk1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xk2, yk2 = torch.min(x2, x2g), torch.min(y2, y2g)
    inter = torch.zeros_like(x1)
    m = (yk2 > yk1) & (xk2 > xk1)
    inter[m] = (xk2[m] - xk1[m]) * (yk2[m] - yk1[m])
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + eps)
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    area_c = (xc2 - xc1) * (yc2 - yc1)
    return (1 - (iou - (area_c - union) / (area_c + eps))).sum()


def loss_components(pred_bboxes, logits, gt_bboxes, gt_obj, ce_loss, ce_w, giou_w, obj_w):
    """GROVE.py:339-381 (use_temp_objectness branch).  L1 reuses the GIoU weight (:375)."""
    dev = ce_loss.device
    ce = ce_loss * ce_w
    giou = torch.tensor(0.0, device=dev)
    l1 = torch.tensor(0.0, device=dev)
    obj = torch.tensor(0.0, device=dev)
    nb = nmax = 0
    for v in range(len(pred_bboxes)):
        for f in range(len(pred_bboxes[v])):
            pb, lg = pred_bboxes[v][f], logits[v][f]
            gb, go = gt_bboxes[v][f].to(dev), gt_obj[v][f].to(dev)
            assert gb.shape[0] == go.sum()
            if gb.shape[0] != 0:
                sel = pb[go.bool()]
                giou = giou + giou_loss_sum(box_cxcywh_to_xyxy(sel), box_cxcywh_to_xyxy(gb))
                l1 = l1 + (sel - gb).abs().sum()
            obj = obj + F.binary_cross_entropy_with_logits(lg, go, reduction="sum")
            nb += gb.shape[0]
            nmax += pb.shape[0]
    giou = giou_w * giou / (nb + 1e-8)
    l1 = giou_w * l1 / (nb + 1e-8)
    obj = obj_w * obj / (nmax + 1e-8)
    return {"loss": ce + giou + l1 + obj, "ce_loss": ce, "giou_loss": giou, "l1_loss": l1, "temp_objectness_loss": obj}


# --------------------------------------------------------------------------- #
# Full forward used by smoke()/bench/tests
# --------------------------------------------------------------------------- #
VIT_CFG = {
    "vit_b": dict(embed_dim=768, depth=12, heads=12, global_idx=(2, 5, 8, 11)),   # build_sam.py:37-46
    "vit_l": dict(embed_dim=1024, depth=24, heads=16, global_idx=(5, 11, 17, 23)),  # :26-35
    "vit_h": dict(embed_dim=1280, depth=32, heads=16, global_idx=(7, 15, 23, 31)),  # :15-24
}


def grounding_forward(images, hidden, det_mask, sd: SD, *, depth, heads, global_idx, num_frames=8, window=14):
    """model_forward's grounding half (GROVE.py:162-186) on a reference-named state dict with prefixes
    ``image_encoder.``, ``prompt_encoder.``, ``mask_decoder.``, ``text_hidden_fcs.0.``."""
    emb = image_encoder(images, sd, depth=depth, heads=heads, global_idx=global_idx, window=window, pre="image_encoder.")
    pred = process_hidden_states(hidden, det_mask, sd, num_frames)
    reps = [p.shape[0] for p in pred]
    pe = dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], emb.shape[-1])
    boxes, logits = box_decoder(emb, pe, torch.cat(pred, 0).unsqueeze(1), reps, sd)
    return emb, boxes, logits, reps
