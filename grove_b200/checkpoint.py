"""Checkpoint compatibility tooling: run real SAM / GROVE checkpoints through grove_b200 at another resolution.

Load-time model surgery, executed once per model (never on the hot path), therefore plain torch on whatever device the parameters
live on: the reference does exactly this in train.py:503-576 / infer_iground.py (same calls, same align_corners quirks):

  pos_embed     [1, G, G, D]      bicubic, align_corners=False                                   (train.py:503-529)
  rel_pos_h/w   [2G-1, hd]        bicubic over the table axis, align_corners=True, GLOBAL blocks only (train.py:532-558, 568-574)
  img_size      attribute         (train.py:576)

`load_sam_state_dict` maps a reference state_dict (SAM `image_encoder.* / prompt_encoder.* / mask_decoder.*` keys or GROVE
`model.grounding_encoder.*` / `model.text_hidden_fcs.*` keys) onto the grove_b200 modules, which keep the reference's names.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def resize_abs_pos_embedding(pos_embed: torch.Tensor, target_size: int, patch_size: int = 16) -> torch.Tensor:
    """train.py:503-529"""
    n = target_size // patch_size
    return F.interpolate(pos_embed.permute(0, 3, 1, 2), size=(n, n), mode="bicubic", align_corners=False).permute(0, 2, 3, 1)


def resize_rel_pos_embedding(rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, target_size: int, patch_size: int = 16):
    """train.py:532-558: both tables to 2 * (target_size / patch) - 1 rows"""
    n = 2 * (target_size // patch_size) - 1
    h = F.interpolate(rel_pos_h[None, None].permute(0, 3, 2, 1), size=(n, 1), mode="bicubic", align_corners=True)
    w = F.interpolate(rel_pos_w[None, None].permute(0, 3, 1, 2), size=(1, n), mode="bicubic", align_corners=True)
    return h.permute(0, 3, 2, 1)[0, 0], w.permute(0, 2, 3, 1)[0, 0]


@torch.no_grad()
def interpolate_positional_embeddings(image_encoder, target_size: int = 512, patch_size: int = 16,
                                      global_attn_indexes: Optional[Sequence[int]] = None) -> None:
    """train.py:561-576 on a grove_b200 (or reference) ImageEncoderViT: resizes pos_embed and the GLOBAL blocks' rel-pos tables in place
    (windowed blocks keep their 27-row tables) and sets img_size.  The reference hard-codes ViT-H's global indexes; here they default
    to the encoder's own."""
    gi = list(image_encoder.global_attn_indexes if global_attn_indexes is None else global_attn_indexes)
    if image_encoder.pos_embed is not None:
        image_encoder.pos_embed = nn.Parameter(resize_abs_pos_embedding(image_encoder.pos_embed.clone().contiguous(), target_size, patch_size).contiguous(),
                                               requires_grad=image_encoder.pos_embed.requires_grad)
    for i in gi:
        attn = image_encoder.blocks[i].attn
        h, w = resize_rel_pos_embedding(attn.rel_pos_h.clone().contiguous(), attn.rel_pos_w.clone().contiguous(), target_size, patch_size)
        attn.rel_pos_h = nn.Parameter(h.contiguous(), requires_grad=attn.rel_pos_h.requires_grad)
        attn.rel_pos_w = nn.Parameter(w.contiguous(), requires_grad=attn.rel_pos_w.requires_grad)
    image_encoder.img_size = target_size


_PREFIXES = ("model.grounding_encoder.", "grounding_encoder.", "base_model.model.model.grounding_encoder.")


def load_sam_state_dict(sam: nn.Module, state_dict: Dict[str, torch.Tensor], text_hidden_fcs: Optional[nn.Module] = None) -> Dict[str, list]:
    """strict=False load of a SAM or GROVE checkpoint into a grove_b200 Sam (and optionally text_hidden_fcs).  Returns the missing /
    unexpected / skipped key lists.  Keys of the language model, CLIP tower and projector are skipped (out of scope)."""
    own = sam.state_dict()
    picked, fcs, skipped = {}, {}, []
    for k, v in state_dict.items():
        kk = k
        for p in _PREFIXES:
            if kk.startswith(p):
                kk = kk[len(p):]
                break
        if kk in own:
            picked[kk] = v
        elif "text_hidden_fcs." in k and text_hidden_fcs is not None:
            fcs[k.split("text_hidden_fcs.", 1)[1]] = v
        else:
            skipped.append(k)
    res = sam.load_state_dict(picked, strict=False)
    out = {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys), "skipped": skipped}
    if text_hidden_fcs is not None and fcs:
        r2 = text_hidden_fcs.load_state_dict(fcs, strict=False)
        out["missing"] += ["text_hidden_fcs." + m for m in r2.missing_keys]
    return out
