"""Factories with the reference's names and hyper-parameters (model/SAM/build_sam.py:15-113).

`build_sam_vit_h(checkpoint, use_temp_objectness)` returns an object with `.image_encoder`, `.prompt_encoder`,
`.mask_decoder` (GROVE.py:55).  The reference hard-wires img_size=1024 for the encoder and a 512-pixel prompt
encoder (build_sam.py:66-69) and lets train.py:561-576 resize the positional tables; `image_size` here sets both
consistently for callers that want a self-consistent model at another resolution (default = reference values).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .image_encoder import ImageEncoderViT
from .mask_decoder import MaskDecoder
from .prompt_encoder import PromptEncoder
from .transformer import TwoWayTransformer


class Sam(nn.Module):
    """Parameter container (model/SAM/modeling/sam.py:18; GROVE only uses it as such, GROVE.py:55)."""
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder, prompt_encoder, mask_decoder, pixel_mean=(123.675, 116.28, 103.53), pixel_std=(58.395, 57.12, 57.375)):
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)


def _build_sam(encoder_embed_dim, encoder_depth, encoder_num_heads, encoder_global_attn_indexes, checkpoint=None, use_temp_objectness=True,
               image_size=None):
    prompt_embed_dim = 256
    enc_size = 1024 if image_size is None else image_size
    pe_size = 512 if image_size is None else image_size      # build_sam.py:66-69
    vit_patch_size = 16
    emb = pe_size // vit_patch_size
    sam = Sam(
        image_encoder=ImageEncoderViT(depth=encoder_depth, embed_dim=encoder_embed_dim, img_size=enc_size, mlp_ratio=4,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=encoder_num_heads, patch_size=vit_patch_size,
                                      qkv_bias=True, use_rel_pos=True, global_attn_indexes=encoder_global_attn_indexes, window_size=14,
                                      out_chans=prompt_embed_dim),
        prompt_encoder=PromptEncoder(embed_dim=prompt_embed_dim, image_embedding_size=(emb, emb), input_image_size=(pe_size, pe_size),
                                     mask_in_chans=16),
        mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                 transformer=TwoWayTransformer(depth=2, embedding_dim=prompt_embed_dim, mlp_dim=2048, num_heads=8),
                                 transformer_dim=prompt_embed_dim, iou_head_depth=3, iou_head_hidden_dim=256, decoding_type="query",
                                 use_temp_objectness=use_temp_objectness),
    )
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            state_dict = torch.load(f)
        sam.load_state_dict(state_dict, strict=False, assign=True)
    return sam


def build_sam_vit_h(checkpoint=None, use_temp_objectness=True, image_size=None):
    return _build_sam(1280, 32, 16, [7, 15, 23, 31], checkpoint, use_temp_objectness, image_size)


build_sam = build_sam_vit_h


def build_sam_vit_l(checkpoint=None, use_temp_objectness=True, image_size=None):
    return _build_sam(1024, 24, 16, [5, 11, 17, 23], checkpoint, use_temp_objectness, image_size)


def build_sam_vit_b(checkpoint=None, use_temp_objectness=True, image_size=None):
    return _build_sam(768, 12, 12, [2, 5, 8, 11], checkpoint, use_temp_objectness, image_size)


sam_model_registry = {"default": build_sam_vit_h, "vit_h": build_sam_vit_h, "vit_l": build_sam_vit_l, "vit_b": build_sam_vit_b}
