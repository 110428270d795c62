"""Prompt encoder — text-prompt path (reference: model/SAM/modeling/prompt_encoder.py).

GROVE only ever calls `prompt_encoder(points=None, boxes=None, masks=None, text_embeds=...)` and
`get_dense_pe()` (GROVE.py:182,275-277); point / box / mask prompts are out of scope and raise.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple, Type

import torch
import torch.nn as nn

from .. import ops
from .common import LayerNorm2d, PackCache, _ContainerOnly, f32


class PositionEmbeddingRandom(_ContainerOnly):
    """prompt_encoder.py:189-238 — random-Fourier positional encoding; the grid encoding runs in grove_dense_pe."""

    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def grid_tokens(self, G: int) -> torch.Tensor:
        """token-major fp32 [G*G, 2*num_pos_feats]; constant per (matrix, grid) -> computed once and reused (the reference
        recomputes it every training step, GROVE.py:182, and hoists it in inference, infer_iground.py:157)."""
        if not hasattr(self, "_pack"):
            self._pack = PackCache()
        return self._pack.get(("pe", G), [self.positional_encoding_gaussian_matrix], lambda g: ops.dense_pe(f32(g), G))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int], mask_in_chans: int,
                 activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        # parameters of the unused prompt types are kept so reference checkpoints load key-for-key (prompt_encoder.py:45-65)
        self.num_point_embeddings: int = 4
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans), activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self) -> torch.Tensor:
        """prompt_encoder.py:67-76 -> [1, embed_dim, G, G] (channels-last view of the token-major table)."""
        G = self.image_embedding_size[0]
        if self.image_embedding_size[1] != G:
            raise NotImplementedError("square embedding grids only")
        pe = self.pe_layer.grid_tokens(G)
        out = pe.view(1, G, G, self.embed_dim).permute(0, 3, 1, 2)   # a view: MaskDecoder recovers the cached token-major table
        dt = self.pe_layer.positional_encoding_gaussian_matrix.dtype
        return out if dt == torch.float32 else out.to(dt)

    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor], masks: Optional[torch.Tensor],
                text_embeds: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """prompt_encoder.py:140-186, text path: sparse = text_embeds (returned in fp32 like the reference's concat
        with an fp32 empty tensor, :164-167,176-177); dense = no_mask_embed broadcast as a VIEW (:182-184)."""
        if points is not None or boxes is not None or masks is not None:
            raise NotImplementedError("grove_b200 implements the text-prompt path only (GROVE.py:275-277 passes None for the rest)")
        if text_embeds is None:
            raise ValueError("text_embeds is required")
        bs = text_embeds.shape[0]
        sparse = text_embeds.to(torch.float32)
        dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense
