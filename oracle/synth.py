"""TEST INFRASTRUCTURE ONLY — deterministic synthetic weights and inputs.

Weights are drawn from numpy's PCG64 (stable across platforms and numpy versions), keyed by the
*reference's parameter names*, so the build container (where the reference runs) and the GPU box
(where it does not exist) construct bit-identical state dicts without shipping them.  The name /
shape tables below are checked against the reference modules' own ``state_dict()`` by
``oracle/make_golden.py``.  Zero-initialised reference parameters (pos_embed, rel_pos_*, adapter
alpha) are deliberately given non-zero values so parity is not vacuous.
"""
from __future__ import annotations

import zlib
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

Shapes = Dict[str, Tuple[int, ...]]


def encoder_param_shapes(embed_dim: int, depth: int, heads: int, global_idx: Sequence[int], grid: int,
                         window: int = 14, out_chans: int = 256, pre: str = "image_encoder.") -> Shapes:
    """Parameter names/shapes of ImageEncoderViT (model/SAM/modeling/image_encoder.py:63-170)."""
    D, hd = embed_dim, embed_dim // heads
    s: Shapes = {pre + "pos_embed": (1, grid, grid, D),
                 pre + "patch_embed.proj.weight": (D, 3, 16, 16), pre + "patch_embed.proj.bias": (D,)}
    for i in range(depth):
        b = f"{pre}blocks.{i}."
        S = grid if i in global_idx else window
        s.update({b + "norm1.weight": (D,), b + "norm1.bias": (D,),
                  b + "attn.rel_pos_h": (2 * S - 1, hd), b + "attn.rel_pos_w": (2 * S - 1, hd),
                  b + "attn.qkv.weight": (3 * D, D), b + "attn.qkv.bias": (3 * D,),
                  b + "attn.proj.weight": (D, D), b + "attn.proj.bias": (D,),
                  b + "norm2.weight": (D,), b + "norm2.bias": (D,),
                  b + "mlp.lin1.weight": (4 * D, D), b + "mlp.lin1.bias": (4 * D,),
                  b + "mlp.lin2.weight": (D, 4 * D), b + "mlp.lin2.bias": (D,)})
    for k in range(len(global_idx)):
        a = f"{pre}adapters.{k}."
        s.update({a + "alpha": (1,), a + "conv3d.weight": (D, D, 3, 3, 3), a + "conv3d.bias": (D,)})
    s.update({pre + "neck.0.weight": (out_chans, D, 1, 1), pre + "neck.1.weight": (out_chans,), pre + "neck.1.bias": (out_chans,),
              pre + "neck.2.weight": (out_chans, out_chans, 3, 3), pre + "neck.3.weight": (out_chans,), pre + "neck.3.bias": (out_chans,)})
    return s


def decoder_param_shapes(dim: int = 256, mlp: int = 2048, depth: int = 2,
                         pe_pre: str = "prompt_encoder.", md_pre: str = "mask_decoder.") -> Shapes:
    """The parameters the query path touches (prompt_encoder.py:65,198-201; mask_decoder.py:53-55,80-85;
    transformer.py:45-60,134-149,205-208)."""
    s: Shapes = {pe_pre + "pe_layer.positional_encoding_gaussian_matrix": (2, dim // 2),
                 pe_pre + "no_mask_embed.weight": (1, dim),
                 md_pre + "iou_token.weight": (1, dim), md_pre + "mask_tokens.weight": (4, dim)}

    def attn(p, internal):
        for n in ("q_proj", "k_proj", "v_proj"):
            s[p + n + ".weight"] = (internal, dim)
            s[p + n + ".bias"] = (internal,)
        s[p + "out_proj.weight"] = (dim, internal)
        s[p + "out_proj.bias"] = (dim,)

    t = md_pre + "transformer."
    for i in range(depth):
        l = f"{t}layers.{i}."
        attn(l + "self_attn.", dim)
        attn(l + "cross_attn_token_to_image.", dim // 2)
        attn(l + "cross_attn_image_to_token.", dim // 2)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            s[l + n + ".weight"] = (dim,)
            s[l + n + ".bias"] = (dim,)
        s.update({l + "mlp.lin1.weight": (mlp, dim), l + "mlp.lin1.bias": (mlp,),
                  l + "mlp.lin2.weight": (dim, mlp), l + "mlp.lin2.bias": (dim,)})
    attn(t + "final_attn_token_to_image.", dim // 2)
    s.update({t + "norm_final_attn.weight": (dim,), t + "norm_final_attn.bias": (dim,),
              md_pre + "bbox_prediction_head.0.weight": (dim, dim), md_pre + "bbox_prediction_head.0.bias": (dim,),
              md_pre + "bbox_prediction_head.2.weight": (4, dim), md_pre + "bbox_prediction_head.2.bias": (4,),
              md_pre + "temporal_objectness_head.weight": (1, dim), md_pre + "temporal_objectness_head.bias": (1,)})
    return s


def text_fcs_shapes(hidden: int = 4096, out_dim: int = 256, pre: str = "text_hidden_fcs.0.") -> Shapes:
    """GROVE.py:75-79."""
    return {pre + "0.weight": (hidden, hidden), pre + "0.bias": (hidden,),
            pre + "2.weight": (out_dim, hidden), pre + "2.bias": (out_dim,)}


def _draw(name: str, shape, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    n = rng.standard_normal(shape, dtype=np.float64)
    leaf = name.rsplit(".", 1)[-1]
    if "norm" in name or ".neck.1." in name or ".neck.3." in name:
        return (1.0 + 0.1 * n) if leaf == "weight" else 0.05 * n
    if leaf == "alpha":
        return np.full(shape, 0.5)
    if leaf == "pos_embed":
        return 0.02 * n
    if leaf in ("rel_pos_h", "rel_pos_w"):
        return 0.05 * n
    if any(t in name for t in ("positional_encoding_gaussian_matrix", "iou_token.", "mask_tokens.", "no_mask_embed.")):
        return n  # N(0,1) like nn.Embedding / randn buffers
    if leaf == "bias":
        return 0.02 * n
    fan_in = int(np.prod(shape[1:]))
    return n / np.sqrt(3.0 * fan_in)  # same variance as torch's default kaiming_uniform(a=sqrt(5)) init


def synth_state_dict(shapes: Shapes, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(_draw(k, v, seed)).to(dtype) for k, v in shapes.items()}


def synth_tensor(name: str, shape, seed: int = 0, scale: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode()), 7]))
    return torch.from_numpy(scale * rng.standard_normal(shape, dtype=np.float64)).to(dtype)


def synth_uniform(name: str, shape, lo: float, hi: float, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode()), 11]))
    return torch.from_numpy(rng.uniform(lo, hi, shape)).to(dtype)


def det_positions(L: int, n_det: int, seed: int = 0):
    """Fixed [DET] token positions >= 575."""
    rng = np.random.Generator(np.random.PCG64([seed, 1234]))
    return sorted(int(p) for p in rng.choice(np.arange(576, L - 1), size=n_det, replace=False))
