"""the two video-specific pieces GROVE adds to the CLIP global encoder, on the same CUDA library as the grounding path.

* `SpatioTemporalConvAdapter` -- model/llava/model/multimodal_encoder/modeling_clip.py:591-612: applied after every third CLIP encoder
  layer (:705-708) to the layer's output tuple; the cls token passes through, the 16x16 patch tokens of 8 consecutive frames go through
  `tanh(alpha) * relu(Conv3d(D, D, 3x3x3, 'same')) + x`.  Same arithmetic as the SAM adapter, so it runs on the same implicit-GEMM
  tcgen05 kernel (27 shifted 5-D TMA boxes, zero fill = 'same' padding, gate and residual fused in the epilogue).
* `AdaptiveAvgPooling3D` -- pooling.py:6-25: '(b t) (h w) c' video features -> `nn.AdaptiveAvgPool3d((t, 8, 9))` -> 'b (t h w) c'
  (576 tokens for 8 frames); one HBM-bound kernel with both rearranges folded into its indexing.

Same constructor signatures and parameter names as the reference classes (`conv3d.weight`, `conv3d.bias`, `alpha`).  CUDA only."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .modeling.common import PackCache, bf16, f32


class SpatioTemporalConvAdapter(nn.Module):
    """modeling_clip.py:591-612.  forward(x: tuple) -> tuple, x[0] = [(b t), 1 + h*w, c] with t = 8 and h = w = 16 as the reference
    hard-codes (:603)."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.conv3d = nn.Conv3d(in_channels, out_channels, kernel_size, padding="same")
        self.relu = nn.ReLU()
        self.alpha = nn.Parameter(torch.zeros([1]))
        self.tanh = nn.Tanh()
        self._pack = PackCache()

    @torch.no_grad()
    def forward(self, x):
        inp = x[0]
        if not inp.is_cuda:
            raise RuntimeError("grove_b200.clip_adapters runs on CUDA only (no CPU fallback)")
        c3 = self.conv3d
        D = c3.in_channels
        if tuple(c3.kernel_size) != (3, 3, 3) or c3.out_channels != D:
            raise NotImplementedError("the CLIP adapter is Conv3d(D, D, 3x3x3, 'same') (modeling_clip.py:594)")
        BT, L, C = inp.shape
        if L != 1 + 256 or C != D or BT % 8:
            raise ValueError(f"expected [(b*8), 1 + 16*16, {D}] (modeling_clip.py:603 fixes t=8, h=16), got {tuple(inp.shape)}")
        cls_embed = inp[:, :1]
        seq = inp[:, 1:].contiguous()
        xb = seq.to(torch.bfloat16).reshape(BT * 256, D)
        xs = seq.to(torch.float32).reshape(BT * 256, D)                 # fp32 residual / output of the fused epilogue
        wc = self._pack.get("w", [c3.weight], lambda w: bf16(w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)))
        bc = self._pack.get("b", [c3.bias], f32)
        al = self._pack.get("alpha", [self.alpha], f32)
        ops.conv_gemm(xb, wc, xs, V=BT // 8, T=8, G=16, kt=3, bias=bc, act="relu", gate_alpha=al, resid=xs)
        out = torch.cat((cls_embed, xs.view(BT, 256, D).to(inp.dtype)), dim=1)
        return (out,)


class AdaptiveAvgPooling3D(nn.Module):
    """pooling.py:6-25 (no parameters)."""

    def __init__(self, num_frames=8, output_tokens=576):
        super().__init__()
        self.num_frames = num_frames
        self.output_size = (num_frames, 8, 9)                           # pooling.py:13 (output_tokens is unused there too)

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("grove_b200.clip_adapters runs on CUDA only (no CPU fallback)")
        BT, N, C = x.shape
        h = w = int(N ** 0.5)                                           # pooling.py:18
        T = self.num_frames
        if h * w != N or BT % T:
            raise ValueError(f"expected [(b*{T}), h*w, c] with square h*w, got {tuple(x.shape)}")
        xin = x.contiguous()
        if xin.dtype not in (torch.bfloat16, torch.float32):
            xin = xin.float()
        OT, OH, OW = self.output_size
        out = torch.empty(BT // T, OT * OH * OW, C, device=x.device, dtype=xin.dtype)
        ops.adaptive_avgpool3d_tokens(xin, out, B=BT // T, T=T, H=h, W=w, OT=OT, OH=OH, OW=OW)
        return out.to(x.dtype)
