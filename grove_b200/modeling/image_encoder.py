"""SAM ViT image encoder with spatio-temporal Conv3d adapters — B200 pipeline behind the reference API.

Mirrors model/SAM/modeling/image_encoder.py: same class names, constructor signatures, parameter names and
`forward(images[V,3,T,H,W]) -> [V*T, out_chans, H/16, W/16]` contract (GROVE.py:134-136).  The forward pass is one
fused CUDA pipeline over a token-major fp32 residual stream:

  im2col -> tcgen05 GEMM (+bias +abs-pos)                                   (PatchEmbed :461-492, :176-177)
  per block: LN -> GEMM(qkv) -> fused rel-pos attention -> GEMM(proj,+res)  (Block :243-259, Attention :301-326)
             LN -> GEMM(lin1,+GELU) -> GEMM(lin2,+res)                      (MLPBlock common.py:13-26)
  after each global block: implicit-GEMM Conv3d, +bias, ReLU, tanh(alpha) gate, +res   (:40-59, :179-182)
  neck: GEMM 1x1 -> LN -> implicit-GEMM 3x3 -> LN                           (:152-168)

Window partition/unpartition (:329-384) never materialise: the window kernel addresses the unpartitioned
tensor and synthesises the zero-padded tokens.  Forward only (round 1).
"""
from __future__ import annotations

from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from .. import ops
from .common import LayerNorm2d, MLPBlock, PackCache, _ContainerOnly, bf16, f32, no_grad_entry


class SpatioTemporalConvAdapter(_ContainerOnly):
    """image_encoder.py:40-59.  out = tanh(alpha) * relu(conv3d(x)) + x over groups of 8 frames."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.conv3d = nn.Conv3d(in_channels, out_channels, kernel_size, padding="same")
        self.relu = nn.ReLU()
        self.alpha = nn.Parameter(torch.zeros([1]))
        self.tanh = nn.Tanh()


class PatchEmbed(_ContainerOnly):
    """image_encoder.py:461-492."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans: int = 3, embed_dim: int = 768) -> None:
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class Attention(_ContainerOnly):
    """image_encoder.py:262-326."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False, rel_pos_zero_init: bool = True,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if self.use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))


class Block(_ContainerOnly):
    """image_encoder.py:194-259."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True, norm_layer: Type[nn.Module] = nn.LayerNorm,
                 act_layer: Type[nn.Module] = nn.GELU, use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos, rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size


def is_conv_adapter(m) -> bool:
    """Any module exposing `.conv3d` (an nn.Conv3d) and `.alpha` is run as the spatio-temporal Conv3d adapter: train.py:170-176
    replaces `adapters` with fresh instances of the REFERENCE's own class, so the check is structural, not by type."""
    return isinstance(getattr(m, "conv3d", None), nn.Conv3d) and isinstance(getattr(m, "alpha", None), torch.Tensor)


def _resize_rel_pos(rel_pos: torch.Tensor, size: int) -> torch.Tensor:
    """get_rel_pos's table resize (image_encoder.py:399-408): linear interpolation when len != 2*size-1 (host-side, once)."""
    L = 2 * size - 1
    if rel_pos.shape[0] == L:
        return rel_pos
    r = torch.nn.functional.interpolate(rel_pos.float().reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
    return r.reshape(-1, L).permute(1, 0)


class ImageEncoderViT(nn.Module):
    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, out_chans: int = 256, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True,
                 use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 global_attn_indexes: Tuple[int, ...] = (), adapter_type: str = "conv") -> None:
        super().__init__()
        if patch_size != 16 or in_chans != 3:
            raise NotImplementedError("grove_b200 builds the 16x16 RGB patch embed only (build_sam.py:65-84)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("grove_b200 fuses the exact-erf GELU of the reference MLP (common.py:18)")
        self.img_size = img_size
        self.embed_dim = embed_dim
        self.out_chans = out_chans
        self.num_heads = num_heads
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size), in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                                     act_layer=act_layer, use_rel_pos=use_rel_pos, rel_pos_zero_init=rel_pos_zero_init,
                                     window_size=window_size if i not in global_attn_indexes else 0,
                                     input_size=(img_size // patch_size, img_size // patch_size)))
        if adapter_type == "conv":
            self.adapters = nn.ModuleList([SpatioTemporalConvAdapter(embed_dim, embed_dim, (3, 3, 3)) for _ in range(len(global_attn_indexes))])
        elif adapter_type == "transformer":
            raise NotImplementedError("the reference's TemporalAdapter is broken (image_encoder.py:32 uses an undefined self.relu) and never selected")
        else:
            self.adapters = nn.ModuleList([nn.Identity() for _ in range(len(global_attn_indexes))])
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False), LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False), LayerNorm2d(out_chans))
        self.global_attn_indexes = global_attn_indexes
        self._pack = PackCache()
        self._use_graphs = False
        self._graph = None          # (signature, CUDAGraph, static patches, static embeddings, kernels per replay)
        # dtype of the residual stream between blocks.  bfloat16 (default; the reference runs the whole model in bf16, train.py:618) halves
        # the residual GEMMs' epilogue traffic and lets the LayerNorms fold into the consuming GEMMs; every sum is formed in fp32 and rounded
        # once per block half.  Measured end-to-end drift vs the fp32 oracle is the same as with torch.float32 (boxes ~1.5e-4, logits ~1e-3,
        # budget 1e-2 / 2e-2) -- it is set by the bf16 tensor-core operands, not by the stream.  The training step
        # (encoder_train.py) keeps its own fp32 stream.
        self.residual_dtype = torch.bfloat16

    # ------------------------------------------------------------------ weight packing (cached)
    def _linear(self, key, lin: nn.Linear):
        w = self._pack.get(key + ".w", [lin.weight], bf16)
        b = self._pack.get(key + ".b", [lin.bias], f32) if lin.bias is not None else None
        return w, b

    def _ln(self, key, ln):
        return self._pack.get(key + ".g", [ln.weight], f32), self._pack.get(key + ".b", [ln.bias], f32)

    def _ln_folded(self, key, ln: nn.LayerNorm, lin: nn.Linear):
        """(W * gamma as bf16 [N, K], b + W.beta fp32 [N], column sums of the ROUNDED W * gamma fp32 [N]) for a LayerNorm -> Linear pair
        whose normalisation runs inside the GEMM epilogue (grove_gemm_epilogue.ln_stats)."""
        def build(w, b, g, be):
            wg = (w.float() * g.float()[None, :]).to(torch.bfloat16).contiguous()
            bias = (b.float() if b is not None else torch.zeros(w.shape[0], device=w.device)) + w.float() @ be.float()
            return wg, bias.contiguous(), wg.float().sum(1).contiguous()
        if lin.bias is None:
            return self._pack.get(key + ".fold", [lin.weight, ln.weight, ln.bias], lambda w, g, be: build(w, None, g, be))
        return self._pack.get(key + ".fold", [lin.weight, lin.bias, ln.weight, ln.bias], build)

    # ------------------------------------------------------------------ CUDA-graph replay of the block stack (serving)
    def enable_cuda_graphs(self, on: bool = True) -> None:
        """Serving mode: capture the ~100 launches from the patch-embed GEMM to the neck into one CUDA graph per input shape and
        replay it (no per-kernel launch gaps, no host work per kernel; measured +0.2 .. +2.5 % on the same box).  The graph bakes in the
        packed weights' addresses; it is re-captured when any parameter is reassigned, moved or modified in place (same signature rule
        as PackCache).  If capture fails the encoder falls back to kernel-by-kernel launches."""
        self._use_graphs = bool(on)
        self._graph = None

    def _encode_patches_cached(self, patches: torch.Tensor, Fr: int, G: int) -> torch.Tensor:
        if not self._use_graphs:
            return self._encode_patches(patches, Fr, G)
        sig = (Fr, G, patches.device, tuple(patches.shape), self.residual_dtype, tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if self._graph is None or self._graph[0] != sig:
            self._graph = None
            self._encode_patches(patches, Fr, G)            # warm-up: packs the weights, sets the kernels' shared-memory attributes
            torch.cuda.synchronize(patches.device)
            static_in = torch.empty_like(patches)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            try:
                with torch.cuda.graph(graph):
                    static_out = self._encode_patches(static_in, Fr, G)
            except Exception as e:                           # capture is an optimisation: fall back to kernel-by-kernel launches
                import warnings
                warnings.warn(f"grove_b200: CUDA-graph capture of the encoder failed ({e}); launching kernel by kernel")
                self._use_graphs = False
                torch.cuda.synchronize(patches.device)
                return self._encode_patches(patches, Fr, G)
            self._graph = (sig, graph, static_in, static_out, ops.launch_count() - n0)
        _, graph, static_in, static_out, n_kernels = self._graph
        static_in.copy_(patches)
        graph.replay()
        ops.add_launch_count(n_kernels)
        return static_out.clone()

    # ------------------------------------------------------------------ forward
    @no_grad_entry("ImageEncoderViT.forward", lambda self: self.adapters.parameters())
    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """[V,3,T,H,W] -> token-major embeddings [V*T, G*G, out_chans] bf16"""
        if x.dim() != 5 or x.shape[1] != 3:
            raise ValueError(f"expected images of shape [V,3,T,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("grove_b200.ImageEncoderViT runs on CUDA only (no CPU fallback)")
        V, _, T, H, W = x.shape
        if H != W or H % 16:
            raise ValueError("square inputs with side a multiple of 16 are required")
        with ops.device_of(x):
            img = x.to(torch.bfloat16).contiguous()
            patches = torch.empty(V * T * (H // 16) ** 2, 768, device=x.device, dtype=torch.bfloat16)
            ops.im2col_patch16(img, patches)
            return self._encode_patches_cached(patches, V * T, H // 16)

    @torch.no_grad()
    def forward_frames(self, frames: torch.Tensor, transform=None) -> torch.Tensor:
        """Decoded RGB frames, uint8 [V,T,h,w,3] on the GPU -> [V*T, out_chans, G, G]: the reference's host-side
        ResizeLongestSide.apply_image + grounding_enc_processor + .bfloat16() (transforms.py:27-34, HowTo100M.py:168-178, train.py:751-753)
        are fused into the patch embed's operand (grove_b200.preprocess); bit-identical to forward() on the host-processed tensor."""
        from ..preprocess import ResizeLongestSide
        if frames.dim() != 5 or frames.shape[-1] != 3 or frames.dtype != torch.uint8:
            raise ValueError(f"expected uint8 frames of shape [V,T,h,w,3], got {tuple(frames.shape)} {frames.dtype}")
        if not frames.is_cuda:
            raise RuntimeError("grove_b200.ImageEncoderViT runs on CUDA only (no CPU fallback)")
        V, T, h, w, _ = frames.shape
        tr = transform if transform is not None else ResizeLongestSide(self.img_size)
        G = self.img_size // 16
        with ops.device_of(frames):
            patches = tr.patches(frames.reshape(V * T, h, w, 3).contiguous(), self.img_size)
            tok = self._encode_patches_cached(patches, V * T, G)
        out = tok.view(V * T, G, G, self.out_chans).permute(0, 3, 1, 2)
        want = self.pos_embed.dtype if self.pos_embed is not None else torch.bfloat16
        return out if want == torch.bfloat16 else out.to(want)

    def _encode_patches(self, patches: torch.Tensor, Fr: int, G: int) -> torch.Tensor:
        """patches bf16 [Fr*G*G, 768] (k = c*256 + py*16 + px) -> token-major embeddings [Fr, G*G, out_chans] bf16"""
        D, heads = self.embed_dim, self.num_heads
        hd, N = D // heads, G * G
        has_conv = any(is_conv_adapter(a) for a in self.adapters)
        if has_conv and Fr % 8:
            raise ValueError("the spatio-temporal adapter groups frames by 8 (image_encoder.py:52): V*T must be a multiple of 8")
        dev = patches.device
        M = Fr * N
        lowp = self.residual_dtype == torch.bfloat16
        if self.residual_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("residual_dtype must be torch.float32 or torch.bfloat16")
        xs = torch.empty(M, D, device=dev, dtype=self.residual_dtype)   # residual stream
        wpe = self._pack.get("pe.w", [self.patch_embed.proj.weight], lambda w: bf16(w.reshape(w.shape[0], -1)))
        bpe = self._pack.get("pe.b", [self.patch_embed.proj.bias], f32)
        pos = None
        if self.pos_embed is not None:
            if self.pos_embed.shape[1] != G or self.pos_embed.shape[2] != G:
                raise ValueError(f"pos_embed is {tuple(self.pos_embed.shape)} but the input grid is {G}x{G} (train.py:561-565 resizes it)")
            pos = self._pack.get("pos", [self.pos_embed], lambda p: f32(p.reshape(N, D)))
        ops.gemm(patches, wpe, xs, bias=bpe, resid=pos, resid_row_mod=N if pos is not None else 0)
        del patches

        h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
        att = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        mlp_dim = self.blocks[0].mlp.lin1.out_features
        hid = torch.empty(M, mlp_dim, device=dev, dtype=torch.bfloat16)
        # fp32 stream: bf16 copies of it feed the conv / neck tensor-core operands; bf16 stream: the stream is its own operand and the
        # adapter (which reads its neighbours' rows) writes into a second buffer
        xb = None if lowp else torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        xb2 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        last = len(self.blocks) - 1
        # bf16 stream: the residual GEMMs' epilogues emit per-row (sum, sum of squares) of what they write, and the LayerNorm that follows is
        # folded into the GEMM that consumes it (A = the raw stream, W carries gamma) -- no LayerNorm pass, no normalised copy in HBM
        stats = torch.empty(M, D // 128, 2, device=dev, dtype=torch.float32) if lowp else None
        have_stats = False
        for i, blk in enumerate(self.blocks):
            k = f"b{i}"
            if have_stats:
                wq, bq, cq = self._ln_folded(k + ".qkv", blk.norm1, blk.attn.qkv)
                ops.gemm(xs, wq, qkv, bias=bq, ln_fold=(stats, cq, blk.norm1.eps))
            else:
                g1, b1 = self._ln(k + ".n1", blk.norm1)
                ops.layernorm(xs, g1, b1, h, blk.norm1.eps)
                wq, bq = self._linear(k + ".qkv", blk.attn.qkv)
                ops.gemm(h, wq, qkv, bias=bq)
            S = blk.window_size if blk.window_size > 0 else G
            if blk.window_size > 0:
                bqb = self._pack.get(k + ".qkvb16", [blk.attn.qkv.bias], bf16)
                tab = self._pack.get(k + ".reltab", [blk.attn.rel_pos_h, blk.attn.rel_pos_w],
                                     lambda a, b, S=S: ops.window_rel_table(_resize_rel_pos(a, S), _resize_rel_pos(b, S)))
                ops.attn_window_tc(qkv, bqb, tab, att, F=Fr, G=G, heads=heads, hd=hd, ws=blk.window_size)
            else:
                rh = self._pack.get(k + ".rh", [blk.attn.rel_pos_h], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
                rw = self._pack.get(k + ".rw", [blk.attn.rel_pos_w], lambda t, S=S: bf16(_resize_rel_pos(t, S)))
                ops.attn_global(qkv, rh, rw, att, F=Fr, G=G, heads=heads, hd=hd)
            wp, bp = self._linear(k + ".proj", blk.attn.proj)
            if lowp:
                ops.gemm(att, wp, xs, bias=bp, resid=xs, ln_stats_out=stats)
                w1, bb1, c1 = self._ln_folded(k + ".l1", blk.norm2, blk.mlp.lin1)
                ops.gemm(xs, w1, hid, bias=bb1, act="gelu", ln_fold=(stats, c1, blk.norm2.eps))
            else:
                ops.gemm(att, wp, xs, bias=bp, resid=xs)
                g2, b2 = self._ln(k + ".n2", blk.norm2)
                ops.layernorm(xs, g2, b2, h, blk.norm2.eps)
                w1, bb1 = self._linear(k + ".l1", blk.mlp.lin1)
                ops.gemm(h, w1, hid, bias=bb1, act="gelu")
            w2, bb2 = self._linear(k + ".l2", blk.mlp.lin2)
            adapter = self.adapters[self.global_attn_indexes.index(i)] if i in self.global_attn_indexes else None
            conv = is_conv_adapter(adapter)
            if adapter is not None and not conv and not isinstance(adapter, nn.Identity):
                raise NotImplementedError(f"unsupported adapter module {type(adapter).__name__}")
            have_stats = lowp and not conv and i != last       # the adapter rewrites the stream: the next norm1 runs as its own pass
            ops.gemm(hid, w2, xs, bias=bb2, resid=xs, out2=xb if ((conv or i == last) and not lowp) else None,
                     ln_stats_out=stats if have_stats else None)
            if conv and lowp:
                c3 = adapter.conv3d
                if tuple(c3.kernel_size) != (3, 3, 3) or c3.in_channels != D or c3.out_channels != D:
                    raise NotImplementedError("adapter Conv3d must be DxDx3x3x3 (image_encoder.py:139-143)")
                wc = self._pack.get(k + ".c3w", [c3.weight], lambda w: bf16(w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)))
                bc = self._pack.get(k + ".c3b", [c3.bias], f32)
                al = self._pack.get(k + ".alpha", [adapter.alpha], f32)
                ops.conv_gemm(xs, wc, xb2, V=Fr // 8, T=8, G=G, kt=3, bias=bc, act="relu", gate_alpha=al, resid=xs)
                xs, xb2 = xb2, xs
            elif conv:
                c3 = adapter.conv3d
                if tuple(c3.kernel_size) != (3, 3, 3) or c3.in_channels != D or c3.out_channels != D:
                    raise NotImplementedError("adapter Conv3d must be DxDx3x3x3 (image_encoder.py:139-143)")
                wc = self._pack.get(k + ".c3w", [c3.weight], lambda w: bf16(w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)))
                bc = self._pack.get(k + ".c3b", [c3.bias], f32)
                al = self._pack.get(k + ".alpha", [adapter.alpha], f32)
                ops.conv_gemm(xb, wc, xs, V=Fr // 8, T=8, G=G, kt=3, bias=bc, act="relu", gate_alpha=al, resid=xs,
                              out2=xb2 if i == last else None)
                if i == last:
                    xb, xb2 = xb2, xb
        # neck
        C = self.out_chans
        wn0 = self._pack.get("n0", [self.neck[0].weight], lambda w: bf16(w.reshape(w.shape[0], -1)))
        y0 = torch.empty(M, C, device=dev, dtype=torch.float32)
        ops.gemm(xs if lowp else xb, wn0, y0)
        g, b = self._ln("n1", self.neck[1])
        y1 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        ops.layernorm(y0, g, b, y1, self.neck[1].eps)
        wn2 = self._pack.get("n2", [self.neck[2].weight], lambda w: bf16(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)))
        ops.conv_gemm(y1, wn2, y0, V=Fr, T=1, G=G, kt=1)
        g, b = self._ln("n3", self.neck[3])
        emb = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        ops.layernorm(y0, g, b, emb, self.neck[3].eps)
        return emb.view(Fr, N, C)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference contract: [V,3,T,H,W] -> [V*T, out_chans, G, G].  The result is a channels-last view of the
        token-major buffer (no NCHW copy): MaskDecoder consumes it zero-copy; `.contiguous()` gives plain NCHW."""
        tok = self.forward_tokens(x)
        Fr, N, C = tok.shape
        G = int(round(N ** 0.5))
        out = tok.view(Fr, G, G, C).permute(0, 3, 1, 2)
        want = self.pos_embed.dtype if self.pos_embed is not None else torch.bfloat16
        return out if want == torch.bfloat16 else out.to(want)
