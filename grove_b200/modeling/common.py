"""Parameter containers shared by encoder and decoder (reference: model/SAM/modeling/common.py).

The modules below keep the reference's constructor signatures and parameter names (so reference
checkpoints load with the same keys) but own no math: the enclosing ImageEncoderViT / MaskDecoder lower
the whole forward onto the CUDA library.  Calling one of them on its own is a caller error.
"""
from __future__ import annotations

import functools
from typing import Type

import torch
import torch.nn as nn


class _ContainerOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is a parameter container in grove_b200: its math is fused into the "
                           "parent module's CUDA pipeline (there is no per-layer PyTorch path)")


class MLPBlock(_ContainerOnly):
    """common.py:13-26 — lin1 / GELU / lin2."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()


class LayerNorm2d(_ContainerOnly):
    """common.py:31-43 — channel-first LayerNorm, eps 1e-6."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps


class PackCache:
    """Kernel-friendly copies of parameters (bf16 / fp32, repacked), rebuilt when a parameter is reassigned,
    modified in place (optimizer step, load_state_dict) or moved."""

    def __init__(self):
        self._store = {}

    def get(self, key, tensors, fn):
        sig = tuple((t.data_ptr(), t._version, t.device, t.dtype) for t in tensors)
        hit = self._store.get(key)
        if hit is None or hit[0] != sig:
            with torch.no_grad():
                hit = (sig, fn(*tensors))
            self._store[key] = hit
        return hit[1]


def f32(t):
    return t.detach().to(torch.float32).contiguous()


def bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


_WARNED_NO_AUTOGRAD = set()


def warn_no_autograd(what: str, params, *inputs) -> None:
    """The forward entry points run the CUDA library under no_grad and return tensors without a grad_fn.  The reference trains the adapters,
    the mask decoder and text_hidden_fcs THROUGH these calls (GROVE.py:134-136, train.py:279-296), so a trainer that only swaps the import
    would silently lose those gradients: say so (once per entry point) whenever autograd is recording and something here wants a gradient."""
    import warnings
    if what in _WARNED_NO_AUTOGRAD or not torch.is_grad_enabled():
        return
    if any(getattr(t, "requires_grad", False) for t in inputs) or any(p.requires_grad for p in params):
        _WARNED_NO_AUTOGRAD.add(what)
        warnings.warn(f"grove_b200: {what} does not record an autograd graph — its outputs carry no gradient to the trainable grounding "
                      "parameters or inputs.  For training call GroundingBranch.grounding_loss(...) (forward + the library's own backward pass, "
                      "delivers .grad to adapters / mask decoder / text_hidden_fcs and to last_hidden_state); for inference wrap the call in "
                      "torch.no_grad().", stacklevel=3)


def no_grad_entry(what: str, params_of):
    """@torch.no_grad() for a forward entry point, with the autograd check done BEFORE gradients are switched off"""
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(self, *a, **k):
            if torch.is_grad_enabled():
                tensors = [x for x in list(a) + list(k.values()) if isinstance(x, torch.Tensor)]
                tensors += [y for x in list(a) + list(k.values()) if isinstance(x, (list, tuple)) for y in x if isinstance(y, torch.Tensor)]
                warn_no_autograd(what, list(params_of(self)), *tensors)
            with torch.no_grad():
                return fn(self, *a, **k)
        return wrapper
    return deco
