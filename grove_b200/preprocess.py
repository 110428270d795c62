"""Frame pre-processing on the GPU, fused into the patch embed's operand.

Mirrors the reference's `ResizeLongestSide` (model/SAM/utils/transforms.py:17-113) and `grounding_enc_processor`
(dataset/video_grounding_datasets/HowTo100M.py:168-178, infer_iground.py:304-318).  The reference resizes every frame on the
host with PIL, normalises in fp32, pads and ships 6 MB per 1024^2 frame to the GPU, where `.bfloat16()` (train.py:751-753)
makes another copy.  Here the decoded uint8 frames go to the GPU as they are (3 bytes per source pixel) and two
kernels produce the PatchEmbed GEMM's A operand directly: a horizontal resampling pass (uint8 -> uint8) and a fused vertical
pass + (x - mean) / std + zero pad + bf16 + 16x16 patchify.  The arithmetic is Pillow's (22-bit fixed-point bilinear, restated from
its Resample.c) and the reference's (fp32 subtract, divide, round to bf16): results are bit-identical to the host pipeline.
The resampling coefficient tables are built on the host in float64 exactly like Pillow's precompute_coeffs and cached per size.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Tuple

import numpy as np
import torch

from . import ops
from ._lib import check, lib

IMG_MEAN = (123.675, 116.28, 103.53)      # HowTo100M.py:86-87, infer_iground.py:305-306
IMG_STD = (58.395, 57.12, 57.375)
_PRECISION_BITS = 32 - 8 - 2


@functools.lru_cache(maxsize=64)
def _bilinear_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the triangle filter, vectorised: (bounds [out,2], coeffs [out,ksize])"""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    centers = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((centers - support + 0.5).astype(np.int64), 0)          # C casts truncate; the operands are non-negative here
    xmin = np.where(centers - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((centers + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((x + xmin[:, None] - centers[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(x < cnt[:, None], w, 0.0)
    # Pillow sums the taps one by one in index order; np.cumsum reproduces that order of additions
    ww = np.cumsum(w, axis=1)[:, -1:]
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kk = (0.5 + w * float(1 << _PRECISION_BITS)).astype(np.int64).astype(np.int32)     # weights of the triangle filter are >= 0
    bounds = np.stack([xmin, cnt], 1).astype(np.int32)
    return bounds, kk


def _identity_tables(n: int) -> Tuple[np.ndarray, np.ndarray]:
    return np.stack([np.arange(n), np.ones(n, dtype=np.int64)], 1).astype(np.int32), np.full((n, 1), 1 << _PRECISION_BITS, dtype=np.int32)


class ResizeLongestSide:
    """transforms.py:17-113.  `apply_image` takes / returns uint8 frames on the GPU ([h,w,3] or [F,h,w,3])."""

    def __init__(self, target_length: int) -> None:
        self.target_length = target_length
        self._dev_tables = {}

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        scale = long_side_length * 1.0 / max(oldh, oldw)
        newh, neww = oldh * scale, oldw * scale
        return int(newh + 0.5), int(neww + 0.5)

    def _tables(self, in_size, out_size, device):
        key = (in_size, out_size, str(device))
        t = self._dev_tables.get(key)
        if t is None:
            b, k = _identity_tables(in_size) if in_size == out_size else _bilinear_tables(in_size, out_size)
            t = (torch.from_numpy(b).to(device), torch.from_numpy(np.ascontiguousarray(k)).to(device), k.shape[1])
            self._dev_tables[key] = t
        return t

    def _resize_rows(self, frames: torch.Tensor, new_w: int) -> torch.Tensor:
        Fr, h, w, _ = frames.shape
        if new_w == w:
            return frames
        b, k, ks = self._tables(w, new_w, frames.device)
        out = torch.empty(Fr, h, new_w, 3, device=frames.device, dtype=torch.uint8)
        check(lib().grove_resize_rows_u8(C.c_void_p(frames.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(k.data_ptr()),
                                         ks, Fr * h, w, new_w, C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)),
              "grove_resize_rows_u8")
        return out

    def patches(self, frames: torch.Tensor, img_size: int, mean=IMG_MEAN, std=IMG_STD) -> torch.Tensor:
        """uint8 frames [F,h,w,3] (decoded RGB, CUDA) -> bf16 [F*(img/16)^2, 768]: apply_image + grounding_enc_processor + .bfloat16()
        + the patch embed's im2col, without the intermediate images"""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda or not frames.is_contiguous():
            raise RuntimeError("grove_b200: frames must be a contiguous CUDA uint8 tensor [F,h,w,3] (there is no CPU path)")
        Fr, h, w, _ = frames.shape
        nh, nw = self.get_preprocess_shape(h, w, self.target_length)
        if nh > img_size or nw > img_size:
            raise ValueError(f"resized frames ({nh}x{nw}) do not fit the {img_size}x{img_size} encoder input")
        x = self._resize_rows(frames, nw)
        b, k, ks = self._tables(h, nh, frames.device)
        G = img_size // 16
        out = torch.empty(Fr * G * G, 768, device=frames.device, dtype=torch.bfloat16)
        m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        check(lib().grove_frames_to_patches_u8(C.c_void_p(x.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(k.data_ptr()), ks, C.c_void_p(out.data_ptr()),
                                               Fr, h, nw, nh, img_size, m3, s3, C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)),
              "grove_frames_to_patches_u8")
        return out

    def apply_image(self, image: torch.Tensor) -> torch.Tensor:
        """uint8 [h,w,3] or [F,h,w,3] on the GPU -> resized uint8 (transforms.py:27-34)"""
        single = image.dim() == 3
        fr = image.unsqueeze(0) if single else image
        if fr.dtype != torch.uint8 or not fr.is_cuda:
            raise RuntimeError("grove_b200: apply_image expects CUDA uint8 frames (there is no CPU path)")
        fr = fr.contiguous()
        Fr, h, w, _ = fr.shape
        nh, nw = self.get_preprocess_shape(h, w, self.target_length)
        x = self._resize_rows(fr, nw)
        if nh != h:   # the vertical pass is the horizontal pass of the transposed problem: rows of length h, w*... -> run it per column block
            xt = x.permute(0, 2, 1, 3).contiguous()                     # [F, w', h, 3]
            b, k, ks = self._tables(h, nh, fr.device)
            o = torch.empty(Fr, nw, nh, 3, device=fr.device, dtype=torch.uint8)
            check(lib().grove_resize_rows_u8(C.c_void_p(xt.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(k.data_ptr()),
                                             ks, Fr * nw, h, nh, C.c_void_p(torch.cuda.current_stream(fr.device).cuda_stream)), "grove_resize_rows_u8")
            x = o.permute(0, 2, 1, 3).contiguous()
        return x[0] if single else x

    # coordinate helpers (transforms.py:36-98): plain arithmetic, device-agnostic
    def apply_coords(self, coords, original_size):
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(old_h, old_w, self.target_length)
        coords = coords.clone().to(torch.float) if isinstance(coords, torch.Tensor) else np.array(coords, dtype=float, copy=True)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes(self, boxes, original_size):
        return self.apply_coords(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)

    apply_coords_torch, apply_boxes_torch = apply_coords, apply_boxes
