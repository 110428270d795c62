"""TEST INFRASTRUCTURE ONLY — numpy restatement of the frame pre-processing that precedes the grounding path:

  ResizeLongestSide.apply_image            model/SAM/utils/transforms.py:27-34,100-113  (torchvision resize of a PIL image, bilinear)
  grounding_enc_processor                  dataset/video_grounding_datasets/HowTo100M.py:168-178, infer_iground.py:304-318
  images.bfloat16()                        train.py:751-753

The resize arithmetic lives in a third-party dependency that is not under /root/reference: Pillow (unpinned by the reference;
README.md:53 pins torch/torchvision only), reached through torchvision.transforms.functional.resize -> PIL.Image.resize(BILINEAR).
Its published algorithm (src/libImaging/Resample.c) is restated here: a triangle filter whose support is scaled by the
down-sampling factor, coefficients normalised in double and rounded to 22-bit fixed point, a horizontal pass followed by a
vertical pass, each accumulating in int32 from 1 << 21 and clipping to uint8.  Pinned bit-for-bit by tests/test_oracle_golden.py
against frames resized by Pillow 12.2 through the reference's own call chain (tests/golden/preprocess.npz).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMG_MEAN = np.array([123.675, 116.28, 103.53], dtype=np.float32)   # HowTo100M.py:86-87 / infer_iground.py:305-306
IMG_STD = np.array([58.395, 57.12, 57.375], dtype=np.float32)


def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
    """transforms.py:100-113"""
    scale = long_side_length * 1.0 / max(oldh, oldw)
    newh, neww = oldh * scale, oldw * scale
    return int(newh + 0.5), int(neww + 0.5)


def bilinear_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the full box [0, in_size).
    Returns (bounds int32 [out,2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    """one pass of ImagingResample{Horizontal,Vertical}_8bpc over `axis` of an [H,W,C] uint8 image"""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """PIL.Image.resize((new_w, new_h), BILINEAR) of an [H,W,3] uint8 image: horizontal pass, then vertical pass (each skipped when
    that side keeps its length)."""
    h, w = img.shape[:2]
    out = img
    if new_w != w:
        out = _resample_axis(out, *bilinear_coeffs(w, new_w), axis=1)
    if new_h != h:
        out = _resample_axis(out, *bilinear_coeffs(h, new_h), axis=0)
    return out


def apply_image(img: np.ndarray, target_length: int) -> np.ndarray:
    """ResizeLongestSide.apply_image (transforms.py:27-34)"""
    nh, nw = get_preprocess_shape(img.shape[0], img.shape[1], target_length)
    return resize_bilinear_u8(img, nh, nw)


def grounding_enc_processor(frames_u8: np.ndarray, img_size: int) -> np.ndarray:
    """HowTo100M.py:168-178 on resized frames [T,h,w,3] uint8 -> float32 [3,T,img,img]: (x - mean) / std, then zero pad right / bottom"""
    x = frames_u8.transpose(3, 0, 1, 2).astype(np.float32)
    x = (x - IMG_MEAN[:, None, None, None]) / IMG_STD[:, None, None, None]
    out = np.zeros((3, x.shape[1], img_size, img_size), dtype=np.float32)
    out[:, :, :x.shape[2], :x.shape[3]] = x
    return out
