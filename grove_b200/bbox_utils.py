"""Box format helpers with the reference's names and signatures (utils/bbox_utils.py:4-80).  Host-side utilities for
numpy arrays / tensors on any device; the device-side fused equivalent on the hot path is grove_box_postprocess."""
import numpy as np
import torch


def _like(x, type):
    return np.zeros_like(x) if type == "np" else torch.zeros_like(x)


def normalize_bboxes(bboxes, image_width, image_height, type="np"):
    out = _like(bboxes, type)
    out[:, 0] = bboxes[:, 0] / image_width
    out[:, 1] = bboxes[:, 1] / image_height
    out[:, 2] = bboxes[:, 2] / image_width
    out[:, 3] = bboxes[:, 3] / image_height
    return out


def unnormalize_bboxes(normalized_bboxes, image_width, image_height, type="np"):
    out = _like(normalized_bboxes, type)
    out[:, 0] = normalized_bboxes[:, 0] * image_width
    out[:, 1] = normalized_bboxes[:, 1] * image_height
    out[:, 2] = normalized_bboxes[:, 2] * image_width
    out[:, 3] = normalized_bboxes[:, 3] * image_height
    return out


def box_cxcywh_to_xyxy(boxes, type="np"):
    cx, cy, w, h = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    parts = (cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2)
    return np.stack(parts, axis=-1) if type == "np" else torch.stack(parts, dim=-1)


def box_xyxy_to_cxcywh(boxes, type="np"):
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    parts = ((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1)
    return np.stack(parts, axis=-1) if type == "np" else torch.stack(parts, dim=-1)
