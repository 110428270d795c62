"""Box-IoU / decision utilities of the reference's eval scripts on the GPU, bit-exact with the
numpy / torch originals: same names, argument meaning and return types, inputs/outputs as numpy arrays."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("grove_b200.box_eval needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def np_box_iou(boxes1: np.ndarray, boxes2: np.ndarray) -> np.ndarray:
    """eval_vidstg.py:47-63."""
    dt = np.result_type(boxes1.dtype, boxes2.dtype)
    dt = np.float32 if dt == np.float32 else np.float64
    a = torch.from_numpy(np.ascontiguousarray(boxes1, dtype=dt)).to(_dev())
    b = torch.from_numpy(np.ascontiguousarray(boxes2, dtype=dt)).to(_dev())
    return ops.box_iou(a, b, 0).cpu().numpy()


def compute_iou_matrix(pred_bboxes, gt_bboxes) -> np.ndarray:
    """eval_iground.py:58-63 (rows = first argument), float64, +1 convention of compute_iou :39-56."""
    a = torch.tensor(np.asarray(pred_bboxes, dtype=np.float64).reshape(-1, 4), device=_dev())
    b = torch.tensor(np.asarray(gt_bboxes, dtype=np.float64).reshape(-1, 4), device=_dev())
    return ops.box_iou(a, b, 1).cpu().numpy()


def compute_iou(pred_bbox, gt_bbox) -> float:
    return float(compute_iou_matrix([pred_bbox], [gt_bbox])[0, 0])


def greedy_matches(ious: np.ndarray, text_sims: np.ndarray, iou_threshold: float, text_sim_threshold: float):
    """the matching loop of find_best_matches, eval_iground.py:85-96 (the BERT similarity matrix is the caller's)."""
    if ious.size == 0 or text_sims.size == 0:
        return []
    i = torch.from_numpy(np.ascontiguousarray(ious, dtype=np.float64)).to(_dev())
    s = torch.from_numpy(np.ascontiguousarray(text_sims, dtype=np.float64)).to(_dev())
    return ops.greedy_match(i, s, iou_threshold, text_sim_threshold)


def bbox_overlaps_batch(anchors, gt_boxes, frm_mask=None):
    """eval_anet.py:22-119, 3-D branch: anchors [b,N,5], gt_boxes [b,K,5], frm_mask [b,N,K] (1 = different frame)."""
    anchors = torch.as_tensor(anchors, dtype=torch.float32)
    gt_boxes = torch.as_tensor(gt_boxes, dtype=torch.float32)
    if anchors.dim() != 3:
        raise NotImplementedError("only the 3-D (per-frame proposals) branch is used by the GROVE pipeline (eval_anet.py:204)")
    out = []
    for bi in range(anchors.shape[0]):
        a = anchors[bi].contiguous().to(_dev())
        g = gt_boxes[bi].contiguous().to(_dev())
        m = None if frm_mask is None else torch.as_tensor(frm_mask[bi]).to(torch.uint8).contiguous().to(_dev())
        out.append(ops.box_iou(a, g, 2, frm_mask=m).cpu())
    return torch.stack(out)
