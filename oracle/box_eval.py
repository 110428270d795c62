"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's box-IoU / decision utilities
(SURVEY.md §2 row 9).  Pinned by tests/test_oracle_golden.py against the reference's functions
executed in the build container (oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np


def np_box_iou(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """eval_vidstg.py:13-63 (np_box_area, _box_inter_union, np_box_iou): xyxy, no +1, dtype follows
    numpy promotion of the inputs; division by zero union yields nan/inf exactly as numpy does."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[:, :2])
    rb = np.minimum(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clip(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (a1[:, None] + a2 - inter)


def viou_recalls(ious, n_gt_frames: int, thresholds):
    """eval_vidstg.py:157-178: IoUs accumulated in frame order (`gt_viou += iou`, Python floats), divided by max(n, 1),
    strict '>' against each threshold."""
    acc = 0
    for v in ious:
        acc += float(v)
    v = acc / max(n_gt_frames, 1)
    return v, [1 if v > t else 0 for t in thresholds]


def video_viou(pred_boxes, gt_boxes, thresholds):
    """eval_vidstg.py:157-178 for one video: pred_boxes / gt_boxes [n,4] xyxy aligned per ground-truth frame; a prediction that is all
    zeros scores 0 without an IoU (`pred_box.any()`)."""
    ious = []
    for p, g in zip(np.asarray(pred_boxes, dtype=np.float64), np.asarray(gt_boxes, dtype=np.float64)):
        ious.append(float(np_box_iou(p[None], g[None])[0][0]) if p.any() else 0)
    v, over = viou_recalls(ious, len(ious), thresholds)
    return v, over, ious


def localization_accuracy(pred_boxes, gt_boxes, kinds=None):
    """eval_youcookinteractions.py:8-51 on flat arrays: kinds 0 normal, 1 empty ground truth (skipped), 2 missing prediction, 3 NaN
    prediction (both valid-but-wrong).  Returns (accuracy %, correct, valid, per-pair flags)."""
    n = len(pred_boxes)
    kinds = np.zeros(n, dtype=np.int64) if kinds is None else kinds
    correct = valid = 0
    flags = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        if kinds[i] == 1:
            continue
        valid += 1
        if kinds[i] == 2 or np.any(np.isnan(pred_boxes[i])):
            continue
        if center_in_box([float(v) for v in pred_boxes[i]], [float(v) for v in gt_boxes[i]]):
            correct += 1
            flags[i] = 1
    return ((correct / valid) * 100 if valid else 0.0), correct, valid, flags


def giou_loss_xyxy(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """torchvision.ops.generalized_box_iou_loss per row, fp32, on the coordinates as given (eps 1e-7)."""
    b1, b2 = b1.astype(np.float32), b2.astype(np.float32)
    x1, y1, x2, y2 = b1.T
    x1g, y1g, x2g, y2g = b2.T
    xk1, yk1, xk2, yk2 = np.maximum(x1, x1g), np.maximum(y1, y1g), np.minimum(x2, x2g), np.minimum(y2, y2g)
    inter = np.where((yk2 > yk1) & (xk2 > xk1), (xk2 - xk1) * (yk2 - yk1), np.float32(0))
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + np.float32(1e-7))
    ac = (np.maximum(x2, x2g) - np.minimum(x1, x1g)) * (np.maximum(y2, y2g) - np.minimum(y1, y1g))
    return np.float32(1) - (iou - (ac - union) / (ac + np.float32(1e-7)))


def val_giou_and_objectness_accuracy(pred_bboxes, logits, gt_bboxes, gt_obj):
    """train.py:821-840: nested [V][T] lists; GIoU loss of pred[gt_obj] vs gt ON THE COORDINATES AS GIVEN (the reference passes cxcywh),
    accumulated per frame in fp32; objectness hits = ((sigmoid(logit) > 0.5) == label).  Returns (giou_sum, hits, num_bboxes, num_max)."""
    giou, hits, nb, nmax = np.float32(0), 0, 0, 0
    for v in range(len(pred_bboxes)):
        for f in range(len(pred_bboxes[v])):
            pb, lg = np.asarray(pred_bboxes[v][f], dtype=np.float32), np.asarray(logits[v][f], dtype=np.float32)
            go, gb = np.asarray(gt_obj[v][f]).astype(np.int64), np.asarray(gt_bboxes[v][f], dtype=np.float32).reshape(-1, 4)
            if gb.shape[0]:
                giou = np.float32(giou + giou_loss_xyxy(pb[go.astype(bool)], gb).sum(dtype=np.float32))
            sg = np.float32(1) / (np.float32(1) + np.exp(-lg, dtype=np.float32))
            hits += int(((sg > 0.5).astype(np.int64) == go).sum())
            nb += gb.shape[0]
            nmax += pb.shape[0]
    return float(giou), hits, nb, nmax


def compute_iou_plus1(p, g) -> float:
    """eval_iground.py:39-56: +1 pixel convention on 4-vectors treated as xyxy; 0.0 when the union is 0."""
    xA, yA = max(p[0], g[0]), max(p[1], g[1])
    xB, yB = min(p[2], g[2]), min(p[3], g[3])
    inter = max(0, xB - xA + 1) * max(0, yB - yA + 1)
    aA = (p[2] - p[0] + 1) * (p[3] - p[1] + 1)
    aB = (g[2] - g[0] + 1) * (g[3] - g[1] + 1)
    den = float(aA + aB - inter)
    return 0.0 if den == 0 else inter / den


def compute_iou_matrix(rows, cols) -> np.ndarray:
    """eval_iground.py:58-63 (float64 matrix)."""
    m = np.zeros((len(rows), len(cols)))
    for i, r in enumerate(rows):
        for j, c in enumerate(cols):
            m[i, j] = compute_iou_plus1(r, c)
    return m


def greedy_match(ious: np.ndarray, sims: np.ndarray, iou_thr: float, sim_thr: float):
    """eval_iground.py:85-96 (find_best_matches loop): repeatedly take the global first-max of the IoU
    matrix (np.argmax, row-major tie-break), stop when it is '<' either threshold, zero its row+col."""
    ious, sims = ious.copy(), sims.copy()
    out = []
    while ious.size > 0 and sims.size > 0:
        i, j = np.unravel_index(np.argmax(ious), ious.shape)
        if ious[i, j] < iou_thr or sims[i, j] < sim_thr:
            break
        out.append((int(i), int(j)))
        ious[i, :] = 0
        ious[:, j] = 0
        sims[i, :] = 0
        sims[:, j] = 0
    return out


def bbox_overlaps_batch(anchors: np.ndarray, gt: np.ndarray, frm_mask=None) -> np.ndarray:
    """eval_anet.py:22-119, 3-D branch (anchors [b,N,5], gt [b,K,5], frm_mask [b,N,K] with 1 = different
    frame): fp32, +1 convention, zero-area gt -> 0, zero-area anchor -> -1 (applied in that order)."""
    a = anchors.astype(np.float32)
    g = gt.astype(np.float32)
    gx = g[:, :, 2] - g[:, :, 0] + 1
    gy = g[:, :, 3] - g[:, :, 1] + 1
    ax = a[:, :, 2] - a[:, :, 0] + 1
    ay = a[:, :, 3] - a[:, :, 1] + 1
    g_area = (gx * gy)[:, None, :]
    a_area = (ax * ay)[:, :, None]
    iw = np.minimum(a[:, :, None, 2], g[:, None, :, 2]) - np.maximum(a[:, :, None, 0], g[:, None, :, 0]) + 1
    iw[iw < 0] = 0
    ih = np.minimum(a[:, :, None, 3], g[:, None, :, 3]) - np.maximum(a[:, :, None, 1], g[:, None, :, 1]) + 1
    ih[ih < 0] = 0
    ua = a_area + g_area - iw * ih
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = iw * ih / ua
    if frm_mask is not None:
        ov = ov * (1 - frm_mask).astype(np.float32)
    gz = ((gx == 1) & (gy == 1))[:, None, :]
    az = ((ax == 1) & (ay == 1))[:, :, None]
    ov = np.where(np.broadcast_to(gz, ov.shape), np.float32(0), ov)
    ov = np.where(np.broadcast_to(az, ov.shape), np.float32(-1), ov)
    return ov.astype(np.float32)


def center_in_box(pred_xyxy, gt_xyxy) -> bool:
    """eval_youcookinteractions.py:43-48: inclusive centre-in-box test."""
    cx = (pred_xyxy[0] + pred_xyxy[2]) / 2
    cy = (pred_xyxy[1] + pred_xyxy[3]) / 2
    return bool(gt_xyxy[0] <= cx <= gt_xyxy[2] and gt_xyxy[1] <= cy <= gt_xyxy[3])


def sliding_segment_with_mask(num_frames=48, num_segments=8):
    """infer_iground.py:110-148: strided sparse windows + first-seen masks."""
    seg, rem = num_frames // num_segments, num_frames % num_segments
    all_idx, masks, seen = [], [], set()
    for off in range(seg):
        idx = [i * seg + off for i in range(num_segments)]
        masks.append([0 if k in seen else 1 for k in idx])
        all_idx.append(idx)
        seen.update(idx)
    for off in range(rem):
        idx = [k for k in (i * seg + seg + off for i in range(num_segments)) if k < num_frames]
        if idx:
            masks.append([0 if k in seen else 1 for k in idx])
            all_idx.append(idx)
            seen.update(idx)
    return all_idx, masks
