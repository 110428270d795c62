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


def center_in_box(pred_boxes, gt_boxes) -> np.ndarray:
    """The decision of eval_youcookinteractions.py:43-48 for n (prediction, ground truth) pairs of xyxy boxes: uint8 [n], 1 iff the centre
    of the prediction lies inside the ground-truth box, bounds INCLUSIVE; Python-float (double) arithmetic, NaN predictions -> 0."""
    p = torch.from_numpy(np.ascontiguousarray(np.asarray(pred_boxes, dtype=np.float64).reshape(-1, 4))).to(_dev())
    g = torch.from_numpy(np.ascontiguousarray(np.asarray(gt_boxes, dtype=np.float64).reshape(-1, 4))).to(_dev())
    assert p.shape == g.shape
    return ops.center_in_box(p, g).cpu().numpy()


def evaluate_dataset_localization(pred_boxes_dict, gt_data, dataset):
    """eval_youcookinteractions.py:8-51, same arguments and return value (accuracy %, total_correct, total_valid): frames without a
    ground-truth box are skipped, a missing (None) or NaN prediction counts as valid but wrong, the first box of a prediction is used."""
    pred, gt = [], []
    total_valid = 0
    for clip in gt_data:
        uid = f"{clip['video_id']}_{clip[f'segment_{dataset}_idx']}"
        pred_boxes = pred_boxes_dict.get(uid, []).get("final_boxes", [])
        for pb, gb in zip(pred_boxes, clip["segment_bboxes"]):
            if not gb:
                continue
            total_valid += 1
            if pb is None or np.any(np.isnan(pb)):
                continue
            pred.append(np.asarray(pb[0], dtype=np.float64)[:4])
            gt.append(np.asarray(gb, dtype=np.float64)[:4])
    total_correct = int(center_in_box(np.stack(pred), np.stack(gt)).sum()) if pred else 0
    accuracy = (total_correct / total_valid) * 100 if total_valid else 0.0
    return accuracy, total_correct, total_valid


def viou_over_threshold(pred_boxes, gt_boxes, iou_thresholds=(0.3, 0.5)):
    """One video of VidSTGiouEvaluator.evaluate (eval_vidstg.py:157-186): pred_boxes / gt_boxes [n,4] xyxy, one row per ground-truth frame in
    frame order (an all-zero prediction scores 0).  Returns (gt_viou, {thr: 0|1 with the reference's strict '>'}, per-frame IoUs)."""
    p = torch.from_numpy(np.ascontiguousarray(np.asarray(pred_boxes, dtype=np.float64).reshape(-1, 4))).to(_dev())
    g = torch.from_numpy(np.ascontiguousarray(np.asarray(gt_boxes, dtype=np.float64).reshape(-1, 4))).to(_dev())
    assert p.shape == g.shape
    thr = torch.tensor(list(iou_thresholds), dtype=torch.float64, device=_dev())
    ious, viou, over = ops.viou_decisions(p, g, thr)
    over = over.cpu().tolist()
    return float(viou.item()), {t: int(o) for t, o in zip(iou_thresholds, over)}, ious.cpu().numpy()


def val_giou_and_objectness_accuracy(pred_bboxes, logits_temp_objectness, gt_bboxes, gt_temp_objectness):
    """The validation sums of train.py:821-840 over nested [V][T] lists: (giou_sum, temp_objectness_sum, num_bboxes, num_max_bboxes).
    giou_sum adds torchvision's GIoU loss of pred[gt_objectness.bool()] vs the ground truth ON THE COORDINATES AS GIVEN -- the reference feeds
    cxcywh predictions (and `.int()`-cast ground truth) straight in; that quirk is reproduced, not corrected.  temp_objectness_sum counts
    (sigmoid(logit) > 0.5) == label and is exact."""
    dev = _dev()
    pb, lg, gt_rows, sel_rows, lab_rows = [], [], [], [], []
    num_bboxes = num_max = 0
    for v, (pv, lv) in enumerate(zip(pred_bboxes, logits_temp_objectness)):
        for f, (pf, lf) in enumerate(zip(pv, lv)):
            pf, lf = torch.as_tensor(pf), torch.as_tensor(lf)
            go = torch.as_tensor(gt_temp_objectness[v][f]).detach().cpu().to(torch.int32)
            gb = torch.as_tensor(gt_bboxes[v][f]).detach().cpu().float().reshape(-1, 4)
            g_full = torch.zeros(pf.shape[0], 4)
            g_full[go.bool()] = gb
            pb.append(pf.detach().float().reshape(-1, 4)); lg.append(lf.detach().float().reshape(-1))
            gt_rows.append(g_full); sel_rows.append(go.bool().to(torch.uint8)); lab_rows.append(go)
            num_bboxes += gb.shape[0]
            num_max += pf.shape[0]
    if not pb:
        return 0.0, 0, 0, 0
    giou, acc = ops.val_metrics(torch.cat(pb).to(dev).contiguous(), torch.cat(lg).to(dev).contiguous(), torch.cat(gt_rows).to(dev),
                                torch.cat(sel_rows).to(dev), torch.cat(lab_rows).to(dev))
    return giou, acc, num_bboxes, num_max
