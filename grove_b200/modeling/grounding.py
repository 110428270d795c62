"""The grounding half of model/GROVE.py behind the same method names.

`GroundingBranch` owns what GROVEBaseModel owns for this path — `grounding_encoder` (build_sam_vit_*) and
`text_hidden_fcs` (GROVE.py:55,75-79) — and exposes `get_grounding_encoder_embs`, `_create_det_token_mask`,
`_process_hidden_states`, `_generate_and_postprocess_masks`, `_compute_loss_components_video` with the reference's
argument meaning, so model/GROVE.py can delegate to it while the LLaVA decoder stays untouched.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from .build_sam import sam_model_registry
from .common import PackCache, bf16, f32


class GroundingBranch(nn.Module):
    def __init__(self, vit: str = "vit_h", hidden_size: int = 4096, out_dim: int = 256, num_frames: int = 8, use_temp_objectness: bool = True,
                 temp_objectness_threshold: float = 0.5, det_token_idx: int = 32005, ce_loss_weight: float = 1.0, giou_loss_weight: float = 2.0,
                 temp_objectness_loss_weight: float = 2.0, vision_pretrained: Optional[str] = None, image_size: Optional[int] = None):
        super().__init__()
        self.config = SimpleNamespace(hidden_size=hidden_size, out_dim=out_dim, num_frames=num_frames, use_temp_objectness=use_temp_objectness,
                                      temp_objectness_threshold=temp_objectness_threshold)
        self.det_token_idx = det_token_idx
        self.ce_loss_weight, self.giou_loss_weight, self.temp_objectness_loss_weight = ce_loss_weight, giou_loss_weight, temp_objectness_loss_weight
        self.grounding_encoder = sam_model_registry[vit](vision_pretrained, use_temp_objectness=use_temp_objectness, image_size=image_size)
        # GROVE.py:75-79
        self.text_hidden_fcs = nn.ModuleList([nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(inplace=True),
                                                            nn.Linear(hidden_size, out_dim), nn.Dropout(0.0))])
        self._pack = PackCache()

    # GROVE.py:134-136
    def get_grounding_encoder_embs(self, images: torch.Tensor):
        return self.grounding_encoder.image_encoder(images)

    # GROVE.py:200-205 (right_pad=1: teacher-forced) / :427-430 (right_pad=0: generate)
    def _create_det_token_mask(self, input_ids: torch.Tensor, right_pad: int = 1):
        mask = input_ids[:, 1:] == self.det_token_idx
        parts = [torch.zeros((mask.shape[0], 575), dtype=torch.bool, device=mask.device), mask]
        if right_pad:
            parts.append(torch.zeros((mask.shape[0], right_pad), dtype=torch.bool, device=mask.device))
        return torch.cat(parts, dim=1)

    # GROVE.py:248-268
    @torch.no_grad()
    def _process_hidden_states(self, output_hidden_states, det_token_mask, offset=None, infer=False):
        """Returns (hidden_states, pred_embeddings_list) like the reference.  The reference projects all V*L tokens and
        keeps the [DET] rows; here the rows are gathered first and only they go through text_hidden_fcs (identical values
        for the kept rows).  `hidden_states` is therefore `[projected [DET] rows]`, not the full [V,L,out_dim] tensor —
        no caller reads it (GROVE.py:179,432)."""
        hidden = output_hidden_states[-1]
        V, L, Hd = hidden.shape
        T = self.config.num_frames
        idx = det_token_mask.reshape(-1).nonzero().flatten().to(torch.int32)          # row-major order == boolean-mask order
        counts = det_token_mask.int().sum(-1).tolist()                                 # host sync, as in the reference's slicing loop
        n = idx.numel()
        fcs = self.text_hidden_fcs[0]
        out_dim = fcs[2].out_features
        dev = hidden.device
        if n == 0:
            proj = torch.zeros(0, out_dim, device=dev, dtype=torch.float32)
        else:
            rows = ((n + 127) // 128) * 128
            a = torch.zeros(rows, Hd, device=dev, dtype=torch.bfloat16)
            ops.gather_rows_bf16(hidden.reshape(V * L, Hd).contiguous(), idx, a)
            w0 = self._pack.get("fc0.w", [fcs[0].weight], bf16); b0 = self._pack.get("fc0.b", [fcs[0].bias], f32)
            w2 = self._pack.get("fc2.w", [fcs[2].weight], bf16); b2 = self._pack.get("fc2.b", [fcs[2].bias], f32)
            h = torch.empty(rows, Hd, device=dev, dtype=torch.bfloat16)
            ops.gemm(a, w0, h, bias=b0, act="relu")
            p = torch.empty(rows, out_dim, device=dev, dtype=torch.float32)
            ops.gemm(h, w2, p, bias=b2)
            proj = p[:n]
        proj = proj.to(hidden.dtype)
        # repeat_interleave(num_frames) of both hidden states and mask (:253-254): every frame of video v gets v's rows
        per_video, s = [], 0
        for c in counts:
            per_video.append(proj[s:s + c])
            s += c
        pred_embeddings_list = [pv for pv in per_video for _ in range(T)]
        return [proj], pred_embeddings_list

    # GROVE.py:270-331
    @torch.no_grad()
    def _generate_and_postprocess_masks(self, pred_embeddings, image_embeddings, orig_sizes, dense_pe, infer=False):
        T = self.config.num_frames
        bs = len(pred_embeddings)
        num_masks_per_embed = [e.shape[0] for e in pred_embeddings]
        pred = torch.cat(pred_embeddings, dim=0).unsqueeze(1)
        ge = self.grounding_encoder
        sparse, dense = ge.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=pred)
        sparse = sparse.to(pred.dtype)
        out = ge.mask_decoder(image_embeddings=image_embeddings, image_pe=dense_pe, sparse_prompt_embeddings=sparse,
                              dense_prompt_embeddings=dense, multimask_output=False, reps=num_masks_per_embed)
        if self.config.use_temp_objectness:
            bbox_preds, logits = out
        else:
            bbox_preds, logits = out, None
        B = bbox_preds.shape[0]
        if infer:
            sizes = []
            for i in range(bs):
                w, h = orig_sizes[i // T]
                sizes += [[float(w), float(h)]] * num_masks_per_embed[i]
            size_wh = torch.tensor(sizes, dtype=torch.float32, device=bbox_preds.device).reshape(B, 2)
            lg = logits.float().contiguous() if logits is not None else torch.full((B,), 1e9, device=bbox_preds.device)
            xyxy, keep = ops.box_postprocess(bbox_preds.float().contiguous(), lg, size_wh, self.config.temp_objectness_threshold)
            xyxy = xyxy.to(bbox_preds.dtype)
            keep = keep.bool()
        bbox_pred_list, logit_list, s = [], [], 0
        for i in range(0, bs, T):
            fb, fl = [], []
            for j in range(T):
                n = num_masks_per_embed[i + j]
                if infer:
                    fb.append(xyxy[s:s + n][keep[s:s + n]] if logits is not None else xyxy[s:s + n])
                else:
                    fb.append(bbox_preds[s:s + n])
                if logits is not None:
                    fl.append(logits[s:s + n])
                s += n
            bbox_pred_list.append(fb)
            logit_list.append(fl)
        return (bbox_pred_list, logit_list) if self.config.use_temp_objectness else bbox_pred_list

    # GROVE.py:339-381
    @torch.no_grad()
    def _compute_loss_components_video(self, pred_bboxes, logits_temp_objectness, gt_bboxes_list, gt_temp_objectness_list, output):
        ce_loss = output.loss * self.ce_loss_weight
        dev = ce_loss.device
        if not (self.config.use_temp_objectness and logits_temp_objectness is not None):
            raise NotImplementedError("grove_b200 builds the use_temp_objectness=True loss (the configuration every GROVE script uses)")
        pb, lg, gt_rows, sel_rows, lab_rows = [], [], [], [], []
        num_bboxes = num_max = 0
        for v, (pv, lv) in enumerate(zip(pred_bboxes, logits_temp_objectness)):
            for f, (pf, lf) in enumerate(zip(pv, lv)):
                gb = torch.as_tensor(gt_bboxes_list[v][f]).detach().cpu().float().reshape(-1, 4)
                go = torch.as_tensor(gt_temp_objectness_list[v][f]).detach().cpu()
                assert gb.shape[0] == go.sum(), f"Number of ground truth bboxes and objectness labels do not match: {gb.shape[0]} vs {go.sum()}"
                sel = go.bool()
                g_full = torch.zeros(pf.shape[0], 4)
                g_full[sel] = gb
                pb.append(pf); lg.append(lf); gt_rows.append(g_full); sel_rows.append(sel.to(torch.uint8)); lab_rows.append(go.float())
                num_bboxes += gb.shape[0]
                num_max += pf.shape[0]
        boxes = torch.cat(pb).float().contiguous()
        sums = ops.box_losses(boxes, torch.cat(lg).float().contiguous(), torch.cat(gt_rows).to(dev), torch.cat(sel_rows).to(dev),
                              torch.cat(lab_rows).to(dev))
        giou = self.giou_loss_weight * sums[0] / (num_bboxes + 1e-8)
        l1 = self.giou_loss_weight * sums[1] / (num_bboxes + 1e-8)          # the L1 term reuses the GIoU weight (GROVE.py:375)
        obj = self.temp_objectness_loss_weight * sums[2] / (num_max + 1e-8)
        return {"loss": ce_loss + giou + l1 + obj, "ce_loss": ce_loss, "giou_loss": giou, "l1_loss": l1, "temp_objectness_loss": obj}

    # the grounding half of model_forward / evaluate (GROVE.py:162-186, 432-444)
    @torch.no_grad()
    def ground(self, images, last_hidden_state, det_token_mask, orig_sizes=None, infer=False):
        emb = self.get_grounding_encoder_embs(images)
        _, pred = self._process_hidden_states([last_hidden_state], det_token_mask, None, infer=infer)
        dense_pe = self.grounding_encoder.prompt_encoder.get_dense_pe()
        return emb, self._generate_and_postprocess_masks(pred, emb, orig_sizes, dense_pe, infer=infer)
