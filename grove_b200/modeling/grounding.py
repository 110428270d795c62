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
from .common import PackCache, bf16, f32, no_grad_entry


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
        self._use_graphs = False
        self._graphs = {}
        self._row_cache = {}

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

    # ------------------------------------------------------------------ [DET] rows: host-side bookkeeping, device-side projection
    @staticmethod
    def _det_rows(det_token_mask: torch.Tensor):
        """(idx, counts): flat row indices (int32, HOST tensor, row-major = boolean-mask order) of the [DET] positions of a [V, L] mask and
        the number per video.  A CUDA mask costs one device->host read of V*L bytes; the serving loop builds the mask from host-resident
        `input_ids` and pays nothing.  Either way this runs BEFORE any kernel of the step is enqueued, so nothing downstream blocks the host
        between the encoder and the decoder (the reference slices device tensors in a Python loop, GROVE.py:262-267)."""
        m = det_token_mask
        if m.is_cuda:
            m = m.cpu()
        idx = m.reshape(-1).nonzero().flatten().to(torch.int32)
        return idx, [int(c) for c in m.sum(-1).tolist()]

    def _project_rows(self, hidden: torch.Tensor, idx_dev: torch.Tensor, n: int) -> torch.Tensor:
        """text_hidden_fcs (GROVE.py:75-79) on the n gathered [DET] rows only -> fp32 [n, out_dim] (identical values for the kept rows)"""
        V, L, Hd = hidden.shape
        fcs = self.text_hidden_fcs[0]
        out_dim = fcs[2].out_features
        dev = hidden.device
        if n == 0:
            return torch.zeros(0, out_dim, device=dev, dtype=torch.float32)
        rows = ((n + 127) // 128) * 128
        a = torch.zeros(rows, Hd, device=dev, dtype=torch.bfloat16)
        ops.gather_rows_bf16(hidden.reshape(V * L, Hd).contiguous(), idx_dev, a)
        w0 = self._pack.get("fc0.w", [fcs[0].weight], bf16); b0 = self._pack.get("fc0.b", [fcs[0].bias], f32)
        w2 = self._pack.get("fc2.w", [fcs[2].weight], bf16); b2 = self._pack.get("fc2.b", [fcs[2].bias], f32)
        h = torch.empty(rows, Hd, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, w0, h, bias=b0, act="relu")
        p = torch.empty(rows, out_dim, device=dev, dtype=torch.float32)
        ops.gemm(h, w2, p, bias=b2)
        return p[:n]

    # GROVE.py:248-268
    @no_grad_entry("GroundingBranch._process_hidden_states", lambda self: self.text_hidden_fcs.parameters())
    def _process_hidden_states(self, output_hidden_states, det_token_mask, offset=None, infer=False):
        """Returns (hidden_states, pred_embeddings_list) like the reference.  The reference projects all V*L tokens and
        keeps the [DET] rows; here the rows are gathered first and only they go through text_hidden_fcs (identical values
        for the kept rows).  `hidden_states` is therefore `[projected [DET] rows]`, not the full [V,L,out_dim] tensor —
        no caller reads it (GROVE.py:179,432)."""
        hidden = output_hidden_states[-1]
        T = self.config.num_frames
        idx, counts = self._det_rows(det_token_mask)
        with ops.device_of(hidden):
            proj = self._project_rows(hidden, idx.to(hidden.device), idx.numel()).to(hidden.dtype)
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
        return self._nested_outputs(bbox_preds, logits, num_masks_per_embed, orig_sizes, infer)

    def _postprocess(self, bbox_preds, logits, reps, orig_sizes):
        """GROVE.py:307-315 for all instances at once: boxes x (w, h) -> xyxy, keep = sigmoid(logit) > threshold.  Returns (xyxy fp32 [B,4],
        keep uint8 [B]); without objectness every box is kept."""
        T = self.config.num_frames
        B = bbox_preds.shape[0]
        if orig_sizes is None:
            raise ValueError("infer=True needs orig_sizes = [(w, h)] per video (GROVE.py:307)")
        sizes = []
        for i, n in enumerate(reps):
            w, h = orig_sizes[i // T]
            sizes += [[float(w), float(h)]] * n
        dev = bbox_preds.device
        size_wh = torch.tensor(sizes, dtype=torch.float32).reshape(B, 2).to(dev, non_blocking=True)
        lg = logits.float().contiguous() if logits is not None else torch.full((B,), 1e9, device=dev)
        with ops.device_of(bbox_preds):
            return ops.box_postprocess(bbox_preds.float().contiguous(), lg, size_wh, self.config.temp_objectness_threshold)

    def _nested_outputs(self, bbox_preds, logits, reps, orig_sizes, infer, post=None):
        """the slicing loop of GROVE.py:297-331: nested [V][T] lists of views into the flat [B,4] / [B] outputs.  `post` = (xyxy, keep)
        when the caller already ran the post-process (ground() thresholds the fp32 records, not their bf16 rounding)."""
        T = self.config.num_frames
        bs = len(reps)
        if infer:
            xyxy, keep = post if post is not None else self._postprocess(bbox_preds, logits, reps, orig_sizes)
            xyxy = xyxy.to(bbox_preds.dtype)
            keep = keep.bool()
        bbox_pred_list, logit_list, s = [], [], 0
        for i in range(0, bs, T):
            fb, fl = [], []
            for j in range(T):
                n = reps[i + j]
                if infer:
                    fb.append(xyxy[s:s + n][keep[s:s + n]] if logits is not None else xyxy[s:s + n])
                else:
                    fb.append(bbox_preds[s:s + n])
                if logits is not None:
                    fl.append(logits[s:s + n])       # all P logits of the frame, also when infer drops boxes (GROVE.py:316)
                s += n
            bbox_pred_list.append(fb)
            logit_list.append(fl)
        return (bbox_pred_list, logit_list) if self.config.use_temp_objectness else bbox_pred_list

    # GROVE.py:339-408
    def _compute_loss_components_video(self, pred_bboxes, logits_temp_objectness, gt_bboxes_list, gt_temp_objectness_list, output):
        """Both branches of the reference: with objectness (GROVE.py:341-381) and without (:382-408; GIoU + L1 only, the ground-truth
        objectness labels still select the predictions that have a box)."""
        if torch.is_grad_enabled() and any(getattr(x, "requires_grad", False) for v in pred_bboxes for x in v):
            raise RuntimeError("grove_b200: _compute_loss_components_video does not build an autograd graph; train through "
                               "GroundingBranch.grounding_loss(...) (the CUDA library's own backward pass)")
        with torch.no_grad():
            return self._loss_components(pred_bboxes, logits_temp_objectness, gt_bboxes_list, gt_temp_objectness_list, output)

    def _loss_components(self, pred_bboxes, logits_temp_objectness, gt_bboxes_list, gt_temp_objectness_list, output):
        ce_loss = output.loss * self.ce_loss_weight
        dev = ce_loss.device
        with_obj = bool(self.config.use_temp_objectness and logits_temp_objectness is not None)
        pb, lg, gt_rows, sel_rows, lab_rows = [], [], [], [], []
        num_bboxes = num_max = 0
        for v, pv in enumerate(pred_bboxes):
            for f, pf in enumerate(pv):
                gb = torch.as_tensor(gt_bboxes_list[v][f]).detach().cpu().float().reshape(-1, 4)
                go = torch.as_tensor(gt_temp_objectness_list[v][f]).detach().cpu()
                assert gb.shape[0] == go.sum(), f"Number of ground truth bboxes and objectness labels do not match: {gb.shape[0]} vs {go.sum()}"
                sel = go.bool()
                g_full = torch.zeros(pf.shape[0], 4)
                g_full[sel] = gb
                pb.append(pf); gt_rows.append(g_full); sel_rows.append(sel.to(torch.uint8)); lab_rows.append(go.float())
                lg.append(logits_temp_objectness[v][f] if with_obj else torch.zeros(pf.shape[0], device=pf.device))
                num_bboxes += gb.shape[0]
                num_max += pf.shape[0]
        boxes = torch.cat(pb).float().contiguous()
        with ops.device_of(boxes):
            sums = ops.box_losses(boxes, torch.cat(lg).float().contiguous(), torch.cat(gt_rows).to(dev), torch.cat(sel_rows).to(dev),
                                  torch.cat(lab_rows).to(dev))
        giou = self.giou_loss_weight * sums[0] / (num_bboxes + 1e-8)
        l1 = self.giou_loss_weight * sums[1] / (num_bboxes + 1e-8)          # the L1 term reuses the GIoU weight (GROVE.py:375,403)
        if not with_obj:
            return {"loss": ce_loss + giou + l1, "ce_loss": ce_loss, "giou_loss": giou, "l1_loss": l1}
        obj = self.temp_objectness_loss_weight * sums[2] / (num_max + 1e-8)
        return {"loss": ce_loss + giou + l1 + obj, "ce_loss": ce_loss, "giou_loss": giou, "l1_loss": l1, "temp_objectness_loss": obj}

    # ------------------------------------------------------------------ training step (BASELINE config 4)
    def _loss_inputs(self, reps, T, gt_bboxes_list, gt_temp_objectness_list, dev):
        """flatten the nested ground truth of _compute_loss_components_video (GROVE.py:350-371) into per-instance rows"""
        gt_rows, sel_rows, lab_rows = [], [], []
        num_bboxes = num_max = 0
        i = 0
        for v in range(len(gt_bboxes_list)):
            for f in range(T):
                n = reps[i]; i += 1
                gb = torch.as_tensor(gt_bboxes_list[v][f]).detach().cpu().float().reshape(-1, 4)
                go = torch.as_tensor(gt_temp_objectness_list[v][f]).detach().cpu()
                assert go.numel() == n, f"objectness labels ({go.numel()}) do not match the number of predictions ({n})"
                assert gb.shape[0] == go.sum(), f"Number of ground truth bboxes and objectness labels do not match: {gb.shape[0]} vs {go.sum()}"
                sel = go.bool()
                g_full = torch.zeros(n, 4)
                g_full[sel] = gb
                gt_rows.append(g_full); sel_rows.append(sel.to(torch.uint8)); lab_rows.append(go.float())
                num_bboxes += gb.shape[0]
                num_max += n
        return torch.cat(gt_rows).to(dev), torch.cat(sel_rows).to(dev), torch.cat(lab_rows).to(dev), num_bboxes, num_max

    @torch.no_grad()
    def grounding_loss_and_grads(self, images, last_hidden_state, det_token_mask, gt_bboxes_list, gt_temp_objectness_list, ce_loss=None,
                                 upstream: float = 1.0, apply: bool = True, _cotangent=None, reducer=None):
        """One training step of the grounding branch — model_forward's grounding half + _compute_loss_components_video + backward
        (GROVE.py:162-198, 339-381; train.py:761-770) — on the CUDA library, without autograd.

        Returns (losses, d_last_hidden_state, grads): `losses` has the reference's keys; `d_last_hidden_state` ([V,L,hidden], fp32) is
        the cotangent handed back to the language model; `grads` (GradStore) holds fp32 gradients of every trainable parameter of the
        adapters, the mask decoder and text_hidden_fcs.  With apply=True they are also added to `.grad` (scaled by `upstream`)."""
        from .decoder_train import GradStore, _wants
        from .encoder_train import encode_backward, encode_train
        if not self.config.use_temp_objectness:
            raise NotImplementedError("grove_b200 builds the use_temp_objectness=True loss (the configuration every GROVE script uses)")
        ge = self.grounding_encoder
        dev = images.device
        T = self.config.num_frames
        grads = GradStore(getattr(self, "_grad_arena_elems", 0), next(self.parameters()).device)
        # ---- forward
        emb_tok, enc_tape = encode_train(ge.image_encoder, images)
        V, L, Hd = last_hidden_state.shape
        idx = det_token_mask.reshape(-1).nonzero().flatten().to(torch.int32)
        counts = det_token_mask.int().sum(-1).tolist()
        n = idx.numel()
        if n == 0:
            raise ValueError("the training step needs at least one [DET] token")
        fcs = self.text_hidden_fcs[0]
        out_dim = fcs[2].out_features
        rows = ((n + 127) // 128) * 128
        a = torch.zeros(rows, Hd, device=dev, dtype=torch.bfloat16)
        ops.gather_rows_bf16(last_hidden_state.reshape(V * L, Hd).contiguous(), idx, a)
        w0 = self._pack.get("fc0.w", [fcs[0].weight], bf16); b0 = self._pack.get("fc0.b", [fcs[0].bias], f32)
        w2 = self._pack.get("fc2.w", [fcs[2].weight], bf16); b2 = self._pack.get("fc2.b", [fcs[2].bias], f32)
        h = torch.empty(rows, Hd, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, w0, h, bias=b0, act="relu")
        p = torch.empty(rows, out_dim, device=dev, dtype=torch.float32)
        ops.gemm(h, w2, p, bias=b2)
        proj = p[:n].to(last_hidden_state.dtype).float()           # the module's output dtype (bf16 in production), as the reference
        # instance b = (video v, frame t, phrase j) reads projected row offset_v + j  (repeat_interleave over frames, GROVE.py:253-257)
        row_of, reps, s0 = [], [], 0
        for c in counts:
            for _ in range(T):
                row_of += list(range(s0, s0 + c))
                reps.append(c)
            s0 += c
        Fr = emb_tok.shape[0]
        if len(reps) != Fr:
            raise ValueError(f"{V} videos x {T} frames do not match the {Fr} encoded frames")
        row_of_t = torch.tensor(row_of, dtype=torch.long, device=dev)
        text = proj[row_of_t].contiguous()                          # [B, out_dim]
        B = text.shape[0]
        no_mask = ge.prompt_encoder.no_mask_embed.weight.reshape(-1).to(torch.float32).contiguous()
        dense_pe = ge.prompt_encoder.get_dense_pe()
        boxes, logits, dec_tape = ge.mask_decoder.predict_masks_train(emb_tok, dense_pe, text, no_mask, reps)
        # ---- losses (GROVE.py:339-381) and their derivative
        gt, sel, lab, num_bboxes, num_max = self._loss_inputs(reps, T, gt_bboxes_list, gt_temp_objectness_list, dev)
        sums = ops.box_losses(boxes, logits, gt, sel, lab)
        gw, ow = self.giou_loss_weight, self.temp_objectness_loss_weight
        giou = gw * sums[0] / (num_bboxes + 1e-8)
        l1 = gw * sums[1] / (num_bboxes + 1e-8)                     # the L1 term reuses the GIoU weight (GROVE.py:375)
        obj = ow * sums[2] / (num_max + 1e-8)
        ce = (ce_loss if ce_loss is not None else torch.zeros((), device=dev)) * self.ce_loss_weight
        losses = {"loss": ce + giou + l1 + obj, "ce_loss": ce, "giou_loss": giou, "l1_loss": l1, "temp_objectness_loss": obj}
        dboxes, dlogits = ops.box_losses_bwd(boxes, logits, gt, sel, lab, gw / (num_bboxes + 1e-8), ow / (num_max + 1e-8))
        if _cotangent is not None:      # tests: a prescribed cotangent of (boxes, logits) instead of the loss' own — a pure VJP check
            dboxes, dlogits = _cotangent[0].float().contiguous(), _cotangent[1].float().contiguous()
        losses["boxes"], losses["logits"] = boxes, logits
        # ---- backward: decoder -> encoder (adapters) and text projection
        d_emb, d_text = ge.mask_decoder.backward(dec_tape, dboxes, dlogits, grads)
        del dec_tape
        on_ready = reducer.ready if reducer is not None else None
        if on_ready is not None:
            on_ready([grads.g[p] for p in ge.mask_decoder.parameters() if p in grads.g])     # the decoder's gradients are final
        encode_backward(ge.image_encoder, enc_tape, d_emb, grads, on_ready=on_ready)
        del enc_tape
        onehot = torch.zeros(B, n, device=dev, dtype=torch.float32)
        onehot[torch.arange(B, device=dev), row_of_t] = 1.0
        d_p = torch.zeros(rows, out_dim, device=dev, dtype=torch.float32)
        ops.small_wgrad(onehot, d_text, d_p[:n])                    # d proj[row] = sum of the cotangents of the instances that read it
        d_p16 = d_p.to(torch.bfloat16)
        if _wants(fcs[2].weight):
            ops.wgrad(d_p16, h, grads.buf(fcs[2].weight))
        if _wants(fcs[2].bias):
            ops.colsum(d_p, grads.buf(fcs[2].bias))
        w2t = self._pack.get("fc2.wT", [fcs[2].weight], lambda w: bf16(w.t()))
        d_h = torch.empty(rows, Hd, device=dev, dtype=torch.bfloat16)
        ops.gemm(d_p16, w2t, d_h, dact_pre=h, dact="relu")          # ReLU mask from its (saved) output
        if _wants(fcs[0].weight):
            ops.wgrad(d_h, a, grads.buf(fcs[0].weight))
        if _wants(fcs[0].bias):
            ops.colsum(d_h, grads.buf(fcs[0].bias))
        w0t = self._pack.get("fc0.wT", [fcs[0].weight], lambda w: bf16(w.t()))
        d_a = torch.empty(rows, Hd, device=dev, dtype=torch.float32)
        ops.gemm(d_h, w0t, d_a)
        d_hidden = torch.zeros(V * L, Hd, device=dev, dtype=torch.float32)
        d_hidden[idx.long()] = d_a[:n]
        if upstream != 1.0:
            d_hidden *= upstream
        if reducer is not None:
            reducer.ready([grads.g[p] for p in self.text_hidden_fcs.parameters() if p in grads.g])
            reducer.finish()
        if apply:
            grads.apply(upstream)
        self._grad_arena_elems = grads.elems                        # the next step's accumulators are views of one zero-filled buffer
        return losses, d_hidden.view(V, L, Hd), grads

    def grounding_loss(self, images, last_hidden_state, det_token_mask, gt_bboxes_list, gt_temp_objectness_list):
        """Autograd bridge: a scalar (giou + l1 + objectness, weighted as GROVE.py:372-378) whose .backward() delivers the CUDA
        library's gradients to the trainable grounding parameters and to `last_hidden_state` (and on into the language model) —
        the drop-in for `model.backward(loss)` (train.py:770)."""
        params = [p for p in list(self.grounding_encoder.image_encoder.adapters.parameters()) + list(self.grounding_encoder.mask_decoder.parameters())
                  + list(self.text_hidden_fcs.parameters()) if p.requires_grad]
        return _GroundingLossFn.apply(self, images, last_hidden_state, det_token_mask, gt_bboxes_list, gt_temp_objectness_list, *params)

    # ------------------------------------------------------------------ the grounding half of model_forward / evaluate (GROVE.py:162-186, 432-444)
    def enable_cuda_graphs(self, on: bool = True) -> None:
        """Serving mode: the WHOLE path of a step — patch im2col, encoder, [DET]-row gather + text projection, box decoder, heads — is captured
        once per (input shapes, [DET] counts, weights) as ONE CUDA graph and replayed; a step is then three small copies into static inputs
        and one graph launch, with no host work between kernels.  Re-captured when any parameter is reassigned, moved or modified in place;
        falls back to kernel-by-kernel launches if capture fails."""
        self._use_graphs = bool(on)
        self._graphs = {}

    def _row_index(self, counts, T, dev):
        """instance (video v, frame t, phrase j) reads projected row offset_v + j (repeat_interleave over frames, GROVE.py:253-257).
        Returns (device long tensor [B], reps list of len V*T); cached per counts."""
        key = (tuple(counts), T, dev)
        hit = self._row_cache.get(key)
        if hit is None:
            if len(self._row_cache) > 64:
                self._row_cache.clear()
            row_of, reps, s0 = [], [], 0
            for c in counts:
                for _ in range(T):
                    row_of += list(range(s0, s0 + c))
                    reps.append(c)
                s0 += c
            hit = (torch.tensor(row_of, dtype=torch.long).to(dev), reps)
            self._row_cache[key] = hit
        return hit

    def _ground_core(self, images, hidden, idx_dev, counts, records_out=None):
        """All kernels of one step, enqueued back to back with no host synchronisation: text projection of the [DET] rows, encoder, decoder +
        heads.  Returns (token-major embeddings bf16 [F, N, C], packed records fp32 [B, 5])."""
        ge = self.grounding_encoder
        T = self.config.num_frames
        row_of, reps = self._row_index(counts, T, hidden.device)
        proj = self._project_rows(hidden, idx_dev, sum(counts)).to(hidden.dtype)      # the module's output dtype (bf16 in production), as the reference
        text = proj.float()[row_of].contiguous() if row_of.numel() else proj.float()
        emb_tok = ge.image_encoder.forward_tokens(images)
        if len(reps) != emb_tok.shape[0]:
            raise ValueError(f"{len(counts)} videos x {T} frames do not match the {emb_tok.shape[0]} encoded frames")
        no_mask = ge.prompt_encoder.no_mask_embed.weight.reshape(-1).to(torch.float32).contiguous()
        rec = ge.mask_decoder.decode_records(emb_tok, ge.prompt_encoder.get_dense_pe(), text, no_mask, reps, records=records_out)
        return emb_tok, rec

    def _param_signature(self):
        """(data_ptr, version) of every parameter: an in-place update, a reload or module surgery all change it and trigger a re-capture.
        Walks `_parameters` / `_modules` directly -- nn.Module.parameters() de-duplicates through named_members and costs 3x as much
        (1.4 ms per step on the bench host, all of it in front of the graph launch)."""
        out, stack = [], [self]
        while stack:
            m = stack.pop()
            for p in m._parameters.values():
                if p is not None:
                    out.append((p.data_ptr(), p._version))
            stack.extend(m._modules.values())
        return tuple(out)

    def _ground_graphed(self, images, hidden, idx, counts, *, copy_out=True):
        """replay (capturing first if needed) the whole-step CUDA graph; `images` / `hidden` / `idx` may live on the host (pinned: the copies
        into the static inputs are the step's only host->device traffic)"""
        dev = next(self.parameters()).device
        key = (tuple(images.shape), images.dtype, tuple(hidden.shape), hidden.dtype, tuple(counts), dev)
        ent = self._graphs.get(key)
        copied = False
        if ent is not None:
            # the uploads into the static inputs go first (they do not depend on the weights): the ~0.5 ms of host work for the parameter
            # signature below then runs under the 0.9 ms image DMA instead of in front of it
            st_images, st_hidden, st_idx = ent["in"]
            st_images.copy_(images, non_blocking=True); st_hidden.copy_(hidden, non_blocking=True)
            if idx.numel():
                st_idx[:idx.numel()].copy_(idx, non_blocking=True)
            copied = True
        sig = self._param_signature()
        if ent is None or ent["sig"] != sig:
            copied = False
            enc = self.grounding_encoder.image_encoder
            st_images = torch.empty(images.shape, dtype=images.dtype, device=dev)
            st_hidden = torch.empty(hidden.shape, dtype=hidden.dtype, device=dev)
            st_idx = torch.zeros(max(idx.numel(), 1), dtype=torch.int32, device=dev)
            st_images.copy_(images, non_blocking=True); st_hidden.copy_(hidden, non_blocking=True); st_idx[:idx.numel()].copy_(idx, non_blocking=True)
            inner, enc._use_graphs = enc._use_graphs, False                  # the encoder's own replay cannot nest inside a capture
            try:
                self._ground_core(st_images, st_hidden, st_idx, counts)      # warm-up: packs weights, fills caches, sets kernel attributes
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    emb_tok, rec = self._ground_core(st_images, st_hidden, st_idx, counts)
                ent = {"sig": sig, "graph": graph, "in": (st_images, st_hidden, st_idx), "out": (emb_tok, rec), "kernels": ops.launch_count() - n0}
            except Exception as e:                                           # capture is an optimisation: fall back to plain launches
                import warnings
                warnings.warn(f"grove_b200: CUDA-graph capture of the grounding step failed ({e}); launching kernel by kernel")
                self._use_graphs = False
                torch.cuda.synchronize(dev)
                return None
            finally:
                enc._use_graphs = inner
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = ent
        if not copied:
            st_images, st_hidden, st_idx = ent["in"]
            st_images.copy_(images, non_blocking=True); st_hidden.copy_(hidden, non_blocking=True)
            if idx.numel():
                st_idx[:idx.numel()].copy_(idx, non_blocking=True)
        ent["graph"].replay()
        ops.add_launch_count(ent["kernels"])
        emb_tok, rec = ent["out"]
        return (emb_tok.clone(), rec.clone()) if copy_out else (emb_tok, rec)

    @torch.no_grad()
    def ground_records(self, images, last_hidden_state, det_token_mask, *, copy_out=True, records_out=None):
        """One step of the path on packed outputs: images [V,3,T,H,W], last_hidden_state [V,L,hidden], det_token_mask bool [V,L] (host or
        device) -> (token-major embeddings bf16 [V*T, N, C], records fp32 [B,5] = cx, cy, w, h, objectness logit; reps list of len V*T).
        With CUDA graphs enabled the inputs may be pinned host tensors.  copy_out=False returns the graph's static output buffers (valid until
        the next call).  `records_out` (fp32 [B,5], e.g. a slot of an all-gather buffer) receives the records: the heads kernel writes it
        directly on the kernel-by-kernel path, the graph path copies its static output there."""
        dev = next(self.parameters()).device
        idx, counts = self._det_rows(det_token_mask)
        reps = [c for c in counts for _ in range(self.config.num_frames)]
        if not last_hidden_state.is_cuda and 0 < idx.numel() < last_hidden_state.shape[0] * last_hidden_state.shape[1]:
            # host-resident hidden states: only the [DET] rows are ever read (text_hidden_fcs runs on the gathered rows), so gather them on
            # the host into a pinned buffer and upload n x hidden instead of V x L x hidden (5.2 MB -> 32 KB per 640-token sequence)
            n, Hd = idx.numel(), last_hidden_state.shape[-1]
            key = (n, Hd, last_hidden_state.dtype)
            ent = self._row_cache.get(("host_rows",) + key)
            if ent is None:
                ent = self._row_cache[("host_rows",) + key] = [torch.empty((n, Hd), dtype=last_hidden_state.dtype, pin_memory=True), None]
            buf, ev = ent
            if ev is not None:
                ev.synchronize()                                    # the previous upload from this pinned buffer has left the host
            torch.index_select(last_hidden_state.reshape(-1, Hd), 0, idx.long(), out=buf)
            with torch.cuda.device(dev):
                last_hidden_state = buf.to(dev, non_blocking=True).view(1, n, Hd)
                ent[1] = torch.cuda.Event()
                ent[1].record()
            idx = torch.arange(n, dtype=torch.int32)
        with torch.cuda.device(dev):
            if self._use_graphs:
                out = self._ground_graphed(images, last_hidden_state, idx, counts, copy_out=copy_out)
                if out is not None:
                    if records_out is not None:
                        records_out.copy_(out[1])
                        return out[0], records_out, reps
                    return out[0], out[1], reps
            images, last_hidden_state = images.to(dev, non_blocking=True), last_hidden_state.to(dev, non_blocking=True)
            emb_tok, rec = self._ground_core(images, last_hidden_state, idx.to(dev, non_blocking=True), counts, records_out)
        return emb_tok, rec, reps

    @torch.no_grad()
    def ground(self, images, last_hidden_state, det_token_mask, orig_sizes=None, infer=False):
        """model_forward's grounding half in the reference's return format: (image_embeddings [V*T, C, G, G], (nested boxes, nested logits))
        (GROVE.py:134-136, 179-186).  The [DET] bookkeeping happens before the first kernel is enqueued and the text projection runs first, so
        the host never waits between the encoder and the decoder."""
        emb_tok, rec, reps = self.ground_records(images, last_hidden_state, det_token_mask)
        Fr, N, C = emb_tok.shape
        G = int(round(N ** 0.5))
        emb = emb_tok.view(Fr, G, G, C).permute(0, 3, 1, 2)
        out_dtype = last_hidden_state.dtype
        boxes = rec[:, :4].to(out_dtype)
        logits = rec[:, 4].to(out_dtype) if self.config.use_temp_objectness else None
        post = self._postprocess(rec[:, :4], rec[:, 4] if logits is not None else None, reps, orig_sizes) if infer else None
        return emb, self._nested_outputs(boxes, logits, reps, orig_sizes, infer, post)

    @torch.no_grad()
    def ground_host_stream(self, host_batches, orig_sizes=None, infer=False):
        """End-to-end serving loop over HOST batches (the shape of infer_iground.py:150-295: decode on the host, ground on the GPU,
        collect numpy-ready results): yields one packed fp32 host tensor per batch, in order — `[sum P, 5]` = (cx, cy, w, h, objectness logit),
        or with infer=True `[sum P, 6]` = (x1, y1, x2, y2 in pixels of orig_sizes, objectness logit, keep) where keep = 1.0 iff
        sigmoid(logit) > temp_objectness_threshold (the rows the reference's infer branch returns, GROVE.py:307-315; rows are NOT dropped so
        the record count stays static).

        `host_batches` iterates `(images, last_hidden_state, input_ids)` in pinned host memory.  Batch i+1 is uploaded on a side
        stream while batch i is computed and the result of batch i is read back asynchronously, so a step costs
        max(copy, compute) instead of their sum; every byte still crosses PCIe inside the loop."""
        dev = next(self.parameters()).device
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        # two device staging sets kept on the module across calls: a fresh 50 MB torch.empty per call occasionally costs a cudaMalloc
        # (15 ms each, measured), i.e. milliseconds per step of a short stream
        slots = self.__dict__.setdefault("_stream_slots", [None, None])
        slot_free = self.__dict__.setdefault("_stream_slot_free", [None, None])   # kept too: an abandoned generator may leave a step in flight

        def upload(batch, k):
            images, hidden, ids = batch
            mask = self._create_det_token_mask(ids)          # on the host when ids are host-resident: no device sync for the [DET] counts
            with torch.cuda.stream(side):
                if slot_free[k] is not None:
                    side.wait_event(slot_free[k])            # the step that read this slot two batches ago has finished
                cur = slots[k]
                if cur is None or any(d.shape != x.shape or d.dtype != x.dtype for d, x in zip(cur, (images, hidden))):
                    cur = slots[k] = tuple(torch.empty(x.shape, dtype=x.dtype, device=dev) for x in (images, hidden))
                for d, x in zip(cur, (images, hidden)):
                    d.copy_(x, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            return cur, mask, ev

        it = iter(host_batches)
        try:
            nxt = upload(next(it), 0)
        except StopIteration:
            return
        pending = None                                       # (pinned host result, event) of the previous batch
        pool, n_out = [None, None], 0
        while nxt is not None:
            (images, hidden), mask, ev = nxt
            main.wait_event(ev)
            k = n_out & 1
            try:
                nxt = upload(next(it), k ^ 1)
            except StopIteration:
                nxt = None
            _, rec, reps = self.ground_records(images, hidden, mask, copy_out=False)
            if infer:
                xyxy, keep = self._postprocess(rec[:, :4], rec[:, 4] if self.config.use_temp_objectness else None, reps, orig_sizes)
                packed = torch.cat([xyxy, rec[:, 4:5], keep.float()[:, None]], 1)
            else:
                packed = rec
            slot_free[k] = torch.cuda.Event()
            slot_free[k].record(main)
            slot = pool[k]                                   # two pinned result buffers, alternated (cudaHostAlloc per step is slow)
            if slot is None or slot.shape[0] < packed.shape[0] or slot.shape[1] != packed.shape[1]:
                slot = pool[k] = torch.empty((max(packed.shape[0], 1), packed.shape[1]), dtype=packed.dtype, pin_memory=True)
            host = slot[:packed.shape[0]]
            n_out += 1
            host.copy_(packed, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].clone()                     # the pinned slot is reused two batches later
            pending = (host, done)
        pending[1].synchronize()
        yield pending[0].clone()


class _GroundingLossFn(torch.autograd.Function):
    """forward = GroundingBranch.grounding_loss_and_grads (loss and all gradients in one pass over the CUDA library);
    backward only scales the stored gradients by the incoming cotangent."""

    @staticmethod
    def forward(ctx, branch, images, hidden, mask, gt_boxes, gt_obj, *params):
        losses, d_hidden, grads = branch.grounding_loss_and_grads(images, hidden.detach(), mask, gt_boxes, gt_obj, apply=False)
        ctx.d_hidden = d_hidden.to(hidden.dtype)
        ctx.pgrads = [grads.grad_of(p) for p in params]
        ctx.pdtypes = [p.dtype for p in params]
        branch.last_losses = losses
        return (losses["giou_loss"] + losses["l1_loss"] + losses["temp_objectness_loss"]).detach().clone()

    @staticmethod
    def backward(ctx, g):
        gp = [None if t is None else (t * g).to(dt) for t, dt in zip(ctx.pgrads, ctx.pdtypes)]
        return (None, None, ctx.d_hidden * g, None, None, None, *gp)
