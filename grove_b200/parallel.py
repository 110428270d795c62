"""Sharding of the grounding path over the GPUs of one box.

The path shards into independent units — one 8-frame window of one video (the Conv3d adapter couples the 8 frames of a
window and nothing else, image_encoder.py:52-56) — so there is NO data-path collective.  The only exchange is the optional
all-gather of packed per-frame records [frames, phrases, 5] = (cx, cy, w, h, objectness logit) when the windows of ONE video
are split across ranks (BASELINE config 5); it replaces the reference's pickled `all_gather_object` of per-clip dicts
(infer_iground.py:290-293).  Training (BASELINE config 4) is data parallel over clips: every rank runs the whole training step on its own clip and the
gradients of the trainable grounding parameters are averaged with `allreduce_gradients` — flat fp32 buckets, one NCCL
all-reduce each (the reference delegates this to DeepSpeed ZeRO-2's bucketed reduce, train.py:466-486, reduce_bucket_size 5e8).
Everything here is device-agnostic torch (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def sliding_segment_with_mask(num_frames: int = 48, num_segments: int = 8) -> Tuple[List[List[int]], List[List[int]]]:
    """The reference's long-clip schedule (infer_iground.py:110-148): ceil(F/8) strided sparse windows
    {i*(F//8) + off}, plus a 0/1 mask that drops frames already produced by an earlier window."""
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


def units_of_rank(num_units: int, world: int, rank: int) -> List[int]:
    """unit u -> rank u mod world (videos for configs 3/4, windows of one clip for config 5)."""
    return list(range(rank, num_units, world))


def ground_sharded_clip(window_fn: Callable[..., Optional[torch.Tensor]], num_frames: int, num_phrases: int, *,
                        group: Optional[dist.ProcessGroup] = None, device=None, timings: Optional[dict] = None) -> torch.Tensor:
    """Ground one long clip whose 8-frame windows are split across the ranks of `group` (window w -> rank w mod n).

    `window_fn(frame_ids, out)` runs the per-window hot path (encoder + decoder) on THIS rank; `out` is the window's `[len(frame_ids), P, 5]`
    fp32 slot INSIDE the all-gather send buffer — the heads kernel writes its packed records straight into it (window_fn returns None), or
    window_fn returns a tensor of that shape which is copied in.  Returns, on every rank, the records of all frames in temporal order:
    `[num_frames, P, 5]` fp32.  One all-gather of the pre-packed buffer on the compute stream; no other communication.

    Any clip length works: the reference's schedule (infer_iground.py:110-148) always emits 8-frame windows — for lengths that are not
    a multiple of 8 the remainder windows re-visit frames already produced, and the first-seen masks drop the repeats (:264-266).
    `timings` (optional dict) receives CUDA events around the collective: {"allgather": (start, end)}."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    windows, masks = sliding_segment_with_mask(num_frames, 8)
    wlen = 8
    if any(len(w) != wlen for w in windows):
        raise ValueError("the sliding-window schedule produced a window that is not 8 frames long")
    mine = units_of_rank(len(windows), world, rank)
    per_rank = (len(windows) + world - 1) // world
    buf = None
    for slot, w in enumerate(mine):
        if buf is None:
            dev0 = device if device is not None else _device_of_fn(window_fn)
            buf = torch.zeros(per_rank, wlen, num_phrases, 5, dtype=torch.float32, device=dev0)
        rec = window_fn(windows[w], buf[slot])
        if rec is not None and rec.data_ptr() != buf[slot].data_ptr():
            if tuple(rec.shape) != (wlen, num_phrases, 5):
                raise ValueError(f"window_fn must return [{wlen}, {num_phrases}, 5], got {tuple(rec.shape)}")
            if rec.device != buf.device:         # the first window tells where the path runs
                buf = buf.to(rec.device)
            buf[slot].copy_(rec)
    if buf is None:
        buf = torch.zeros(per_rank, wlen, num_phrases, 5, dtype=torch.float32, device=device or "cpu")
    if world > 1:
        gathered = torch.empty(world * per_rank, wlen, num_phrases, 5, dtype=torch.float32, device=buf.device)
        if timings is not None and buf.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gathered, buf, group=group)
            e1.record()
            timings["allgather"] = (e0, e1)
        else:
            dist.all_gather_into_tensor(gathered, buf, group=group)
    else:
        gathered = buf
    # temporal re-assembly (the reference sorts by first-seen frame index, infer_iground.py:282-288): one gather with a host-built index
    src_rows, dst_rows = [], []
    for w, (idx, msk) in enumerate(zip(windows, masks)):
        base = ((w % world) * per_rank + w // world) * wlen
        for i, mk in enumerate(msk):
            if mk:
                src_rows.append(base + i)
                dst_rows.append(idx[i])
    out = torch.zeros(num_frames, num_phrases, 5, dtype=torch.float32, device=buf.device)
    if src_rows:
        src_t = torch.tensor(src_rows, dtype=torch.long).to(buf.device)
        dst_t = torch.tensor(dst_rows, dtype=torch.long).to(buf.device)
        out[dst_t] = gathered.view(-1, num_phrases, 5)[src_t]
    return out


def _device_of_fn(fn):
    return getattr(fn, "device", None) or "cpu"


def ground_long_clip(branch, clip_images: torch.Tensor, last_hidden_state: torch.Tensor, det_token_mask: torch.Tensor, *,
                     group: Optional[dist.ProcessGroup] = None, timings: Optional[dict] = None) -> torch.Tensor:
    """BASELINE config 5 end to end: one long clip `[1, 3, F, H, W]` with P phrases, its 8-frame windows split across the ranks of `group`,
    every window through `GroundingBranch.ground_records` (encoder + text projection + box decoder + heads), records written into the
    all-gather buffer, one NCCL all-gather, temporal re-assembly.  Returns `[F, P, 5]` fp32 on every rank — the tensor that replaces the
    reference's per-window Python dict merge + pickled `all_gather_object` (infer_iground.py:245-293)."""
    if clip_images.dim() != 5 or clip_images.shape[0] != 1:
        raise ValueError(f"expected one clip [1, 3, F, H, W], got {tuple(clip_images.shape)}")
    num_frames = clip_images.shape[2]
    idx, counts = branch._det_rows(det_token_mask)
    P = counts[0]
    dev = next(branch.parameters()).device
    host_mask = det_token_mask.cpu()         # one read-back for the whole clip; every window reuses the host copy

    def window_fn(frame_ids, out):
        images = clip_images[:, :, frame_ids]
        if images.is_cuda:
            images = images.contiguous()
        branch.ground_records(images, last_hidden_state, host_mask, records_out=out.view(-1, 5), copy_out=False)
        return None
    window_fn.device = dev
    return ground_sharded_clip(window_fn, num_frames, P, group=group, device=dev, timings=timings)


def pack_records(boxes_nested, logits_nested) -> torch.Tensor:
    """nested [V][T] lists of [P,4] boxes / [P] logits (the return value of _generate_and_postprocess_masks with infer=False,
    GROVE.py:297-331) for ONE video -> packed [T, P, 5].  With infer=True the reference drops boxes per frame (ragged): use
    GroundingBranch.ground_records / ground_host_stream, whose records keep every row plus a keep flag."""
    fb, fl = boxes_nested[0], logits_nested[0]
    if any(b.shape[0] != l.shape[0] for b, l in zip(fb, fl)):
        raise ValueError("pack_records needs unfiltered per-frame outputs (infer=False): boxes and logits differ in length")
    return torch.stack([torch.cat([b.float(), l.float()[:, None]], 1) for b, l in zip(fb, fl)])


def allreduce_gradients(grads: Sequence[torch.Tensor], *, group: Optional[dist.ProcessGroup] = None, bucket_elems: int = 1 << 26,
                        average: bool = True) -> None:
    """In-place mean (or sum) over the ranks of `group` of a list of fp32 gradient tensors.  The tensors are packed into flat
    buckets of at most `bucket_elems` elements (256 MB) so that a ViT-H adapter (27 * 1280^2 = 44 M elements) is one collective and
    the ~250 small decoder tensors share one; every rank must pass the tensors in the same order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or not grads:
        return
    world = dist.get_world_size(group)
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        if len(bucket) > 1 or flat.data_ptr() != bucket[0].data_ptr():
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        bucket, size = [], 0

    for g in grads:
        if g.dtype != torch.float32 or not g.is_contiguous():
            raise ValueError("allreduce_gradients expects contiguous fp32 gradient tensors (GradStore accumulators)")
        if size + g.numel() > bucket_elems and bucket:
            flush()
        bucket.append(g)
        size += g.numel()
    flush()


class GradientReducer:
    """Data-parallel gradient averaging OVERLAPPED with the backward pass (BASELINE config 4; the reference delegates this to DeepSpeed
    ZeRO-2's overlap_comm, train.py:466-486).  The training step calls `ready(tensors)` as soon as a group of fp32 gradient accumulators is
    final — the mask decoder right after its backward, each Conv3d adapter right after its weight gradient, text_hidden_fcs at the end —
    and the group is all-reduced on a side stream while the encoder walk continues on the compute stream (a ViT-H adapter is 44 M fp32
    values = 177 MB: its all-reduce hides under the ~8 frozen blocks the walk still has to cross).  `finish()` joins the streams.
    Small tensors of a group share one flat bucket; on CPU tensors (gloo tests) everything runs synchronously."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, bucket_elems: int = 1 << 26, average: bool = True):
        self.group, self.bucket_elems, self.average = group, bucket_elems, average
        self._stream = None
        self.reduced_elems = 0
        self.calls = 0

    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def ready(self, grads: Sequence[torch.Tensor]) -> None:
        grads = [g for g in grads if g is not None]
        if not grads or not self._active():
            return
        self.calls += 1
        self.reduced_elems += sum(g.numel() for g in grads)
        if not grads[0].is_cuda:
            allreduce_gradients(grads, group=self.group, bucket_elems=self.bucket_elems, average=self.average)
            return
        dev = grads[0].device
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))            # the gradients are final at this point of the compute stream
        self._stream.wait_event(ev)
        with torch.cuda.stream(self._stream):
            for g in grads:
                g.record_stream(self._stream)
            allreduce_gradients(grads, group=self.group, bucket_elems=self.bucket_elems, average=self.average)

    def finish(self) -> None:
        """the compute stream waits for every outstanding reduction (call before the optimizer step reads the gradients)"""
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)


def allreduce_gradstore(store, parameters: Sequence[torch.Tensor], **kw) -> None:
    """all-reduce the accumulators of a modeling.decoder_train.GradStore, visiting `parameters` in their (rank-independent) order"""
    allreduce_gradients([store.g[p] for p in parameters if p in store.g], **kw)
