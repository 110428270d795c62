"""Worker of tests/test_gpu_nccl.py, launched under torchrun with one rank per GPU (NCCL over NVLink):

  (1) BASELINE config 5 — a 128-frame x 16-phrase clip (ViT-B, 512^2), its sixteen 8-frame windows split over the ranks, the packed
      [frames_local * P, 5] records all-gathered once (grove_b200.parallel.ground_long_clip); every rank must hold the same [128, 16, 5]
      tensor, and rank 0 checks it against oracle.grounding_forward per window (boxes 1e-2, logits 2e-2);
  (2) BASELINE config 4's data-parallel step — every rank runs the training step on its own clip, the fp32 gradient accumulators are
      averaged with the bucketed NCCL all-reduce (parallel.allreduce_gradstore) and compared with the mean of the per-rank gradients.

Prints one line `NCCL_WORKER_OK {json}` on rank 0.  Not collected by pytest (no test_ prefix)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dist.init_process_group("nccl", device_id=dev)
    from grove_b200 import parallel
    from oracle import synth
    from test_gpu_model import _branch_with_weights, _long_clip_case, _oracle_long_clip
    report = {"world": world}
    # ---------------- (1) config 5
    Ftot, P, img, seed = 128, 16, 512, 23
    gb, cfg, full = _branch_with_weights("vit_b", img, seed)
    clip, hid, mask = _long_clip_case(gb, cfg, full, Ftot, P, img, seed)
    timings = {}
    out = parallel.ground_long_clip(gb, clip, hid, mask, timings=timings)
    torch.cuda.synchronize()
    e0, e1 = timings["allgather"]
    report["allgather_us_first"] = 1e3 * e0.elapsed_time(e1)
    timings = {}
    out2 = parallel.ground_long_clip(gb, clip, hid, mask, timings=timings)
    torch.cuda.synchronize()
    e0, e1 = timings["allgather"]
    report["allgather_us"] = 1e3 * e0.elapsed_time(e1)
    report["allgather_bytes_per_rank"] = (Ftot // world) * P * 5 * 4
    assert torch.equal(out, out2)
    ref0 = out.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(out, ref0), f"rank {rank} holds different records than rank 0"
    if rank == 0:
        ref = _oracle_long_clip(cfg, full, clip, hid, mask, Ftot, P)
        eb, el = float((out[..., :4] - ref[..., :4]).abs().max()), float((out[..., 4] - ref[..., 4]).abs().max())
        report["config5_max_dbox"], report["config5_max_dlogit"] = eb, el
        assert eb < 1e-2 and el < 2e-2, (eb, el)
        safe = ref[..., 4].abs() > 3 * max(el, 1e-4)
        assert torch.equal((out[..., 4] > 0)[safe], (ref[..., 4] > 0)[safe])
    del clip, out, out2, ref0
    # ---------------- (2) config 4: data-parallel training step, gradient all-reduce on NCCL
    from test_gpu_training import _setup
    gbt, cfg, full, images, hidden, mask, gt_boxes, gt_obj = _setup("vit_b", 512, 1, 8, 2, 40)      # same weights on every rank (seed)
    images = synth.synth_tensor(f"nccl.images.{rank}", tuple(images.shape), 40 + rank).cuda()        # a different clip per rank
    _, _, grads = gbt.grounding_loss_and_grads(images.to(torch.bfloat16), hidden.to(torch.bfloat16), mask, gt_boxes, gt_obj, apply=False)
    params = [p for p in gbt.parameters() if p in grads.g]
    probe = [gbt.grounding_encoder.image_encoder.adapters[0].conv3d.weight, gbt.grounding_encoder.mask_decoder.bbox_prediction_head[2].weight,
             gbt.text_hidden_fcs[0][2].bias]
    local_g = [grads.g[p].clone() for p in probe]
    gathered = [[torch.empty_like(g) for _ in range(world)] for g in local_g]
    for g, lst in zip(local_g, gathered):
        dist.all_gather(lst, g)
    parallel.allreduce_gradstore(grads, params)
    torch.cuda.synchronize()
    errs = []
    for p, lst in zip(probe, gathered):
        mean = torch.stack(lst).mean(0)
        errs.append(float((grads.g[p] - mean).abs().max() / (mean.abs().max() + 1e-30)))
        assert float((lst[0] - lst[-1]).abs().max()) > 0          # the ranks really had different gradients
    report["allreduce_rel_err"] = max(errs)
    assert max(errs) < 1e-5, errs
    report["grad_tensors"] = len(params)
    # the same step with the reduction OVERLAPPED with the backward pass (parallel.GradientReducer: decoder, then each adapter as the encoder
    # walk finishes it, then text_hidden_fcs, all-reduced on a side stream): equals the mean of the per-rank gradients up to the run-to-run
    # noise of the step itself (a few float atomics, ~1e-4 of a gradient's norm)
    red = parallel.GradientReducer()
    _, _, grads2 = gbt.grounding_loss_and_grads(images.to(torch.bfloat16), hidden.to(torch.bfloat16), mask, gt_boxes, gt_obj, apply=False, reducer=red)
    torch.cuda.synchronize()
    errs2 = []
    for p, lst in zip(probe, gathered):
        mean = torch.stack(lst).mean(0)
        errs2.append(float((grads2.g[p] - mean).norm() / (mean.norm() + 1e-30)))
    report["overlapped_rel_err"], report["overlapped_calls"], report["overlapped_elems"] = max(errs2), red.calls, red.reduced_elems
    assert max(errs2) < 2e-3 and red.calls >= 6, (errs2, red.calls)
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK " + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
