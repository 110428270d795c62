"""Benchmark of the grounding hot path (BASELINE.json: grounding-path frames/s; config[1] = SAM ViT-B encoder + box
decoder, 1 video x 8 frames at 1024^2, 4 phrases, bf16 operands).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One process per GPU (torchrun for N>1); every rank grounds its own clip per step (videos shard with no data-path
collective -> weak scaling).  Prints ONE JSON line on rank 0.  `--impl reference` times the reference's CPU
implementation of the path (the oracle port — the reference is pure PyTorch, oracle/ restates it) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

IMG, FRAMES, PHRASES, SEQ_L, VIT = 1024, 8, 4, 640, "vit_b"
VIDEOS = 1   # videos per GPU per step (config 3: 2)
METRIC, UNIT = "grounding_path_frames_per_s", "frames/s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE representative launch of the dominant kernel: the residual-stream GEMM form that takes
# 38 % of the step (proj, M=32768 N=768 K=768, bf16 residual + row statistics, EPI=8), from the round-2 `ncu --set full` capture summarised in
# profiles/r2_gemm_proj_summary.txt.  Algorithmic bytes of that launch: A 50.3 + residual 50.3 + output 50.3 + W 1.2 + statistics 1.6 = 153.7 MB;
# the 126 MB L2 still holds most of the 50 MB output when the kernel ends, hence the small write figure.
NCU_TRAFFIC_BYTES = 101887488 + 10890752
NCU_TRAFFIC_NOTE = ("proj GEMM launch (M=32768,N=768,K=768, bf16 residual stream): 101.9 MB read + 10.9 MB written vs 153.7 MB algorithmic "
                    "(output still dirty in L2 at kernel end; profiles/r2_gemm_proj_summary.txt)")


def useful_flops_per_frame(D, depth, n_glob, G):
    """multiply-add = 2 FLOPs, real query rows only, all 196 keys per window."""
    N = G * G
    patch = 2 * N * 768 * D
    linear = depth * 2 * N * D * 12 * D
    attn_win = (depth - n_glob) * (4 * N * 196 * D + 4 * N * 14 * D)
    attn_glob = n_glob * (4 * N * N * D + 4 * N * G * D)
    adapters = n_glob * 2 * 27 * D * D * N
    neck = 2 * N * D * 256 + 2 * N * 9 * 256 * 256
    return patch + linear + attn_win + attn_glob + adapters + neck


def decoder_flops_per_instance(N):
    return 670720 * N + 3.55e7


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"], "src": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        top = sm[len(sm) // 2:] if sm else []      # samples under load = the upper half (idle samples precede/follow the region)
        return {"sm_mhz": top[len(top) // 2] if top else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_model(dev):
    from oracle import synth
    from oracle.grounding import VIT_CFG
    from grove_b200.modeling.grounding import GroundingBranch
    cfg = VIT_CFG[VIT]
    gb = GroundingBranch(vit=VIT, num_frames=FRAMES, image_size=IMG)
    sd = synth.synth_state_dict({**synth.encoder_param_shapes(cfg["embed_dim"], cfg["depth"], cfg["heads"], cfg["global_idx"], IMG // 16),
                                 **synth.decoder_param_shapes()}, 21)
    fsd = synth.synth_state_dict(synth.text_fcs_shapes(), 21)
    gb.grounding_encoder.load_state_dict(sd, strict=False)
    gb.text_hidden_fcs.load_state_dict({k[len("text_hidden_fcs."):]: v for k, v in fsd.items()})
    return gb.to(dev), {**sd, **fsd}, cfg


def synth_inputs(n_sets, seed0=100):
    from oracle import synth
    sets = []
    for i in range(n_sets):
        images = synth.synth_tensor(f"bench.images.{i}", (VIDEOS, 3, FRAMES, IMG, IMG), seed0 + i).to(torch.bfloat16)
        hidden = synth.synth_tensor(f"bench.hidden.{i}", (VIDEOS, SEQ_L, 4096), seed0 + i).to(torch.bfloat16)
        ids = torch.full((VIDEOS, SEQ_L - 575), 7, dtype=torch.long)
        for v in range(VIDEOS):
            for p in synth.det_positions(SEQ_L, PHRASES, seed0 + i + 1000 * v):
                ids[v, p - 575 + 1] = 32005
        sets.append((images, hidden, ids))
    return sets


# ------------------------------------------------------------------ CPU reference arm (oracle port of the reference's PyTorch path)
def cpu_reference_step(sd, cfg, inputs, dev="cpu"):
    """One LAYER-SAMPLED pass of the reference algorithm over a full 8-frame 1024^2 clip on the host cores: patch embed,
    ONE windowed block, ONE global block + its Conv3d adapter, the neck, text projection and the full box decoder are
    executed on the real activation shapes; the step time is  t_patch + n_win*t_win + n_glob*(t_glob + t_adapter) +
    t_neck + t_text + t_dec  (the blocks of one kind are identical in cost).  Returns (seconds, parts)."""
    import torch.nn.functional as F
    from oracle import grounding as og
    images, hidden, ids = inputs
    images, hidden = images.float().to(dev), hidden.float().to(dev)
    pre = "image_encoder."
    t = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        t[name] = time.perf_counter() - t0
        return r

    with torch.no_grad():
        def patch():
            x = images.permute(0, 2, 1, 3, 4).reshape(FRAMES, 3, IMG, IMG)
            x = F.conv2d(x, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=16).permute(0, 2, 3, 1)
            return x + sd[pre + "pos_embed"]
        x = timed("patch", patch)
        gi = cfg["global_idx"][0]
        x = timed("win", lambda: og.vit_block(x, sd, f"{pre}blocks.0.", cfg["heads"], 14))
        x = timed("glob", lambda: og.vit_block(x, sd, f"{pre}blocks.{gi}.", cfg["heads"], 0))
        x = timed("adapter", lambda: og.conv_adapter(x, sd, f"{pre}adapters.0."))

        def neck():
            y = F.conv2d(x.permute(0, 3, 1, 2), sd[pre + "neck.0.weight"])
            y = og.layer_norm_2d(y, sd[pre + "neck.1.weight"], sd[pre + "neck.1.bias"])
            y = F.conv2d(y, sd[pre + "neck.2.weight"], padding=1)
            return og.layer_norm_2d(y, sd[pre + "neck.3.weight"], sd[pre + "neck.3.bias"])
        emb = timed("neck", neck)
        mask = og.create_det_token_mask(ids, 32005)
        pred = timed("text", lambda: og.process_hidden_states(hidden, mask, sd, FRAMES))
        reps = [p.shape[0] for p in pred]

        def dec():
            pe = og.dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], emb.shape[-1])
            return og.box_decoder(emb, pe, torch.cat(pred, 0).unsqueeze(1), reps, sd)
        timed("dec", dec)
    n_glob = len(cfg["global_idx"])
    total = t["patch"] + (cfg["depth"] - n_glob) * t["win"] + n_glob * (t["glob"] + t["adapter"]) + t["neck"] + t["text"] + t["dec"]
    return total, t


SAMPLE_DESC = ("layer-sampled 8-frame 1024^2 clip per step: patch embed, 1 of 8 windowed blocks, 1 of 4 global blocks + Conv3d adapter, neck, "
               "text projection, full box decoder (8 frames x 4 phrases) executed in fp32 on the host cores; step time = sum of parts x block counts")
FULL_DESC = ("ONE full, un-sampled pass of the oracle port over the whole workload (8 frames at 1024^2, all 12 blocks, 4 adapters, neck, text "
             "projection, box decoder for 8 x 4 instances), fp32, all host threads")


def cpu_reference_full_pass(sd, cfg, inputs):
    """One complete fp32 pass of the reference algorithm (oracle port) over one clip on the host cores; returns seconds."""
    from oracle import grounding as og
    images, hidden, ids = inputs
    mask = og.create_det_token_mask(ids, 32005)
    t0 = time.perf_counter()
    with torch.no_grad():
        og.grounding_forward(images.float(), hidden.float(), mask, sd, depth=cfg["depth"], heads=cfg["heads"], global_idx=cfg["global_idx"],
                             num_frames=FRAMES)
    return time.perf_counter() - t0


def run_reference(args, rank):
    """The reference arm: the reference's algorithm for the path (oracle port; the reference is pure PyTorch and does not travel) on the host
    cores.  `value` comes from ONE full un-sampled pass; the K timed steps are bounded layer-sampled passes (so the run ends within minutes)
    whose real wall time is `ms_per_step`; their extrapolation to a full clip is a secondary field."""
    if rank != 0:
        return
    from oracle import synth
    from oracle.grounding import VIT_CFG
    cfg = VIT_CFG[VIT]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synth_state_dict({**synth.encoder_param_shapes(cfg["embed_dim"], cfg["depth"], cfg["heads"], cfg["global_idx"], IMG // 16),
                                 **synth.decoder_param_shapes(), **synth.text_fcs_shapes()}, 21)
    inputs = synth_inputs(1)[0]
    for _ in range(args.warmup):
        cpu_reference_step(sd, cfg, inputs)
    wall0 = time.perf_counter()
    tot = 0.0
    for _ in range(args.steps):
        s, _ = cpu_reference_step(sd, cfg, inputs)
        tot += s
    wall = time.perf_counter() - wall0
    full_s = cpu_reference_full_pass(sd, cfg, inputs)
    v = FRAMES / full_s
    _emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": workload_config(args.gpus),
           "step_definition": "each of the K timed steps is a bounded layer-sampled pass (ms_per_step is its real wall time); `value` is measured on "
                              "one full un-sampled pass run after them",
           "full_pass_ms": 1e3 * full_s,
           "sampled_extrapolation": {"value": FRAMES * args.steps / tot if tot > 0 else None, "unit": UNIT, "ms_per_full_step_extrapolated": 1e3 * tot / max(args.steps, 1),
                                     "how": SAMPLE_DESC},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": FULL_DESC},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def workload_config(n):
    name = {"vit_b": "ViT-B", "vit_l": "ViT-L", "vit_h": "ViT-H"}[VIT]
    cfgname = "BASELINE configs[1]" if (VIT, VIDEOS) == ("vit_b", 1) else ("BASELINE configs[2], one GPU's share" if (VIT, VIDEOS) == ("vit_h", 2) else "non-default")
    return {"workload": f"SAM {name} encoder (+4 Conv3d adapters) + text projection + box decoder + heads; {VIDEOS} video(s) x {FRAMES} frames at {IMG}^2, "
                        f"{PHRASES} phrases per GPU ({cfgname})", "frames_per_step_per_gpu": FRAMES * VIDEOS, "phrases": PHRASES,
            "image_size": IMG, "sharding": f"by video, {n} GPU(s), no data-path collective",
            "l2": "4 rotating input sets (220 MB) and a ~3 GB per-step activation working set, both larger than the 126 MB L2",
            "launch": "the whole step (im2col, encoder, [DET] gather + text projection, box decoder, heads) replayed as ONE CUDA graph "
                      "(GroundingBranch.enable_cuda_graphs); [DET] bookkeeping on the host before the launch, no host sync inside the step"}


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries write to fd 1 (NCCL prints its version banner there) to stderr: stdout carries the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="grove_b200", choices=["grove_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch kernel by kernel instead of replaying the whole-step CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (config5: long clip + all-gather; config3: ViT-H share)")
    ap.add_argument("--vit", default="vit_b", choices=["vit_b", "vit_l", "vit_h"], help="non-default workloads are for profiling only")
    ap.add_argument("--videos", type=int, default=1, help="videos per GPU per step (BASELINE configs[2] = vit_h with 2)")
    args = ap.parse_args()
    global VIT, VIDEOS
    VIT, VIDEOS = args.vit, args.videos
    if (VIT, VIDEOS) != ("vit_b", 1):
        args.no_cpu_baseline = True
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints a banner there)
        dist.init_process_group("nccl", device_id=dev)
    from grove_b200 import ops, parallel
    gb, sd, cfg = build_model(dev)
    gb.enable_cuda_graphs(not args.no_graphs)   # serving mode (frozen weights): the whole step is one graph replay
    host_sets = [tuple(t.pin_memory() for t in s) for s in synth_inputs(4, 100 + 10 * rank)]
    dev_sets = [tuple(t.to(dev) for t in s[:2]) for s in host_sets]

    def step_resident(i):
        # images and hidden states resident in HBM; input_ids are host-resident (they come from the tokenizer), so the [DET] mask is built on
        # the host and nothing inside the step waits for the device
        images, hidden = dev_sets[i % len(dev_sets)]
        mask = gb._create_det_token_mask(host_sets[i % len(host_sets)][2])
        return gb.ground(images, hidden, mask, infer=False)

    def step_e2e(i):
        images, hidden, ids = host_sets[i % len(host_sets)]                  # pinned host memory
        mask = gb._create_det_token_mask(ids)
        _, rec, _ = gb.ground_records(images, hidden, mask, copy_out=False)  # uploads, grounds; records [B,5] = boxes + objectness logit
        return rec.cpu()                                                    # device -> host read of the step's result (blocking)

    def loop_e2e(first, n):
        # the serving loop of the public API: pinned host batches in, packed host results out, uploads / read-backs overlapped
        outs = 0
        for r in gb.ground_host_stream(host_sets[(first + i) % len(host_sets)] for i in range(n)):
            outs += r.shape[0]
        assert outs == n * FRAMES * VIDEOS * PHRASES

    def timed_loop(loop, steps, warmup):
        loop(0, warmup)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(warmup, steps)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        launches = ops.launch_count()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e_sync, _ = timed(step_e2e, args.steps, max(args.warmup, 3))       # one blocking call per step (no overlap)
    ms_e2e = timed_loop(loop_e2e, args.steps, max(args.warmup, 3))           # GroundingBranch.ground_host_stream

    # ---- secondary legs: BASELINE configs[4] (long clip split by window + packed NCCL all-gather) and configs[2] (ViT-H, 2 videos per GPU)
    extras = {}
    if not args.no_extras and (VIT, VIDEOS) == ("vit_b", 1):
        try:
            extras["config5"] = bench_config5(gb, dev_sets, dev, world, dist, parallel, ops)
        except Exception as e:   # the secondary legs must never take the headline down
            extras["config5"] = {"error": repr(e)[:200]}
    if not args.no_extras and (VIT, VIDEOS) == ("vit_b", 1):
        try:
            extras["config3"], extras["config4"] = bench_config3_and_4(dev, world, dist, ops, parallel)
        except Exception as e:   # the secondary legs must never take the headline down
            extras["config3"] = {"error": repr(e)[:200]}

    # ---- per-kernel device time of the tensor-core GEMM (the dominant kernel), CUDA events on the launching stream
    rec = []
    orig_gemm, orig_conv = ops.gemm, ops.conv_gemm

    def wrap(fn, flops_of):
        def w(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            rec.append((s, e, flops_of(*a, **k)))
            return r
        return w
    g_w = wrap(orig_gemm, lambda a, w, out, **k: 2.0 * a.shape[0] * a.shape[1] * w.shape[0])
    c_w = wrap(orig_conv, lambda x, wp, out, **k: 2.0 * out.shape[0] * wp.shape[0] * wp.shape[1])
    ops.gemm, ops.conv_gemm = g_w, c_w
    gb.enable_cuda_graphs(False)                                     # per-launch events need kernel-by-kernel launches
    gb.grounding_encoder.image_encoder.enable_cuda_graphs(False)
    for i in range(2):
        rec.clear()
        torch.cuda._sleep(int(4e7))      # ~20 ms of spin: the host queues the whole step behind it, so the events bracket GPU time only
        step_resident(i)
    torch.cuda.synchronize()
    ops.gemm, ops.conv_gemm = orig_gemm, orig_conv
    gemm_ms = sum(s.elapsed_time(e) for s, e, _ in rec)
    gemm_flops = sum(f for _, _, f in rec)
    pk = peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_ms = ms / args.steps
    G = IMG // 16
    TOT = FRAMES * VIDEOS
    flops_step = TOT * useful_flops_per_frame(cfg["embed_dim"], cfg["depth"], len(cfg["global_idx"]), G) + \
        TOT * PHRASES * decoder_flops_per_instance(G * G) + VIDEOS * PHRASES * 35.7e6
    step_tflops = flops_step / (step_ms * 1e-3) / 1e12

    if rank == 0:
        value = world * TOT * args.steps / (ms * 1e-3)
        e2e_v = world * TOT * args.steps / (ms_e2e_sync * 1e-3)
        # what the blocking call uploads: the images, the [DET] rows of the hidden states (gathered on the host: only they are read) and
        # their row indices
        h2d = host_sets[0][0].numel() * host_sets[0][0].element_size() + PHRASES * VIDEOS * host_sets[0][1].shape[-1] * host_sets[0][1].element_size() \
            + 4 * PHRASES * VIDEOS
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": TOT * PHRASES * 5 * 4,
                        "api": "one GroundingBranch.ground_records() call per step on pinned host inputs: upload (images, the [DET] rows of the hidden states "
                               "gathered on the host, their row indices), ground, read the packed boxes + objectness back (blocking)",
                        "pipelined_value": world * TOT * args.steps / (ms_e2e * 1e-3),
                        "pipelined_api": "GroundingBranch.ground_host_stream (next upload / previous read-back overlap the current step)"},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": achieved / pk["bf16_tflops_sustained"], "traffic": NCU_TRAFFIC_BYTES,
                             "traffic_note": NCU_TRAFFIC_NOTE,
                             "kernel": "gemm_bf16_tcgen05_kernel (all GEMM + implicit-conv launches of one step: %d launches, %.3f ms, %.2f TFLOP)"
                                       % (len(rec), gemm_ms, gemm_flops / 1e12),
                             "peak_source": pk["src"] + " cuBLAS bf16 sustained (kernel timed inside a long step)",
                             "step_useful_tflops": step_tflops, "step_frac_of_peak": step_tflops / pk["bf16_tflops_sustained"],
                             "step_frac_of_burst_peak": step_tflops / pk["bf16_tflops"], "step_frac_of_nominal_2250": step_tflops / 2250.0,
                             "gemm_share_of_step": gemm_ms / step_ms}}
        line.update(extras)
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            cpu_sd = {k: v.float() for k, v in sd.items()}
            sec = cpu_reference_full_pass(cpu_sd, cfg, tuple(t.clone() for t in host_sets[0]))
            line["cpu_baseline"] = {"value": FRAMES / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": FULL_DESC}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _max_over_ranks(x, dev, world, dist):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_config5(gb, dev_sets, dev, world, dist, parallel, ops, frames=128, phrases=16, reps=3):
    """BASELINE configs[4]: ONE long clip of 128 frames x 16 phrases (same ViT-B, 1024^2), its sixteen 8-frame windows split over the ranks
    (window w -> rank w mod N), records written into the all-gather buffer, one NCCL all-gather of packed [frames/N * 16, 5] fp32, temporal
    re-assembly.  Strong scaling (total work fixed).  Times the whole clip and the collective alone with CUDA events (max over ranks)."""
    from oracle import synth
    clip = torch.cat([d[0] for d in dev_sets], 2)                    # 32 distinct synthetic frames ...
    clip = torch.cat([clip] * (frames // clip.shape[2]), 2)          # ... tiled to 128 (every rank builds the same clip)
    hidden = dev_sets[0][1]
    ids = torch.full((1, SEQ_L - 575), 7, dtype=torch.long)
    for p in synth.det_positions(SEQ_L, phrases, 77):
        ids[0, p - 575 + 1] = 32005
    mask = gb._create_det_token_mask(ids)
    parallel.ground_long_clip(gb, clip, hidden, mask)               # warm-up (captures the 16-phrase graph, creates the communicator's buffers)
    clip_ms, ag_us = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tm = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = parallel.ground_long_clip(gb, clip, hidden, mask, timings=tm)
        e1.record()
        torch.cuda.synchronize()
        clip_ms.append(_max_over_ranks(e0.elapsed_time(e1), dev, world, dist))
        if "allgather" in tm:
            ag_us.append(_max_over_ranks(1e3 * tm["allgather"][0].elapsed_time(tm["allgather"][1]), dev, world, dist))
    best = min(clip_ms)
    return {"workload": f"1 clip x {frames} frames x {phrases} phrases, ViT-B at {IMG}^2, {frames // 8} windows split over {world} GPU(s)",
            "clip_ms": best, "frames_per_s": frames / (best * 1e-3), "allgather_us": (min(ag_us) if ag_us else None),
            "allgather_us_note": "CUDA events around the collective on the compute stream: includes waiting for the slowest rank's last window "
                                 "(the transfer itself is ~13 us for 20 KB per rank, profiles/r2_nccl_2gpu.json)",
            "allgather_bytes_per_rank": (frames // 8 + world - 1) // world * 8 * phrases * 5 * 4, "scaling": "strong",
            "records_shape": list(out.shape), "collective": "dist.all_gather_into_tensor (NCCL) on the compute stream" if world > 1 else None}


def bench_config3_and_4(dev, world, dist, ops, parallel, videos=2, steps=5, warmup=2):
    """BASELINE configs[2], one GPU's share: SAM ViT-H (the model GROVE builds, GROVE.py:55) + box decoder on 2 videos x 8 frames at 1024^2 with
    4 phrases each; random-init weights of that architecture (torch default init on the device, zero-initialised tables re-randomised).
    Then BASELINE configs[3] on the same model: the training step of the grounding branch (forward + backward + L1/GIoU + objectness loss)
    on one 32-frame clip per GPU at the reference's training resolution (512^2 after interpolate_positional_embeddings, train.py:52,561-576),
    gradients of the trainable grounding parameters averaged over the ranks with the all-reduce overlapped with the backward pass."""
    from grove_b200.modeling.grounding import GroundingBranch
    from oracle import synth
    from oracle.grounding import VIT_CFG
    cfg = VIT_CFG["vit_h"]
    torch.manual_seed(1234)
    gh = GroundingBranch(vit="vit_h", num_frames=FRAMES, image_size=IMG).to(dev)
    enc = gh.grounding_encoder.image_encoder
    with torch.no_grad():
        enc.pos_embed.normal_(std=0.02)
        for blk in enc.blocks:
            blk.attn.rel_pos_h.normal_(std=0.02); blk.attn.rel_pos_w.normal_(std=0.02)
        for a in enc.adapters:
            a.alpha.fill_(0.5)
    gh.enable_cuda_graphs(True)
    g = torch.Generator(device=dev).manual_seed(5 + int(os.environ.get("RANK", 0)))
    sets = []
    for i in range(2):
        images = torch.randn(videos, 3, FRAMES, IMG, IMG, device=dev, generator=g).to(torch.bfloat16)
        hidden = torch.randn(videos, SEQ_L, 4096, device=dev, generator=g).to(torch.bfloat16)
        ids = torch.full((videos, SEQ_L - 575), 7, dtype=torch.long)
        for v in range(videos):
            for p in synth.det_positions(SEQ_L, PHRASES, 300 + i + 10 * v):
                ids[v, p - 575 + 1] = 32005
        sets.append((images, hidden, gh._create_det_token_mask(ids)))
    for i in range(warmup):
        gh.ground(*sets[i % 2], infer=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        gh.ground(*sets[i % 2], infer=False)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world, dist) / steps
    fl = videos * FRAMES * (useful_flops_per_frame(cfg["embed_dim"], cfg["depth"], len(cfg["global_idx"]), IMG // 16)
                            + PHRASES * decoder_flops_per_instance((IMG // 16) ** 2))
    pk = peaks()
    c3 = {"workload": f"SAM ViT-H + box decoder, {videos} videos x {FRAMES} frames at {IMG}^2 x {PHRASES} phrases per GPU (BASELINE configs[2] share)",
          "ms_per_step": ms, "frames_per_s": world * videos * FRAMES / (ms * 1e-3), "steps": steps,
          "step_useful_tflops": fl / (ms * 1e-3) / 1e12, "step_frac_of_peak": fl / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
          "scaling": "weak", "weights": "torch default random init (not the deterministic synth set: 818 M parameters)"}
    try:
        c4 = _bench_config4(gh, dev, world, dist, parallel, g)
    except Exception as e:
        c4 = {"error": repr(e)[:200]}
    del gh
    torch.cuda.empty_cache()
    return c3, c4


def _bench_config4(gh, dev, world, dist, parallel, g, frames=32, img=512, steps=3, warmup=1):
    from grove_b200 import checkpoint as ck
    from oracle import synth
    gh.enable_cuda_graphs(False)
    gh._graphs = {}
    torch.cuda.empty_cache()
    enc = gh.grounding_encoder.image_encoder
    ck.interpolate_positional_embeddings(enc, img)                 # train.py:561-576
    gh.grounding_encoder.prompt_encoder.image_embedding_size = (img // 16, img // 16)     # the reference's factory value (build_sam.py:66-69)
    gh.config.num_frames = frames
    for p in gh.parameters():                                      # GROVE's freeze pattern (train.py:254-296)
        p.requires_grad_(False)
    for p in list(enc.adapters.parameters()) + list(gh.grounding_encoder.mask_decoder.parameters()) + list(gh.text_hidden_fcs.parameters()):
        p.requires_grad_(True)
    images = torch.randn(1, 3, frames, img, img, device=dev, generator=g).to(torch.bfloat16)
    hidden = torch.randn(1, SEQ_L, 4096, device=dev, generator=g).to(torch.bfloat16)
    ids = torch.full((1, SEQ_L - 575), 7, dtype=torch.long)
    for p in synth.det_positions(SEQ_L, PHRASES, 400):
        ids[0, p - 575 + 1] = 32005
    mask = gh._create_det_token_mask(ids).to(dev)
    cg = torch.Generator().manual_seed(9)
    gt_b, gt_o = [[]], [[]]
    for f in range(frames):
        lab = (torch.rand(PHRASES, generator=cg) > 0.5).double()
        if f == 0:
            lab[0] = 1.0
        nb = int(lab.sum())
        gt_b[0].append(torch.cat([torch.rand(nb, 2, generator=cg) * 0.4 + 0.3, torch.rand(nb, 2, generator=cg) * 0.3 + 0.1], 1))
        gt_o[0].append(lab)
    red = None

    def step():
        nonlocal red
        red = parallel.GradientReducer() if world > 1 else None
        return gh.grounding_loss_and_grads(images, hidden, mask, gt_b, gt_o, apply=False, reducer=red)
    for _ in range(warmup):
        losses, _, _ = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    e0.record()
    evs[0].record()
    for i in range(steps):
        losses, _, grads = step()
        evs[i + 1].record()
    e1.record()
    torch.cuda.synchronize()
    if os.environ.get("GROVE_BENCH_DEBUG"):
        print("config4 per-step ms:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(steps)], file=sys.stderr)
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world, dist) / steps
    trainable = sum(t.numel() for t in grads.g.values())
    return {"workload": f"training step of the grounding branch (forward + backward, GIoU + L1 + objectness), SAM ViT-H, 1 clip x {frames} frames at "
                        f"{img}^2 x {PHRASES} phrases per GPU (BASELINE configs[3]: data parallel over clips)",
            "ms_per_step": ms, "frames_per_s": world * frames / (ms * 1e-3), "steps": steps, "loss": float(losses["loss"]),
            "trainable_params": trainable, "scaling": "weak",
            "gradient_allreduce": (f"{red.reduced_elems * 4 / 1e6:.0f} MB fp32 in {red.calls} groups on a side stream, overlapped with the backward pass "
                                   "(parallel.GradientReducer, NCCL)") if red is not None else None,
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}


if __name__ == "__main__":
    main()
