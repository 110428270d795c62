"""Time one training step of the grounding branch (BASELINE config 4: forward + backward + losses) on one GPU.
    python profiles/train_step.py [vit] [frames] [img] [phrases] [--profile]
Prints a JSON line: forward-only ms (inference path), training-step ms, peak memory."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
bench.VIT = args[0] if len(args) > 0 else "vit_b"
bench.FRAMES = int(args[1]) if len(args) > 1 else 8
bench.IMG = int(args[2]) if len(args) > 2 else 1024
bench.PHRASES = int(args[3]) if len(args) > 3 else 4
profile = "--profile" in sys.argv
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
gb, sd, cfg = bench.build_model(dev)
for p in gb.parameters():
    p.requires_grad_(False)
enc = gb.grounding_encoder.image_encoder
for p in list(enc.adapters.parameters()) + list(gb.grounding_encoder.mask_decoder.parameters()) + list(gb.text_hidden_fcs.parameters()):
    p.requires_grad_(True)
images, hidden, ids = (t.to(dev) for t in bench.synth_inputs(1)[0])
mask = gb._create_det_token_mask(ids)
T, P = bench.FRAMES, bench.PHRASES
g = torch.Generator().manual_seed(0)
gt_b, gt_o = [[]], [[]]
for f in range(T):
    lab = (torch.rand(P, generator=g) > 0.5).double()
    lab[0] = 1.0
    nb = int(lab.sum())
    gt_b[0].append(torch.cat([torch.rand(nb, 2, generator=g) * 0.4 + 0.3, torch.rand(nb, 2, generator=g) * 0.3 + 0.1], 1))
    gt_o[0].append(lab)


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd = lambda: gb.ground(images, hidden, mask, infer=False)
step = lambda: gb.grounding_loss_and_grads(images, hidden, mask, gt_b, gt_o, apply=False)
for _ in range(2):
    fwd(); step()
if profile:
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
torch.cuda.reset_peak_memory_stats()
ms_f = timed(fwd, 5)
ms_t = timed(step, 5)
print(json.dumps({"workload": f"{bench.VIT} {T} frames @ {bench.IMG}^2, {P} phrases, 1 GPU", "forward_ms": ms_f, "train_step_ms": ms_t,
                  "bwd_over_fwd": (ms_t - ms_f) / ms_f, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "frames_per_s_train": T / (ms_t * 1e-3)}))
