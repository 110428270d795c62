"""One steady-state step of bench.py's workload between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if len(sys.argv) > 1:
    bench.VIT = sys.argv[1]
if len(sys.argv) > 2:
    bench.VIDEOS = int(sys.argv[2])
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
gb, sd, cfg = bench.build_model(dev)
images, hidden, ids = (t.to(dev) for t in bench.synth_inputs(1)[0])
mask = gb._create_det_token_mask(ids)
for _ in range(2):
    gb.ground(images, hidden, mask, infer=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
gb.ground(images, hidden, mask, infer=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
