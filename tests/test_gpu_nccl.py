"""The multi-GPU path on real NCCL (2 B200s of one box): config 5's window-sharded long clip with the packed all-gather, and
config 4's gradient all-reduce.  Skipped on a one-GPU box; tests/test_parallel_gloo.py covers the host logic on CPU."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under `gpurun --gpus 2`)")
def test_config5_allgather_and_gradient_allreduce_on_nccl():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(here, "nccl_worker.py")], capture_output=True, text=True, timeout=1500, env=env)
    ok = [l for l in r.stdout.splitlines() if l.startswith("NCCL_WORKER_OK ")]
    assert r.returncode == 0 and ok, (r.stdout[-3000:], r.stderr[-3000:])
    rep = json.loads(ok[0][len("NCCL_WORKER_OK "):])
    print(rep)
    out = os.path.join(os.path.dirname(here), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(rep, open(os.path.join(out, "r2_nccl_2gpu.json"), "w"), indent=1)
    assert rep["world"] == 2 and rep["config5_max_dbox"] < 1e-2 and rep["config5_max_dlogit"] < 2e-2 and rep["allreduce_rel_err"] < 1e-5
