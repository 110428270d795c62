"""Run one encoder GEMM shape a few times (for `ncu --set full`): python profiles/gemm_one.py N K kind ctas"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402

N, K, kind, fc = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
M = 32768
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
for _ in range(4):
    if kind == "res16":                                  # bf16 residual stream + row statistics (EPI = 8; GROVE_GEMM_EPI4=1: EPI = 4)
        xs = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        stats = torch.empty(M, N // 128, 2, device="cuda")
        ops.gemm(a, w, xs, bias=bias, resid=xs, ln_stats_out=stats, force_ctas=fc)
    elif kind == "res":
        out = torch.randn(M, N, device="cuda")
        ops.gemm(a, w, out, bias=bias, resid=out, force_ctas=fc)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, out, bias=bias, act="gelu" if kind == "gelu" else None, force_ctas=fc)
torch.cuda.synchronize()
