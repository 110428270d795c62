"""bf16 residual-stream GEMM forms (proj, fc2, adapter conv of ViT-B at 8 x 1024^2) through the C ABI, L2 flushed, CUDA-event timed.
A/B of the two epilogues: GROVE_GEMM_EPI4=1 selects the fp32-staged form, default is the lane = row form (EPI = 8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402
from gemm_micro import bench  # noqa: E402

M = 32768
torch.manual_seed(0)
for name, N, K in (("proj N768 K768", 768, 768), ("fc2 N768 K3072", 768, 3072)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    xs = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    stats = torch.empty(M, N // 128, 2, device="cuda")
    for st in (None, stats):
        ms = bench(lambda: ops.gemm(a, w, xs, bias=bias, resid=xs, ln_stats_out=st))
        print(f"{name:16s} bf16 resid stats={'y' if st is not None else 'n'}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
x = torch.randn(1, 8, 64, 64, 768, device="cuda").to(torch.bfloat16)
wp = (torch.randn(768, 27 * 768, device="cuda") / (27 * 768) ** 0.5).to(torch.bfloat16)
out = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
alpha = torch.tensor([0.5], device="cuda")
bias = torch.randn(768, device="cuda")
ms = bench(lambda: ops.conv_gemm(x, wp, out, V=1, T=8, G=64, kt=3, bias=bias, act="relu", gate_alpha=alpha, resid=out), 5)
print(f"{'conv3d 27x768':16s} bf16 resid: {ms * 1e3:8.1f} us  {2.0 * M * 768 * 27 * 768 / ms / 1e9:8.1f} TFLOP/s", flush=True)
