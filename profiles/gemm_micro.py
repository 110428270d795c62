"""GEMM micro-benchmark: the encoder's GEMM shapes (ViT-B, 8 frames at 1024^2) through the C ABI, CUDA-event timed.
usage: python profiles/gemm_micro.py [force_ctas ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402

M = 32768
SHAPES = [("qkv  N2304 K768  bf16", 2304, 768, "bf16"), ("proj N768  K768  f32+res", 768, 768, "res"), ("fc1  N3072 K768  gelu", 3072, 768, "gelu"),
          ("fc2  N768  K3072 f32+res", 768, 3072, "res"), ("plain N768 K3072 bf16", 768, 3072, "bf16"), ("plain N3072 K3072 bf16", 3072, 3072, "bf16")]


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def main():
    modes = [int(a) for a in sys.argv[1:]] or [1, 2]
    torch.manual_seed(0)
    for name, N, K, kind in SHAPES:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        for fc in modes:
            if kind == "res":
                out = torch.randn(M, N, device="cuda")
                fn = lambda: ops.gemm(a, w, out, bias=bias, resid=out, force_ctas=fc)
            else:
                out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                fn = lambda: ops.gemm(a, w, out, bias=bias, act="gelu" if kind == "gelu" else None, force_ctas=fc)
            ms = bench(fn)
            print(f"{name:28s} ctas={fc}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
    # the adapter conv: 1 x 8 x 64 x 64 x 768 -> 768, 27 taps
    x = torch.randn(1, 8, 64, 64, 768, device="cuda").to(torch.bfloat16)
    wp = (torch.randn(768, 27 * 768, device="cuda") / (27 * 768) ** 0.5).to(torch.bfloat16)
    out = torch.randn(M, 768, device="cuda")
    alpha = torch.tensor([0.5], device="cuda")
    bias = torch.randn(768, device="cuda")
    for fc in modes:
        ms = bench(lambda: ops.conv_gemm(x, wp, out, V=1, T=8, G=64, kt=3, bias=bias, act="relu", gate_alpha=alpha, resid=out, force_ctas=fc), 5)
        print(f"{'conv3d 27x768 -> 768':28s} ctas={fc}: {ms * 1e3:8.1f} us  {2.0 * M * 768 * 27 * 768 / ms / 1e9:8.1f} TFLOP/s", flush=True)
    # cuBLAS reference points (library GEMM, for context only)
    for N, K in ((2304, 768), (768, 3072), (3072, 3072)):
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        ms = bench(lambda: torch.matmul(a, w.t()))
        print(f"cuBLAS N{N} K{K}: {ms * 1e3:8.1f} us {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
