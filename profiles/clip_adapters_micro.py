"""SURVEY.md 8f-3 measurement: the CLIP-side adapter (Conv3d 1024->1024, 8 frames x 16x16 tokens per video) and AdaptiveAvgPooling3D at
production shapes, CUDA events.  python profiles/clip_adapters_micro.py [videos]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200.clip_adapters import AdaptiveAvgPooling3D, SpatioTemporalConvAdapter  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 2
D = 1024
torch.manual_seed(0)
ad = SpatioTemporalConvAdapter(D, D, (3, 3, 3)).cuda().to(torch.bfloat16)
with torch.no_grad():
    ad.alpha.fill_(0.5)
x = torch.randn(V * 8, 257, D, device="cuda").to(torch.bfloat16)
pool = AdaptiveAvgPooling3D(num_frames=8)
feat = torch.randn(V * 8, 576, D, device="cuda").to(torch.bfloat16)


def bench(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


ms = bench(lambda: ad((x,)))
flops = 2.0 * V * 8 * 256 * D * 27 * D
print(f"CLIP SpatioTemporalConvAdapter, {V} videos x 8 frames x 16x16 x {D}: {ms * 1e3:.1f} us per call (module: casts + cat + implicit GEMM), "
      f"{flops / ms / 1e9:.1f} TFLOP/s")
ms = bench(lambda: pool(feat))
byts = feat.numel() * 2 + V * 576 * D * 2
print(f"AdaptiveAvgPooling3D, {V} videos x 8 frames x 24x24 x {D} bf16 -> {V} x 576 x {D}: {ms * 1e3:.1f} us, {byts / ms / 1e6:.0f} GB/s "
      f"(algorithmic bytes {byts / 1e6:.1f} MB)")
