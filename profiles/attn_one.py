"""Run the global-attention kernel alone at the bench shape (8 frames, 12 heads, 64x64 grid): python profiles/attn_one.py [legacy]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402

legacy = len(sys.argv) > 1 and sys.argv[1] == "legacy"
Fr, G, heads, hd = 8, 64, 12, 64
torch.manual_seed(0)
qkv = torch.randn(Fr, G, G, 3, heads, hd, device="cuda").to(torch.bfloat16)
rh = (0.1 * torch.randn(2 * G - 1, hd, device="cuda")).to(torch.bfloat16)
rw = (0.1 * torch.randn(2 * G - 1, hd, device="cuda")).to(torch.bfloat16)
out = torch.empty(Fr, G, G, heads * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_global(qkv, rh, rw, out, F=Fr, G=G, heads=heads, hd=hd, legacy_mma=legacy)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.attn_global(qkv, rh, rw, out, F=Fr, G=G, heads=heads, hd=hd, legacy_mma=legacy)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
flops = 4.0 * Fr * heads * (G * G) ** 2 * hd
print(f"attn_global legacy={legacy}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s (QK^T+PV)")
