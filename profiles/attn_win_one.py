"""Run the windowed-attention kernel alone at the bench shape (8 frames, 12 heads, 64x64 grid, 14x14 windows): python profiles/attn_win_one.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402

Fr, G, heads, hd, ws = 8, 64, 12, 64, 14
torch.manual_seed(0)
D = heads * hd
qkv = torch.randn(Fr, G, G, 3, heads, hd, device="cuda").to(torch.bfloat16)
bias = (0.5 * torch.randn(3 * D, device="cuda")).to(torch.bfloat16)
rh = (0.1 * torch.randn(2 * ws - 1, hd, device="cuda")).to(torch.bfloat16)
rw = (0.1 * torch.randn(2 * ws - 1, hd, device="cuda")).to(torch.bfloat16)
tab = ops.window_rel_table(rh, rw)
out = torch.empty(Fr, G, G, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_window_tc(qkv, bias, tab, out, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.attn_window_tc(qkv, bias, tab, out, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
flops = 4.0 * Fr * G * G * 196 * D
print(f"attn_window_tc: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s (useful QK^T+PV)")
