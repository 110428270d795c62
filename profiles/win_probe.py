"""Timeline of one CTA of the window-attention kernel (build with GROVE_NVCC_EXTRA=-DGROVE_WIN_PROBE python -m grove_b200.build -f):
clock64 at every hand-off of the issuing warp and of three softmax warps, last 8 units of the CTA."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from grove_b200 import ops
from grove_b200._lib import lib
Fr, G, heads, hd, ws = 8, 64, 12, 64, 14
torch.manual_seed(0)
D = heads * hd
qkv = torch.randn(Fr, G, G, 3, heads, hd, device="cuda").to(torch.bfloat16)
bias = (0.5 * torch.randn(3 * D, device="cuda")).to(torch.bfloat16)
rh = (0.1 * torch.randn(27, hd, device="cuda")).to(torch.bfloat16)
rw = (0.1 * torch.randn(27, hd, device="cuda")).to(torch.bfloat16)
tab = ops.window_rel_table(rh, rw)
out = torch.empty(Fr, G, G, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_window_tc(qkv, bias, tab, out, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
l = lib()
l.grove_win_probe_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert l.grove_win_probe_read(buf, 4096) == 0
a = np.array(buf[:], dtype=np.int64)
t0 = a[100 + 16 * 0]
r = lambda x: int(x - t0)
# the probe ring keeps the LAST 8 units of the CTA (cnt & 7); print in slot order
print("MMA per unit slot: [pre QK wait, post, post O_READ, T0S0 issued, T_READ0, T1S1 issued, T_READ1, V ready, P_FULL0, O_READ0', PV0 issued, P_FULL1, O_READ, PV1 issued]")
for c in range(8):
    print("  ", c, [r(a[100 + 16 * c + i]) for i in range(14)])
for w in (0, 1, 4):
    print(f"softmax warp {w}: per unit slot, tile: [pre T wait, post, pre S wait, post, P arrive] ... [pre O wait, post] x2, end")
    for c in range(8):
        base = 1000 + w * 256 + 32 * c
        print("  ", c, [r(a[base + i]) for i in range(5)], [r(a[base + 8 + i]) for i in range(5)], [r(a[base + 5]), r(a[base + 6]), r(a[base + 13]), r(a[base + 14]), r(a[base + 23])])
