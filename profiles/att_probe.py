"""Timeline of one CTA of the global-attention kernel (build with GROVE_NVCC_EXTRA=-DGROVE_ATT_PROBE python -m grove_b200.build -f):
clock64 at every hand-off of the issuing warp and of three softmax warps, per 128-key block."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from grove_b200 import ops
from grove_b200._lib import lib
Fr, G, heads, hd = 8, 64, 12, 64
torch.manual_seed(0)
qkv = torch.randn(Fr, G, G, 3, heads, hd, device="cuda").to(torch.bfloat16)
rh = (0.1 * torch.randn(2 * G - 1, hd, device="cuda")).to(torch.bfloat16)
rw = (0.1 * torch.randn(2 * G - 1, hd, device="cuda")).to(torch.bfloat16)
out = torch.empty(Fr, G, G, heads * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_global(qkv, rh, rw, out, F=Fr, G=G, heads=heads, hd=hd)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8192)()
l = lib()
l.grove_att_probe_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert l.grove_att_probe_read(buf, 8192) == 0
a = np.array(buf[:], dtype=np.int64)
t0 = a[1000]
r = lambda x: int(x - t0)
print("MMA qk [pre-wait, post-wait, issued] per kit:")
for k in range(32):
    print("  ", k, [r(a[1000 + 3 * k + i]) for i in range(3)])
print("MMA pv [pre-wait, post-wait, issued]:")
for b in range(32):
    print("  ", b, [r(a[1300 + 3 * b + i]) for i in range(3)])
for w in (0, 1, 4, 5):
    base = 2100 + w * 400
    print(f"SM w{w} p2 [pre-wait, post-ld, post-exps, end]:")
    for b in range(32):
        print("   ", b, [r(a[base + 4 * b + i]) for i in range(4)])
