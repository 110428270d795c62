"""Box decoder alone at the bench shape (8 frames x 4 phrases, 64x64 grid): eager launches vs CUDA-graph replay.
python profiles/decoder_micro.py [frames] [phrases]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from grove_b200 import ops  # noqa: E402
from grove_b200.modeling.build_sam import sam_model_registry  # noqa: E402

Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G, C = 64, 256
torch.manual_seed(0)
sam = sam_model_registry["vit_b"](None, True, image_size=16 * G).cuda()
md, pe = sam.mask_decoder, sam.prompt_encoder
emb = torch.randn(Fr, G * G, C, device="cuda").to(torch.bfloat16)
text = torch.randn(Fr * P, C, device="cuda")
no_mask = pe.no_mask_embed.weight.reshape(-1).float().contiguous()
dense_pe = pe.get_dense_pe()
reps = [P] * Fr


def run():
    return md.decode_records(emb, dense_pe, text, no_mask, reps)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


ops.reset_launch_count()
run()
torch.cuda.synchronize()
launches = ops.launch_count()
eager = timeit(run)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = run()
replay = timeit(g.replay)
print(f"decoder {Fr} frames x {P} phrases: {launches} grove launches, eager {eager:.1f} us, graph replay {replay:.1f} us")
