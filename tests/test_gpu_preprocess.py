"""Frame pre-processing fused into the patch embed on a B200: bit-exact against the oracle's restatement of the
reference chain (Pillow bilinear resize -> (x - mean) / std -> zero pad -> bf16), which tests/test_oracle_golden.py pins to Pillow."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402
from oracle import preprocess as pp  # noqa: E402


def _patchify(x_bf16, img):
    """[3, T, img, img] bf16 -> [T*G*G, 768] with k = c*256 + py*16 + px (the im2col of PatchEmbed, image_encoder.py:484-491)"""
    G = img // 16
    T = x_bf16.shape[1]
    p = x_bf16.permute(1, 0, 2, 3).reshape(T, 3, G, 16, G, 16).permute(0, 2, 4, 1, 3, 5)
    return p.reshape(T * G * G, 768)


@pytest.mark.parametrize("T,h,w,L", [(2, 360, 640, 512), (1, 720, 1280, 1024), (2, 240, 320, 512), (1, 500, 375, 512), (1, 512, 512, 512),
                                     (1, 1080, 1920, 1024), (1, 100, 37, 64)])
def test_resize_and_fused_patches_bit_exact(T, h, w, L):
    from grove_b200.preprocess import ResizeLongestSide
    rng = np.random.default_rng(h * 7 + w)
    frames = rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8)
    frames[:, ::3] //= 2
    tr = ResizeLongestSide(L)
    dev = torch.from_numpy(frames).cuda()
    ref = np.stack([pp.apply_image(f, L) for f in frames])
    out = tr.apply_image(dev)
    assert tuple(out.shape) == ref.shape
    assert np.array_equal(out.cpu().numpy(), ref)
    # fused: vertical pass + normalise + pad + bf16 + patchify
    patches = tr.patches(dev, L)
    x = torch.from_numpy(pp.grounding_enc_processor(ref, L)).bfloat16()
    want = _patchify(x, L)
    assert torch.equal(patches.cpu().view(torch.int16), want.contiguous().view(torch.int16))


def test_reference_golden_frames():
    """the frames Pillow resized through the reference's own ResizeLongestSide (tests/golden/preprocess.npz)"""
    from grove_b200.preprocess import ResizeLongestSide
    g = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    for tag in ("down", "up", "tall", "same", "odd"):
        T, h, w, L = [int(v) for v in g[f"{tag}.meta"]]
        out = ResizeLongestSide(L).apply_image(torch.from_numpy(g[f"{tag}.frames"]).cuda())
        assert np.array_equal(out.cpu().numpy(), g[f"{tag}.resized"]), tag


def test_encoder_from_uint8_frames_equals_host_pipeline():
    """ImageEncoderViT.forward_frames(uint8) == ImageEncoderViT.forward(reference-style host-processed bf16 tensor), bit for bit"""
    from helpers import encoder_with_weights
    sam, sd, cfg = encoder_with_weights("vit_b", 512, 5)
    enc = sam.image_encoder
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (1, 8, 288, 512, 3), dtype=np.uint8)          # already at the long side: only normalise + pad
    big = rng.integers(0, 256, (1, 8, 360, 640, 3), dtype=np.uint8)             # needs the resize
    for fr in (frames, big):
        res = np.stack([pp.apply_image(f, 512) for f in fr[0]])
        x = torch.from_numpy(pp.grounding_enc_processor(res, 512)).bfloat16().unsqueeze(0).cuda()     # [1,3,T,512,512]
        a = enc(x)
        b = enc.forward_frames(torch.from_numpy(fr).cuda())
        assert torch.equal(a, b)
