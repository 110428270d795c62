"""Per-kernel parity on a B200: each C-ABI entry point against a plain fp32 PyTorch / oracle statement of the same op."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import grounding as og  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from grove_b200 import ops as _ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _relerr(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12))


# ------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("force_ctas", [1, 2])
@pytest.mark.parametrize("M,N,K,max_ctas", [(128, 128, 64, 0), (256, 256, 128, 0), (1100, 768, 768, 0), (4096, 2304, 768, 7),
                                            (8192, 3072, 768, 0), (4096, 768, 3072, 6), (300, 256, 4096, 0), (32768, 128, 256, 0)])
def test_gemm_plain(ops, M, N, K, max_ctas, force_ctas):
    """force_ctas=1: single-CTA 128xBN tiles; force_ctas=2: CTA-pair (cta_group::2) 256x256 tiles when N % 256 == 0"""
    a, w = _rand((M, K), 1, dtype=torch.bfloat16), _rand((N, K), 2, 1 / math.sqrt(K), dtype=torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(a, w, out, max_ctas=max_ctas, force_ctas=force_ctas)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    assert _relerr(out, ref) < 2e-5, _relerr(out, ref)


@pytest.mark.parametrize("force_ctas", [1, 2])
def test_gemm_epilogues(ops, force_ctas):
    M, N, K = 1536, 768, 768
    a, w = _rand((M, K), 3, dtype=torch.bfloat16), _rand((N, K), 4, 1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, resid = _rand((N,), 5), _rand((M, N), 6)
    ref = a.float() @ w.float().t() + bias
    # bias + GELU -> bf16
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, o, bias=bias, act="gelu", force_ctas=force_ctas)
    assert _relerr(o.float(), F.gelu(ref)) < 6e-3
    # bias + in-place fp32 residual + bf16 side copy
    x = resid.clone()
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, x, bias=bias, resid=x, out2=o2, force_ctas=force_ctas)
    assert _relerr(x, resid + ref) < 2e-5
    assert torch.equal(o2, x.to(torch.bfloat16))
    # row-periodic residual (abs-pos embedding), relu, tanh gate
    pos, alpha = _rand((512, N), 7), torch.tensor([0.5], device="cuda")
    o3 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, w, o3, bias=bias, act="relu", gate_alpha=alpha, resid=pos, resid_row_mod=512, force_ctas=force_ctas)
    ref3 = pos.repeat(3, 1) + math.tanh(0.5) * F.relu(ref)
    assert _relerr(o3, ref3) < 2e-5


def test_gemm_few_tiles_long_k_split(ops):
    """The text projection's shapes (a handful of [DET] rows padded to 128, 4096 -> 4096 -> 256): few tiles and 64 k-blocks are spread
    over the SMs as split-K planes + fix-up; bf16 output with ReLU and fp32 output with bias, against fp32 torch."""
    M, K = 128, 4096
    a = _rand((M, K), 51, dtype=torch.bfloat16)
    w0, b0 = _rand((4096, K), 52, 1 / math.sqrt(K), dtype=torch.bfloat16), _rand((4096,), 53)
    h = torch.empty(M, 4096, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w0, h, bias=b0, act="relu")
    assert _relerr(h.float(), F.relu(a.float() @ w0.float().t() + b0)) < 6e-3
    w2, b2 = _rand((256, K), 54, 1 / math.sqrt(K), dtype=torch.bfloat16), _rand((256,), 55)
    o = torch.empty(M, 256, device="cuda", dtype=torch.float32)
    ops.gemm(h, w2, o, bias=b2)
    assert _relerr(o, h.float() @ w2.float().t() + b2) < 2e-5


def test_gemm_tail_split(ops):
    """Wave-quantisation tail: with 12 CTA pairs, 6 m-blocks x 3 n-blocks = 18 tiles leave 6 for a second wave, so the last two
    m-blocks are computed as split-K partial planes + fix-up kernel (long-K problems only); the result must equal the plain schedule
    (fp32 sums in a different order) and the reference."""
    M, N, K = 1536, 768, 6144
    a, w = _rand((M, K), 41, dtype=torch.bfloat16), _rand((N, K), 42, 1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, resid = _rand((N,), 43), _rand((M, N), 44)
    alpha = torch.tensor([0.3], device="cuda")
    ref = resid + math.tanh(0.3) * torch.relu(a.float() @ w.float().t() + bias)
    x = resid.clone()
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, x, bias=bias, resid=x, act="relu", gate_alpha=alpha, out2=o2, max_ctas=24)
    assert _relerr(x, ref) < 2e-5
    assert _relerr(o2.float(), ref) < 6e-3
    y = resid.clone()
    ops.gemm(a, w, y, bias=bias, resid=y, act="relu", gate_alpha=alpha, max_ctas=36)     # 18 pairs: one exact wave, no split
    assert _relerr(x, y) < 5e-6


@pytest.mark.parametrize("force_ctas", [1, 2])
@pytest.mark.parametrize("V,T,G,C,N,kt", [(1, 8, 64, 128, 256, 3), (2, 8, 32, 64, 128, 3), (3, 1, 64, 256, 256, 1), (1, 8, 16, 64, 128, 3)])
def test_conv_implicit_gemm(ops, V, T, G, C, N, kt, force_ctas):
    x = _rand((V, T, G, G, C), 8, dtype=torch.bfloat16)
    bias = _rand((N,), 10)
    if kt == 3:
        w = _rand((N, C, 3, 3, 3), 9, 1 / math.sqrt(27 * C), dtype=torch.bfloat16)
        wp = w.permute(0, 2, 3, 4, 1).reshape(N, -1).contiguous()
        ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), bias, padding=1).permute(0, 2, 3, 4, 1).reshape(-1, N)
    else:
        w = _rand((N, C, 3, 3), 9, 1 / math.sqrt(9 * C), dtype=torch.bfloat16)
        wp = w.permute(0, 2, 3, 1).reshape(N, -1).contiguous()
        ref = F.conv2d(x.float().reshape(V * T, G, G, C).permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    out = torch.empty(V * T * G * G, N, device="cuda", dtype=torch.float32)
    ops.conv_gemm(x, wp, out, V=V, T=T, G=G, kt=kt, bias=bias, force_ctas=force_ctas)
    assert _relerr(out, ref) < 3e-5, _relerr(out, ref)


# ------------------------------------------------------------------ HBM-bound encoder helpers
def test_im2col_layernorm_transposes(ops):
    img = _rand((2, 3, 8, 64, 64), 11, dtype=torch.bfloat16)
    out = torch.empty(16 * 16, 768, device="cuda", dtype=torch.bfloat16)
    ops.im2col_patch16(img, out)
    fr = img.permute(0, 2, 1, 3, 4).reshape(16, 3, 64, 64)
    ref = F.unfold(fr.float(), 16, stride=16).transpose(1, 2).reshape(-1, 768).to(torch.bfloat16)
    assert torch.equal(out, ref)
    for D in (256, 768, 1024, 1280):
        x, g, b = _rand((1000, D), 12, 3.0) + 0.7, _rand((D,), 13) + 1.0, _rand((D,), 14)
        y32 = torch.empty_like(x)
        ops.layernorm(x, g, b, y32, 1e-6)
        assert _relerr(y32, F.layer_norm(x, (D,), g, b, 1e-6)) < 1e-5
        y16 = torch.empty(1000, D, device="cuda", dtype=torch.bfloat16)
        ops.layernorm(x, g, b, y16, 1e-6)
        assert _relerr(y16.float(), F.layer_norm(x, (D,), g, b, 1e-6)) < 5e-3
    t = _rand((3, 1024, 256), 15, dtype=torch.bfloat16)
    nchw = torch.empty(3, 256, 1024, device="cuda", dtype=torch.bfloat16)
    ops.tokens_to_nchw(t, nchw, 3, 1024, 256)
    assert torch.equal(nchw, t.transpose(1, 2).contiguous())
    back = torch.empty_like(t)
    ops.nchw_to_tokens(nchw, back, 3, 1024, 256)
    assert torch.equal(back, t)


# ------------------------------------------------------------------ attention with decomposed rel-pos bias
def _ref_attention(qkv, rel_h, rel_w, B, S, heads, hd):
    """Attention.forward without qkv/proj linears (image_encoder.py:306-323), fp32, on [B,S,S,3,heads,hd]."""
    q, k, v = qkv.float().reshape(B, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, B * heads, S * S, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh, Rw = og.rel_pos_table(S, rel_h.float()), og.rel_pos_table(S, rel_w.float())
    rq = q.reshape(B * heads, S, S, hd)
    attn = (attn.view(B * heads, S, S, S, S) + torch.einsum("bhwc,hkc->bhwk", rq, Rh)[..., :, None]
            + torch.einsum("bhwc,wkc->bhwk", rq, Rw)[..., None, :]).view(B * heads, S * S, S * S)
    o = attn.softmax(-1) @ v
    return o.view(B, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(B, S, S, heads * hd)


@pytest.mark.parametrize("legacy,hd", [(False, 64), (True, 64), (False, 80)])
@pytest.mark.parametrize("G,Fr,heads", [(64, 2, 3), (32, 3, 2)])
def test_global_attention(ops, G, Fr, heads, legacy, hd):
    """legacy=False: the tcgen05/TMEM kernel the modules use (head dim 64 = ViT-B/L, 80 = ViT-H); legacy=True: the mma.sync cross-check"""
    qkv = _rand((Fr, G, G, 3, heads, hd), 20, dtype=torch.bfloat16)
    rh, rw = _rand((2 * G - 1, hd), 21, 0.1, dtype=torch.bfloat16), _rand((2 * G - 1, hd), 22, 0.1, dtype=torch.bfloat16)
    out = torch.full((Fr, G, G, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attn_global(qkv, rh, rw, out, F=Fr, G=G, heads=heads, hd=hd, legacy_mma=legacy)
    torch.cuda.synchronize()
    ref = _ref_attention(qkv, rh, rw, Fr, G, heads, hd)
    err = float((out.float() - ref).abs().max())
    assert err < 2e-2, err          # bf16 P and bf16 output on O(1) values
    assert float((out.float() - ref).abs().mean()) < 2e-3


@pytest.mark.parametrize("hd,G", [(64, 64), (80, 64), (64, 32)])
def test_global_attention_growing_scores(ops, hd, G):
    """Keys whose norm grows along the sequence: the running row maximum of the single-pass softmax rises by far more than its lazy
    rescale threshold (2^8) several times per query tile, so the in-TMEM accumulator correction is exercised; also checks the
    log-sum-exp the backward pass consumes."""
    Fr, heads = 1, 2
    N = G * G
    qkv = _rand((Fr, G, G, 3, heads, hd), 27, dtype=torch.bfloat16).float()
    ramp = torch.linspace(0.25, 9.0, N, device="cuda").view(1, G, G, 1, 1)
    qkv[:, :, :, 1] *= ramp                                  # logits' spread grows ~36x from the first to the last key block
    qkv[:, :, :, 0] *= 1.5
    qkv = qkv.to(torch.bfloat16)
    rh, rw = _rand((2 * G - 1, hd), 28, 0.1, dtype=torch.bfloat16), _rand((2 * G - 1, hd), 29, 0.1, dtype=torch.bfloat16)
    out = torch.full((Fr, G, G, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((Fr * N, heads), float("nan"), device="cuda")
    ops.attn_global(qkv, rh, rw, out, F=Fr, G=G, heads=heads, hd=hd, lse=lse)
    torch.cuda.synchronize()
    ref = _ref_attention(qkv, rh, rw, Fr, G, heads, hd)
    assert float((out.float() - ref).abs().max()) < 3e-2
    assert float((out.float() - ref).abs().mean()) < 2e-3
    # log2-domain log-sum-exp of the biased scores
    q, k, _ = qkv.float().reshape(Fr, N, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, Fr * heads, N, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh, Rw = og.rel_pos_table(G, rh.float()), og.rel_pos_table(G, rw.float())
    rq = q.reshape(Fr * heads, G, G, hd)
    attn = (attn.view(Fr * heads, G, G, G, G) + torch.einsum("bhwc,hkc->bhwk", rq, Rh)[..., :, None]
            + torch.einsum("bhwc,wkc->bhwk", rq, Rw)[..., None, :]).view(Fr * heads, N, N)
    ref_lse = (torch.logsumexp(attn, -1) / math.log(2.0)).view(Fr, heads, N).permute(0, 2, 1).reshape(Fr * N, heads)
    assert float((lse - ref_lse).abs().max()) < 5e-2


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("hd", [64, 80])
@pytest.mark.parametrize("G,Fr,heads", [(64, 2, 2), (32, 1, 3), (28, 1, 1), (64, 3, 16)])
def test_window_attention_with_padding(ops, G, Fr, heads, hd, tc):
    """tc=True: the tcgen05/TMEM kernel the modules use; tc=False: the mma.sync cross-check kernel"""
    ws = 14
    D = heads * hd
    qkv_bias = _rand((3 * D,), 23, 0.5)
    qkv = _rand((Fr, G, G, 3, heads, hd), 24, dtype=torch.bfloat16)
    rh, rw = _rand((2 * ws - 1, hd), 25, 0.1, dtype=torch.bfloat16), _rand((2 * ws - 1, hd), 26, 0.1, dtype=torch.bfloat16)
    out = torch.full((Fr, G, G, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    if tc:
        ops.attn_window_tc(qkv, qkv_bias.to(torch.bfloat16), ops.window_rel_table(rh, rw), out, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
    else:
        ops.attn_window(qkv, qkv_bias.to(torch.bfloat16), rh, rw, out, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
    torch.cuda.synchronize()
    # reference: pad tokens carry qkv = bias (x = 0 after norm1, image_encoder.py:245-249); partition, attend, unpartition
    Gp = ((G + ws - 1) // ws) * ws
    full = qkv_bias.to(torch.bfloat16).float().view(1, 1, 1, 3 * D).expand(Fr, Gp, Gp, 3 * D).clone()
    full[:, :G, :G] = qkv.float().reshape(Fr, G, G, 3 * D)
    win, pad_hw = og.window_partition(full, ws)
    o = _ref_attention(win.reshape(-1, ws, ws, 3, heads, hd), rh, rw, win.shape[0], ws, heads, hd)
    ref = og.window_unpartition(o, ws, pad_hw, (G, G))
    err = float((out.float() - ref).abs().max())
    assert err < 2e-2, err
    assert float((out.float() - ref).abs().mean()) < 2e-3


# ------------------------------------------------------------------ box utilities: bit-exact decisions
def test_box_utilities_bit_exact(ops):
    from conftest import GOLDEN
    import os
    from grove_b200 import box_eval as be
    g = np.load(os.path.join(GOLDEN, "box_eval.npz"))
    assert np.array_equal(be.np_box_iou(g["b1"], g["b2"]), g["iou64"], equal_nan=True)
    assert np.array_equal(be.np_box_iou(g["b1"].astype(np.float32), g["b2"].astype(np.float32)), g["iou32"], equal_nan=True)
    mat = be.compute_iou_matrix(g["p1"].tolist(), g["p2"].tolist())
    assert np.array_equal(mat, g["mat"])
    assert np.array_equal(np.array(be.greedy_matches(mat, g["sims"], 0.3, 0.2), dtype=np.int64).reshape(-1, 2), g["matches"])
    assert np.array_equal(be.bbox_overlaps_batch(g["anc"], g["gtb"], g["frm"]).numpy(), g["ov"])
    assert np.array_equal(be.bbox_overlaps_batch(g["anc"], g["gtb"], np.zeros_like(g["frm"])).numpy(), g["ov_nomask"])
    # larger random problem against the oracle port, incl. threshold decisions
    from oracle import box_eval as ob
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 100, (200, 2)); a = np.concatenate([xy, xy + rng.uniform(0, 60, (200, 2))], 1)
    xy = rng.uniform(0, 100, (150, 2)); b = np.concatenate([xy, xy + rng.uniform(0, 60, (150, 2))], 1)
    assert np.array_equal(be.np_box_iou(a, b), ob.np_box_iou(a, b))
    assert np.array_equal(be.compute_iou_matrix(np.round(a), np.round(b)), ob.compute_iou_matrix(np.round(a), np.round(b)))
    iou = ob.compute_iou_matrix(np.round(a[:40]), np.round(b[:30])); sim = rng.uniform(0, 1, iou.shape)
    assert be.greedy_matches(iou, sim, 0.1, 0.3) == ob.greedy_match(iou, sim, 0.1, 0.3)
    # post-process decisions: sigmoid(logit) > thr on a grid that straddles the threshold
    logits = torch.cat([torch.linspace(-3, 3, 4001), torch.tensor([0.0, 1e-8, -1e-8, 1e-4, -1e-4])]).cuda()
    boxes = torch.rand(logits.numel(), 4, device="cuda")
    size = torch.tensor([[1280.0, 720.0]], device="cuda").expand(logits.numel(), 2).contiguous()
    for thr in (0.5, 0.3, 0.7):
        xyxy, keep = ops.box_postprocess(boxes, logits, size, thr)
        assert torch.equal(keep.bool(), torch.sigmoid(logits) > thr)
        ref = og.box_cxcywh_to_xyxy(og.unnormalize_bboxes(boxes, 1280.0, 720.0))
        assert torch.equal(xyxy, ref)


def test_decision_utilities_bit_exact(ops):
    """centre-in-box, video IoU + strict recall flags and the validation sums on the GPU against the REFERENCE's own code
    (tests/golden/decisions.npz) and, on larger random problems, against the oracle port: decisions and float64 values bit-exact."""
    from conftest import GOLDEN
    import os
    from grove_b200 import box_eval as be
    from oracle import box_eval as ob
    g = np.load(os.path.join(GOLDEN, "decisions.npz"))
    # the reference-shaped call: dicts of clips, empty ground truth skipped, None / NaN predictions valid but wrong
    pred, gt, kinds = g["cib_pred"], g["cib_gt"], g["cib_kinds"]
    gt_data, pred_dict = [], {}
    for c in range(3):
        sl = slice(8 * c, 8 * c + 8)
        gt_data.append({"video_id": f"v{c}", "segment_youcook_idx": c,
                        "segment_bboxes": [[] if k == 1 else tuple(float(v) for v in b) for b, k in zip(gt[sl], kinds[sl])]})
        pred_dict[f"v{c}_{c}"] = {"final_boxes": [None if k == 2 else np.array([p]) for p, k in zip(pred[sl], kinds[sl])]}
    assert list(be.evaluate_dataset_localization(pred_dict, gt_data, "youcook")) == g["cib_result"].tolist()
    ok = np.isin(kinds, (0, 3))
    assert np.array_equal(be.center_in_box(pred[ok], gt[ok]), g["cib_flags"][ok])
    for v in range(g["viou_gt"].shape[0]):
        viou, over, ious = be.viou_over_threshold(g["viou_pred"][v], g["viou_gt"][v], (0.3, 0.5))
        assert viou == g["viou_value"][v] and [over[0.3], over[0.5]] == g["viou_over"][v].tolist()
        assert np.array_equal(ious, g["viou_frame"][v])
    pb, lg, go, gtb = g["val_boxes"], g["val_logits"], g["val_obj"], g["val_gt"]
    V, T = pb.shape[:2]
    for variant, cast in (("f", lambda a: a), ("i", lambda a: a.astype(np.int32))):
        out = be.val_giou_and_objectness_accuracy([[pb[v, f] for f in range(T)] for v in range(V)], [[lg[v, f] for f in range(T)] for v in range(V)],
                                                   [[cast(gtb[v, f][go[v, f].astype(bool)]) for f in range(T)] for v in range(V)],
                                                   [[go[v, f] for f in range(T)] for v in range(V)])
        ref = g[f"val_{variant}"]
        assert list(out[1:]) == ref[1:].tolist()
        np.testing.assert_allclose(out[0], ref[0], rtol=1e-6)
    # larger random problems vs the oracle port
    rng = np.random.default_rng(1)
    n = 5000
    gt2 = np.round(rng.uniform(0, 100, (n, 2))); gt2 = np.concatenate([gt2, gt2 + np.round(rng.uniform(1, 50, (n, 2)))], 1)
    pr2 = np.round(gt2 + rng.uniform(-40, 40, (n, 4)))          # integer coordinates: many centres land exactly on a box edge
    flags = be.center_in_box(pr2, gt2)
    want = np.array([ob.center_in_box(p, q) for p, q in zip(pr2.tolist(), gt2.tolist())], dtype=np.uint8)
    assert np.array_equal(flags, want) and 0 < want.sum() < n
    for trial in range(20):
        m = int(rng.integers(1, 40))
        q = rng.uniform(0, 100, (m, 2)); q = np.concatenate([q, q + rng.uniform(5, 60, (m, 2))], 1)
        p_ = q + rng.uniform(-15, 15, (m, 4))
        p_[rng.uniform(size=m) < 0.1] = 0.0
        thr = (0.1, 0.3, 0.5, 0.7)
        viou, over, ious = be.viou_over_threshold(p_, q, thr)
        rv, rover, rious = ob.video_viou(p_, q, thr)
        assert viou == rv and [over[t] for t in thr] == rover and np.array_equal(ious, np.array(rious, dtype=np.float64))


@pytest.mark.parametrize("force_ctas", [1, 2])
def test_gemm_bf16_residual_stream(ops, force_ctas):
    """The residual-stream GEMMs with the stream kept in bf16 (EPI = 4 on CTA pairs, the general epilogue on single-CTA tiles, the
    tail-split fix-up): out = bf16(resid_bf16 + gate * act(acc + bias)), sum in fp32, ONE rounding — equal to rounding the fp32-stream
    result of the same kernel family, up to one bf16 ulp where the fp32 sums differ in the last bit."""
    M, N, K = 1536, 768, 768
    a, w = _rand((M, K), 3, dtype=torch.bfloat16), _rand((N, K), 4, 1 / math.sqrt(K), dtype=torch.bfloat16)
    bias, resid = _rand((N,), 5), _rand((M, N), 6, dtype=torch.bfloat16)
    ref = resid.float() + a.float() @ w.float().t() + bias
    x = resid.clone()
    ops.gemm(a, w, x, bias=bias, resid=x, force_ctas=force_ctas)           # in place
    d = (x.float() - ref).abs()
    assert float((d / (ref.abs() + 1e-3)).max()) < 8e-3 and float(d.mean()) < 3e-3       # bf16 rounding of an O(1) value
    # ReLU + tanh gate (the adapter's epilogue) into a second buffer
    alpha = torch.tensor([0.5], device="cuda")
    y = torch.empty_like(resid)
    ops.gemm(a, w, y, bias=bias, act="relu", gate_alpha=alpha, resid=resid, force_ctas=force_ctas)
    ref2 = resid.float() + math.tanh(0.5) * F.relu(a.float() @ w.float().t() + bias)
    assert float(((y.float() - ref2).abs() / (ref2.abs() + 1e-3)).max()) < 8e-3
    if force_ctas == 2:
        # ragged M on the CTA-pair epilogue (rows past the end: residual fetch clamped, nothing stored) with the row statistics of the
        # ROUNDED output, one (sum, sum of squares) pair per 128 columns
        Mr = 1400
        xr = torch.full((Mr + 8, N), 3.0, device="cuda", dtype=torch.bfloat16)
        xr[:Mr] = resid[:Mr]
        stats = torch.zeros(Mr, N // 128, 2, device="cuda")
        ops.gemm(a[:Mr], w, xr[:Mr], bias=bias, resid=xr[:Mr], ln_stats_out=stats, force_ctas=2)
        assert bool((xr[Mr:] == 3.0).all())
        assert torch.equal(xr[:Mr], x[:Mr])
        v = xr[:Mr].float().view(Mr, N // 128, 128)
        assert torch.allclose(stats[..., 0], v.sum(-1), rtol=1e-5, atol=2e-3)
        assert torch.allclose(stats[..., 1], (v * v).sum(-1), rtol=1e-5, atol=2e-3)
    # LayerNorm on a bf16 row equals LayerNorm on its fp32 widening
    g, b = _rand((N,), 7) + 1.0, _rand((N,), 8)
    h16, h32 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.layernorm(x, g, b, h16, 1e-6)
    ops.layernorm(x.float().contiguous(), g, b, h32, 1e-6)
    assert torch.equal(h16, h32)


def test_conv_adapter_bf16_stream_tail_split(ops):
    """The Conv3d adapter at its real size (8 frames x 64x64 x 768 -> 384 tiles on 74 CTA pairs: the trailing m-blocks go through the split-K
    planes + fix-up kernel) with a bf16 stream: equals the fp32-stream result rounded to bf16 (<= 1 ulp)."""
    V, T, G, C = 1, 8, 64, 768
    x = _rand((V * T * G * G, C), 11, dtype=torch.bfloat16)
    w = _rand((C, 27 * C), 12, 1 / math.sqrt(27 * C), dtype=torch.bfloat16)
    bias, alpha = _rand((C,), 13), torch.tensor([0.5], device="cuda")
    o32 = torch.empty(V * T * G * G, C, device="cuda", dtype=torch.float32)
    ops.conv_gemm(x, w, o32, V=V, T=T, G=G, kt=3, bias=bias, act="relu", gate_alpha=alpha, resid=x.float().contiguous())
    o16 = torch.empty_like(x)
    ops.conv_gemm(x, w, o16, V=V, T=T, G=G, kt=3, bias=bias, act="relu", gate_alpha=alpha, resid=x)
    d = (o16.float() - o32).abs()
    assert float((d / (o32.abs() + 1e-3)).max()) < 8e-3 and float(d.mean()) < 3e-3
    assert float((o16.float() - o32.to(torch.bfloat16).float()).abs().max()) < 0.04      # at most one bf16 ulp of an O(1..4) value


@pytest.mark.parametrize("act", [None, "gelu"])
def test_gemm_layernorm_fold(ops, act):
    """LayerNorm folded into the consuming GEMM: the residual GEMM's epilogue emits per-row (sum, sum of squares) partials of the bf16
    stream it writes (ln_stats_out), the next GEMM takes the RAW stream as A, W * gamma as weights and applies
    rstd * (acc - mu * colsum) + (b + W.beta) in its epilogue — against LayerNorm -> Linear (-> GELU) in fp32 on the same bf16 stream."""
    M, D, N = 2048, 768, 2304 if act is None else 3072
    a, w = _rand((M, D), 3, dtype=torch.bfloat16), _rand((D, D), 4, 1 / math.sqrt(D), dtype=torch.bfloat16)
    bias = _rand((D,), 5)
    resid = (_rand((M, D), 6) * 2.0 + _rand((1, D), 9) * 3.0).to(torch.bfloat16)      # per-channel offsets: a non-trivial row mean
    stats = torch.empty(M, D // 128, 2, device="cuda", dtype=torch.float32)
    xs = resid.clone()
    ops.gemm(a, w, xs, bias=bias, resid=xs, ln_stats_out=stats)
    x32 = xs.float()
    np.testing.assert_allclose(stats[..., 0].sum(1).cpu().numpy(), x32.sum(1).cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[..., 1].sum(1).cpu().numpy(), (x32 * x32).sum(1).cpu().numpy(), rtol=1e-4)
    g, be = _rand((D,), 7) * 0.3 + 1.0, _rand((D,), 8) * 0.3
    w2, b2 = _rand((N, D), 10, 1 / math.sqrt(D)), _rand((N,), 11)
    wg = (w2 * g[None, :]).to(torch.bfloat16).contiguous()
    bf = (b2 + w2 @ be).contiguous()
    cs = wg.float().sum(1).contiguous()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(xs, wg, out, bias=bf, act=act, ln_fold=(stats, cs, 1e-6))
    ref = F.layer_norm(x32, (D,), g, be, 1e-6) @ w2.t() + b2
    if act == "gelu":
        ref = F.gelu(ref)
    # unfused path on the same stream: LayerNorm kernel -> bf16 -> GEMM
    h = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.layernorm(xs, g, be, h, 1e-6)
    out_u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(h, w2.to(torch.bfloat16).contiguous(), out_u, bias=b2, act=act)
    e_f, e_u = _relerr(out.float(), ref), _relerr(out_u.float(), ref)
    print(f"LN fold act={act}: rel err folded {e_f:.2e}, unfused {e_u:.2e}")
    assert e_f < 6e-3 and e_f < 1.5 * e_u + 1e-3
