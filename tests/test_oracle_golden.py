"""Pins the oracle (oracle/*.py) to outputs of the reference's own modules (tests/golden/*.npz,
written by oracle/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import box_eval, grounding as og, synth


def _g(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _encoder_case(name):
    g = _g(name)
    D, depth, heads, img, seed, *gidx = [int(x) for x in g["meta"]]
    sd = synth.synth_state_dict(synth.encoder_param_shapes(D, depth, heads, gidx, img // 16), seed)
    images = synth.synth_tensor(name + ".images", (1, 3, 8, img, img), seed)
    with torch.no_grad():
        out = og.image_encoder(images, sd, depth=depth, heads=heads, global_idx=gidx, pre="image_encoder.")
    np.testing.assert_allclose(out[:, ::4, ::2, ::2].numpy(), g["out_sub"], rtol=0, atol=2e-4)
    with torch.no_grad():   # tight pin: oracle in float64 vs the reference in float64
        sd64 = {k: v.double() for k, v in sd.items()}
        out64 = og.image_encoder(synth.synth_tensor(name + ".images", (1, 3, 8, img, img), seed).double(), sd64,
                                 depth=depth, heads=heads, global_idx=gidx, pre="image_encoder.")
    np.testing.assert_allclose(out64[:, ::4, ::2, ::2].numpy(), g["out64_sub"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out.double().sum((2, 3)).numpy(), g["out_chan_sum"], rtol=0, atol=2e-2)


def test_encoder_512_verbatim_reference_adapter():
    _encoder_case("enc_tiny512_verbatim")


def test_encoder_256_padded_windows():
    _encoder_case("enc_tiny256_padded")


def _decoder_case(name):
    g = _g(name)
    dim, mlp, G, frames, seed = [int(x) for x in g["meta"]]
    reps = [int(r) for r in g["reps"]]
    sd = synth.synth_state_dict(synth.decoder_param_shapes(dim, mlp), seed)
    emb = synth.synth_tensor(name + ".emb", (frames, dim, G, G), seed)
    txt = synth.synth_tensor(name + ".txt", (sum(reps), 1, dim), seed)
    with torch.no_grad():
        pe = og.dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], G)
        boxes, logits = og.box_decoder(emb, pe, txt, reps, sd)
    st = max(G // 4, 1)
    np.testing.assert_allclose(pe[0, :, ::st, ::st].numpy(), g["dense_pe_sample"], atol=1e-5)
    # fp32 vs the reference's fp32 (both carry fp32 round-off; the reference's own fp32-vs-fp64 gap here is <1e-6) ...
    np.testing.assert_allclose(boxes.numpy(), g["boxes"], atol=5e-6)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-5)
    # ... and the tight pin is float64 oracle vs float64 reference
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        pe64 = og.dense_pe(sd64["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], G)
        b64, l64 = og.box_decoder(emb.double(), pe64, txt.double(), reps, sd64)
    np.testing.assert_allclose(b64.numpy(), g["boxes64"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(l64.numpy(), g["logits64"], rtol=0, atol=1e-9)
    # decisions on the path are bit-exact against the reference's floats
    assert ((torch.sigmoid(l64) > 0.5).numpy() == (torch.sigmoid(torch.from_numpy(g["logits64"])) > 0.5).numpy()).all()


def test_decoder_config1_full_size():
    _decoder_case("dec_cfg1_full")


def test_decoder_ragged_with_empty_frame():
    _decoder_case("dec_ragged")


def test_glue_methods_and_losses():
    g = _g("glue")
    seed, dim, mlp, G, T, hidden, L = 5, 64, 128, 8, 8, 96, 600
    pos = [synth.det_positions(L, 3, seed), synth.det_positions(L, 2, seed + 1)]
    ids = torch.full((2, L - 575), 7, dtype=torch.long)
    for v, pp in enumerate(pos):
        for p in pp:
            ids[v, p - 575 + 1] = 32005
    mask = og.create_det_token_mask(ids, 32005)
    assert (mask.numpy() == g["det_mask"]).all()
    sd = synth.synth_state_dict({**synth.decoder_param_shapes(dim, mlp), **synth.text_fcs_shapes(hidden, dim)}, seed)
    hid = synth.synth_tensor("glue.hidden", (2, L, hidden), seed)
    emb = synth.synth_tensor("glue_dec.emb", (2 * T, dim, G, G), seed)
    with torch.no_grad():
        pred = og.process_hidden_states(hid, mask, sd, T)
        assert [p.shape[0] for p in pred] == g["counts"].tolist()
        np.testing.assert_allclose(torch.cat(pred).numpy(), g["pred_embeddings"], atol=1e-5)
        reps = [p.shape[0] for p in pred]
        pe = og.dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], G)
        boxes, logits = og.box_decoder(emb, pe, torch.cat(pred).unsqueeze(1), reps, sd)
        tb, tl = og.postprocess(boxes, logits, reps, T, None, infer=False)
        ib, _ = og.postprocess(boxes, logits, reps, T, [(1280, 720), (640, 360)], infer=True)
    np.testing.assert_allclose(torch.cat([b for v in tb for b in v]).numpy(), g["train_boxes"], atol=2e-5)
    np.testing.assert_allclose(torch.cat([l for v in tl for l in v]).numpy(), g["train_logits"], atol=2e-4)
    assert [b.shape[0] for v in ib for b in v] == g["infer_counts"].tolist()      # threshold decisions: exact
    np.testing.assert_allclose(torch.cat([b for v in ib for b in v]).numpy(), g["infer_boxes"], atol=2e-2)
    # losses on the reference's ground truth
    gb_all, go_all = torch.from_numpy(g["gt_boxes"]), torch.from_numpy(g["gt_obj"])
    gt_b, gt_o, ob, oo = [], [], 0, 0
    for v in range(2):
        P = len(pos[v])
        fb, fo = [], []
        for f in range(T):
            o = go_all[oo:oo + P]
            oo += P
            n = int(o.sum())
            fb.append(gb_all[ob:ob + n])
            ob += n
            fo.append(o)
        gt_b.append(fb)
        gt_o.append(fo)
    loss = og.loss_components(tb, tl, gt_b, gt_o, torch.tensor(0.25), 1.0, 2.0, 2.0)
    got = np.array([float(loss[k]) for k in ("loss", "ce_loss", "giou_loss", "l1_loss", "temp_objectness_loss")])
    np.testing.assert_allclose(got, g["losses"], rtol=2e-5)


def test_box_eval_utilities_bit_exact():
    g = _g("box_eval")
    assert np.array_equal(box_eval.np_box_iou(g["b1"], g["b2"]), g["iou64"], equal_nan=True)
    assert np.array_equal(box_eval.np_box_iou(g["b1"].astype(np.float32), g["b2"].astype(np.float32)), g["iou32"], equal_nan=True)
    mat = box_eval.compute_iou_matrix(g["p1"].tolist(), g["p2"].tolist())
    assert np.array_equal(mat, g["mat"])
    m = box_eval.greedy_match(mat, g["sims"], 0.3, 0.2)
    assert np.array_equal(np.array(m, dtype=np.int64).reshape(-1, 2), g["matches"])
    assert np.array_equal(box_eval.bbox_overlaps_batch(g["anc"], g["gtb"], g["frm"]), g["ov"])
    assert np.array_equal(box_eval.bbox_overlaps_batch(g["anc"], g["gtb"], np.zeros_like(g["frm"])), g["ov_nomask"])
    for n in (8, 48, 50, 61, 128):
        idx, masks = box_eval.sliding_segment_with_mask(n, 8)
        assert sum(idx, []) == g[f"seg_idx_{n}"].tolist()
        assert sum(masks, []) == g[f"seg_mask_{n}"].tolist()
        assert [len(r) for r in idx] == g[f"seg_len_{n}"].tolist()


def test_training_step_gradients_vs_reference_autograd():
    """BASELINE config 4 in miniature: autograd over the ORACLE in float64 against the reference's own modules + GROVE methods run
    with autograd in float64 (tests/golden/train_tiny512.npz: losses, gradient norms and strided gradient samples of every
    parameter the loss reaches, and of the LLM hidden states).  This is the pin behind tests/test_gpu_training.py."""
    g = _g("train_tiny512")
    D, depth, heads, img, dim, mlp, hidden, L, P, seed = [int(x) for x in g["meta"]]
    T, gidx = 8, (1, 2)
    sd = {**synth.synth_state_dict(synth.encoder_param_shapes(D, depth, heads, gidx, img // 16), seed),
          **synth.synth_state_dict(synth.decoder_param_shapes(dim, mlp), seed), **synth.synth_state_dict(synth.text_fcs_shapes(hidden, dim), seed)}
    sd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    images = synth.synth_tensor("train_tiny512.images", (1, 3, T, img, img), seed).double()
    hid = synth.synth_tensor("train_tiny512.hidden", (1, L, hidden), seed).double().requires_grad_(True)
    ids = torch.full((1, L - 575), 7, dtype=torch.long)
    for p_ in synth.det_positions(L, P, seed):
        ids[0, p_ - 575 + 1] = 32005
    mask = og.create_det_token_mask(ids, 32005)
    _, boxes, logits, reps = og.grounding_forward(images, hid, mask, sd, depth=depth, heads=heads, global_idx=gidx, num_frames=T)
    obj = torch.from_numpy(g["gt_obj"]).reshape(T, P)
    gtb = torch.from_numpy(g["gt_boxes"])
    gt_b, gt_o, s = [[]], [[]], 0
    for f in range(T):
        n = int(obj[f].sum())
        gt_b[0].append(gtb[s:s + n]); gt_o[0].append(obj[f]); s += n
    pb, pl = og.postprocess(boxes, logits, reps, T, None, infer=False)
    losses = og.loss_components(pb, pl, gt_b, gt_o, torch.zeros((), dtype=torch.float64), 1.0, 2.0, 2.0)
    np.testing.assert_allclose([float(losses[k]) for k in ("loss", "ce_loss", "giou_loss", "l1_loss", "temp_objectness_loss")], g["losses"],
                               rtol=1e-6)
    losses["loss"].backward()
    checked = 0
    for key in g.files:
        if not key.startswith("gn:"):
            continue
        name = key[3:]
        t = hid if name == "hidden" else sd[name]
        assert t.grad is not None, name
        gr = t.grad.reshape(-1)
        step = max(gr.numel() // 64, 1)
        scale = max(float(g[key]), 1e-12)
        assert abs(float(gr.norm()) - float(g[key])) <= 2e-6 * scale + 1e-13, name
        np.testing.assert_allclose(gr[::step][:64].numpy(), g["gs:" + name], rtol=0, atol=2e-6 * scale + 1e-13, err_msg=name)
        checked += 1
    assert checked >= 150
    # and nothing else receives a gradient (the prompt encoder is frozen in GROVE and was not recorded: train.py:281-296)
    reached = {k[3:] for k in g.files if k.startswith("gn:")}
    for k, v in sd.items():
        if k not in reached and not k.startswith("prompt_encoder."):
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k


def test_preprocessing_vs_reference_pil_chain():
    """ResizeLongestSide.apply_image (Pillow bilinear through the reference's own class) and grounding_enc_processor + .bfloat16()
    (infer_iground.py:304-318, train.py:751-753): the numpy restatement is bit-exact (tests/golden/preprocess.npz)."""
    from oracle import preprocess as pp
    g = _g("preprocess")
    for tag in ("down", "up", "tall", "same", "odd"):
        T, h, w, L = [int(v) for v in g[f"{tag}.meta"]]
        frames = g[f"{tag}.frames"]
        out = np.stack([pp.apply_image(f, L) for f in frames])
        assert out.shape == g[f"{tag}.resized"].shape, tag
        assert np.array_equal(out, g[f"{tag}.resized"]), tag
    x = torch.from_numpy(pp.grounding_enc_processor(g["proc.frames"], 512)).bfloat16()
    assert list(x.shape) == [int(v) for v in g["proc.shape"]]
    assert np.array_equal(x.view(torch.int16)[:, :, :48, :72].numpy(), g["proc.bits_sub"])
    assert int(g["proc.pad_nonzero"]) == 0 and float(x[:, :, 40:, :].abs().max()) == 0.0


def _clip_inputs(name="clip_adapters"):
    g = _g(name)
    C, b, seed = [int(x) for x in g["meta"]]
    x = synth.synth_tensor(name + ".adapter.x", (b * 8, 257, C), seed)
    w = synth.synth_tensor(name + ".adapter.w", (C, C, 3, 3, 3), seed) * (27 * C) ** -0.5
    bias = synth.synth_tensor(name + ".adapter.b", (C,), seed) * 0.1
    pools = {"pool_a": synth.synth_tensor(f"{name}.pool_a.x", (8, 256, 16), seed), "pool_b": synth.synth_tensor(f"{name}.pool_b.x", (8, 576, 8), seed)}
    return g, x, w, bias, pools


def test_clip_adapters_vs_reference_classes():
    """oracle/clip_adapters.py against the reference's SpatioTemporalConvAdapter (modeling_clip.py:591-612) and
    AdaptiveAvgPooling3D (pooling.py:6-25) executed in float64 (tests/golden/clip_adapters.npz)."""
    from oracle import clip_adapters as oc
    g, x, w, bias, pools = _clip_inputs()
    alpha = torch.tensor([0.5])
    with torch.no_grad():
        y64 = oc.clip_st_adapter(x.double(), w.double(), bias.double(), alpha.double())
        y32 = oc.clip_st_adapter(x, w, bias, alpha)
    np.testing.assert_allclose(y64.numpy(), g["adapter64"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(y32.numpy(), g["adapter64"], rtol=0, atol=2e-5)
    for k, xp in pools.items():
        np.testing.assert_allclose(oc.adaptive_avgpool3d_tokens(xp.double()).numpy(), g[k + "64"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(oc.adaptive_avgpool3d_tokens(xp).numpy(), g[k + "64"], rtol=0, atol=1e-6)


def test_decision_utilities_vs_reference():
    """centre-in-box (eval_youcookinteractions.py:8-51), video IoU with strict '>' recalls (eval_vidstg.py:157-186) and the validation
    GIoU / objectness-accuracy sums (train.py:821-840) against the reference's own code (tests/golden/decisions.npz): decisions exact."""
    g = _g("decisions")
    acc, correct, valid, flags = box_eval.localization_accuracy(g["cib_pred"], g["cib_gt"], g["cib_kinds"])
    assert [acc, correct, valid] == g["cib_result"].tolist()
    assert np.array_equal(flags, g["cib_flags"])
    for v in range(g["viou_gt"].shape[0]):
        viou, over, ious = box_eval.video_viou(g["viou_pred"][v], g["viou_gt"][v], [0.3, 0.5])
        assert viou == g["viou_value"][v] and over == g["viou_over"][v].tolist()          # bit-exact float64, exact flags
        assert np.array_equal(np.array(ious, dtype=np.float64), g["viou_frame"][v])
    assert g["viou_value"][4] == 0.5 and g["viou_over"][4].tolist() == [1, 0]             # the strictness case is really in the fixture
    pb, lg, go, gt = g["val_boxes"], g["val_logits"], g["val_obj"], g["val_gt"]
    V, T = pb.shape[:2]
    for variant, cast in (("f", lambda a: a), ("i", lambda a: a.astype(np.int32))):
        out = box_eval.val_giou_and_objectness_accuracy([[pb[v, f] for f in range(T)] for v in range(V)], [[lg[v, f] for f in range(T)] for v in range(V)],
                                                         [[cast(gt[v, f][go[v, f].astype(bool)]) for f in range(T)] for v in range(V)],
                                                         [[go[v, f] for f in range(T)] for v in range(V)])
        ref = g[f"val_{variant}"]
        assert list(out[1:]) == ref[1:].tolist()                                           # hits and counts: exact
        np.testing.assert_allclose(out[0], ref[0], rtol=1e-6)


def test_infer_postprocess_vs_reference():
    """_generate_and_postprocess_masks(infer=True) at the real decoder width with a threshold that splits the predictions
    (tests/golden/glue_infer.npz): un-normalise, xyxy, keep sigmoid(logit) > thr, ragged per-frame lists."""
    g = _g("glue_infer")
    dim, mlp, G, T, seed = [int(x) for x in g["meta"]]
    reps = g["reps"].tolist()
    sd = synth.synth_state_dict(synth.decoder_param_shapes(dim, mlp), seed)
    emb = synth.synth_tensor("glue_infer_dec.emb", (2 * T, dim, G, G), seed)
    txt = synth.synth_tensor("glue_infer_dec.txt", (sum(reps), 1, dim), seed)
    with torch.no_grad():
        pe = og.dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"], G)
        boxes, logits = og.box_decoder(emb.double(), pe.double(), txt.double(), reps, {k: v.double() for k, v in sd.items()})
        sizes = [tuple(int(x) for x in s) for s in g["sizes"]]
        ib, il = og.postprocess(boxes.float(), logits.float(), reps, T, sizes, infer=True, thr=float(g["thr"]))
    assert 0 < g["infer_counts"].sum() < sum(reps)
    assert [b.shape[0] for v in ib for b in v] == g["infer_counts"].tolist()
    np.testing.assert_allclose(torch.cat([b for v in ib for b in v]).numpy(), g["infer_boxes"], atol=1e-3)
    np.testing.assert_allclose(torch.cat([l for v in il for l in v]).numpy(), g["infer_logits"], atol=2e-5)
