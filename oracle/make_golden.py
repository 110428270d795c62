"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Nothing is copied from the reference: its modules are imported, fed the deterministic synthetic
weights/inputs of oracle/synth.py, and only their *outputs* are stored.  The script also asserts
that oracle/synth.py's name/shape tables equal the reference modules' state_dict() entries, which
pins the checkpoint key-name contract.
"""
from __future__ import annotations

import ast
import os
import sys
import types
from functools import partial
from types import SimpleNamespace as NS

import numpy as np
import torch

REF = os.environ.get("GROVE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

from oracle import synth  # noqa: E402


def _stub_mm():
    """mmdet/mmcv/mmengine are imported by model/layers.py:8-10 only (dead region encoder)."""
    for name, attrs in {"mmdet": {}, "mmdet.models": {"BaseRoIExtractor": object}, "mmcv": {}, "mmcv.cnn": {"ConvModule": object, "Linear": object},
                        "mmcv.ops": {"RoIAlign": object}, "mmengine": {}, "mmengine.model": {"normal_init": lambda *a, **k: None}}.items():
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)


def _functions_from_source(path, names, scope):
    """exec only the named top-level functions of a reference script (its module-level code loads
    BERT / CoreNLP and cannot be imported); nothing is written to the repo."""
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), scope)
    return scope


def _check_shapes(module, shapes, strip):
    sd = module.state_dict()
    for k, shp in shapes.items():
        rk = k[len(strip):]
        assert rk in sd, f"reference has no key {rk}"
        assert tuple(sd[rk].shape) == tuple(shp), (rk, tuple(sd[rk].shape), shp)


def _load(module, sdict, strip):
    missing, unexpected = module.load_state_dict({k[len(strip):]: v for k, v in sdict.items()}, strict=False)
    assert not unexpected, unexpected
    return missing


def encoder_case(name, *, embed_dim, depth, heads, global_idx, img, verbatim_adapter, seed):
    from model.SAM.modeling.image_encoder import ImageEncoderViT, SpatioTemporalConvAdapter
    from einops import rearrange
    G = img // 16
    enc = ImageEncoderViT(depth=depth, embed_dim=embed_dim, img_size=img, mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          num_heads=heads, patch_size=16, qkv_bias=True, use_rel_pos=True, global_attn_indexes=global_idx,
                          window_size=14, out_chans=256).eval()
    if not verbatim_adapter:
        # oracle-side generalisation of image_encoder.py:52 (h inferred instead of h=32); identical arithmetic
        class _Adapter(SpatioTemporalConvAdapter):
            def forward(self, x):
                x = rearrange(x, '(b t) h w c -> b c t h w', t=8)
                x = self.tanh(self.alpha) * self.relu(self.conv3d(x)) + x
                return rearrange(x, 'b c t h w -> (b t) h w c')
        enc.adapters = torch.nn.ModuleList([_Adapter(embed_dim, embed_dim, (3, 3, 3)) for _ in global_idx])
    shapes = synth.encoder_param_shapes(embed_dim, depth, heads, global_idx, G)
    _check_shapes(enc, shapes, "image_encoder.")
    assert set(k[len("image_encoder."):] for k in shapes) == set(enc.state_dict().keys())
    sdict = synth.synth_state_dict(shapes, seed)
    assert not _load(enc, sdict, "image_encoder.")
    images = synth.synth_tensor(name + ".images", (1, 3, 8, img, img), seed)
    with torch.no_grad():
        out = enc(images)
        out64 = enc.double()(images.double())   # the reference itself in float64: the tight pin
    # keep the fixture small: a strided sub-lattice (every 4th channel, every 2nd row/col) + per-channel sums
    np.savez_compressed(os.path.join(OUT, name + ".npz"), out_sub=out[:, ::4, ::2, ::2].numpy(),
                        out_chan_sum=out.double().sum((2, 3)).numpy(), out64_sub=out64[:, ::4, ::2, ::2].numpy(),
                        meta=np.array([embed_dim, depth, heads, img, seed] + list(global_idx)))
    print(name, tuple(out.shape), float(out.abs().mean()), "ref fp32-vs-fp64 gap %.2e" % float((out - out64).abs().max()))


def decoder_case(name, *, dim, mlp, G, frames, reps, seed):
    from model.SAM.modeling import MaskDecoder, PromptEncoder, TwoWayTransformer
    pe = PromptEncoder(embed_dim=dim, image_embedding_size=(G, G), input_image_size=(16 * G, 16 * G), mask_in_chans=16).eval()
    md = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=dim, mlp_dim=mlp, num_heads=8),
                     transformer_dim=dim, iou_head_depth=3, iou_head_hidden_dim=dim, decoding_type="query", use_temp_objectness=True).eval()
    shapes = synth.decoder_param_shapes(dim, mlp)
    _check_shapes(pe, {k: v for k, v in shapes.items() if k.startswith("prompt_encoder.")}, "prompt_encoder.")
    _check_shapes(md, {k: v for k, v in shapes.items() if k.startswith("mask_decoder.")}, "mask_decoder.")
    sdict = synth.synth_state_dict(shapes, seed)
    _load(pe, {k: v for k, v in sdict.items() if k.startswith("prompt_encoder.")}, "prompt_encoder.")
    _load(md, {k: v for k, v in sdict.items() if k.startswith("mask_decoder.")}, "mask_decoder.")
    emb = synth.synth_tensor(name + ".emb", (frames, dim, G, G), seed)
    txt = synth.synth_tensor(name + ".txt", (sum(reps), 1, dim), seed)
    with torch.no_grad():
        dense_pe = pe.get_dense_pe()
        sparse, dense = pe(points=None, boxes=None, masks=None, text_embeds=txt)
        boxes, logits = md(image_embeddings=emb, image_pe=dense_pe, sparse_prompt_embeddings=sparse,
                           dense_prompt_embeddings=dense, multimask_output=False, reps=list(reps))
        pe64, md64 = pe.double(), md.double()   # the reference itself in float64: the tight pin
        sp64, de64 = pe64(points=None, boxes=None, masks=None, text_embeds=txt.double())
        boxes64, logits64 = md64(image_embeddings=emb.double(), image_pe=pe64.get_dense_pe(), sparse_prompt_embeddings=sp64.double(),
                                 dense_prompt_embeddings=de64, multimask_output=False, reps=list(reps))
        pe.float(), md.float()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), boxes=boxes.numpy(), logits=logits.numpy(),
                        boxes64=boxes64.numpy(), logits64=logits64.numpy(),
                        dense_pe_sample=dense_pe[0, :, ::max(G // 4, 1), ::max(G // 4, 1)].numpy(),
                        meta=np.array([dim, mlp, G, frames, seed]), reps=np.array(reps))
    print(name, tuple(boxes.shape), "min|logit|", float(logits.abs().min()), "ref fp32-vs-fp64 gap: boxes %.2e logits %.2e"
          % (float((boxes - boxes64).abs().max()), float((logits - logits64).abs().max())))
    return pe, md, sdict


def glue_case(name, seed):
    """GROVE-level methods called unbound with a duck-typed self."""
    _stub_mm()
    torch.Tensor.cuda = lambda self, *a, **k: self  # GROVE.py:203,260 hard-code .cuda()
    import model.GROVE as RG
    C = RG.GROVEForCausalLM
    dim, mlp, G, T, hidden, L = 64, 128, 8, 8, 96, 600
    pe, md, sdict = decoder_case(name + "_dec", dim=dim, mlp=mlp, G=G, frames=2 * T, reps=[3] * T + [2] * T, seed=seed)
    fshapes = synth.text_fcs_shapes(hidden, dim)
    fsd = synth.synth_state_dict(fshapes, seed)
    fcs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(hidden, hidden), torch.nn.ReLU(inplace=True),
                                                   torch.nn.Linear(hidden, dim), torch.nn.Dropout(0.0))])
    _check_shapes(fcs, fshapes, "text_hidden_fcs.")
    _load(fcs, fsd, "text_hidden_fcs.")
    self = NS(model=NS(text_hidden_fcs=fcs, grounding_encoder=NS(prompt_encoder=pe, mask_decoder=md)),
              config=NS(num_frames=T, use_temp_objectness=True, temp_objectness_threshold=0.5),
              ce_loss_weight=1.0, giou_loss_weight=2.0, temp_objectness_loss_weight=2.0, det_token_idx=32005)
    ids = torch.full((2, L - 575), 7, dtype=torch.long)
    pos = [synth.det_positions(L, 3, seed), synth.det_positions(L, 2, seed + 1)]
    for v, pp in enumerate(pos):
        for p in pp:
            ids[v, p - 575 + 1] = 32005  # mask = ids[:,1:] shifted by the 575 pad
    hid = synth.synth_tensor(name + ".hidden", (2, L, hidden), seed)
    emb = synth.synth_tensor(name + "_dec.emb", (2 * T, dim, G, G), seed)
    with torch.no_grad():
        mask = C._create_det_token_mask(self, ids)
        assert [int(i) for i in mask[0].nonzero().flatten()] == pos[0]
        _, pred_list = C._process_hidden_states(self, [hid], mask, None)
        dense_pe = pe.get_dense_pe()
        tb, tl = C._generate_and_postprocess_masks(self, pred_list, emb, [(1280, 720), (640, 360)], dense_pe, infer=False)
        ib, il = C._generate_and_postprocess_masks(self, pred_list, emb, [(1280, 720), (640, 360)], dense_pe, infer=True)
        # ground truth for the loss: Bernoulli(.5) labels (float64 like HowTo100M.py:129), boxes for the positives
        rng = np.random.Generator(np.random.PCG64([seed, 99]))
        gt_b, gt_o = [], []
        for v in range(2):
            P = len(pos[v])
            gb, go = [], []
            for f in range(T):
                o = (rng.uniform(size=P) < 0.5).astype(np.float64)
                n = int(o.sum())
                cxcy = rng.uniform(0.3, 0.7, (n, 2))
                wh = rng.uniform(0.1, 0.4, (n, 2))
                gb.append(torch.from_numpy(np.concatenate([cxcy, wh], 1)).float())
                go.append(torch.from_numpy(o))
            gt_b.append(gb)
            gt_o.append(go)
        loss = C._compute_loss_components_video(self, tb, tl, gt_b, gt_o, NS(loss=torch.tensor(0.25)))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        det_mask=mask.numpy(), pred_embeddings=torch.cat(pred_list).numpy(), counts=np.array([p.shape[0] for p in pred_list]),
        train_boxes=torch.cat([b for v in tb for b in v]).numpy(), train_logits=torch.cat([l for v in tl for l in v]).numpy(),
        infer_boxes=torch.cat([b for v in ib for b in v]).numpy(), infer_counts=np.array([b.shape[0] for v in ib for b in v]),
        gt_boxes=torch.cat([b for v in gt_b for b in v]).numpy(), gt_obj=torch.cat([o for v in gt_o for o in v]).numpy(),
        losses=np.array([float(loss[k]) for k in ("loss", "ce_loss", "giou_loss", "l1_loss", "temp_objectness_loss")]))
    print(name, {k: float(v) for k, v in loss.items()})


def train_case(name, seed):
    """BASELINE config 4 in miniature, through the REFERENCE's own modules and GROVE methods in float64 with autograd: the pin for the
    oracle's gradients (the GPU training tests compare grove_b200's backward with autograd over the oracle)."""
    _stub_mm()
    torch.Tensor.cuda = lambda self, *a, **k: self
    import model.GROVE as RG
    from model.SAM.modeling import MaskDecoder, PromptEncoder, TwoWayTransformer
    from model.SAM.modeling.image_encoder import ImageEncoderViT
    C = RG.GROVEForCausalLM
    D, depth, heads, gidx, img, dim, mlp, T, hidden, L, P = 64, 3, 2, (1, 2), 512, 256, 128, 8, 96, 600, 2
    G = img // 16
    enc = ImageEncoderViT(depth=depth, embed_dim=D, img_size=img, mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=heads,
                          patch_size=16, qkv_bias=True, use_rel_pos=True, global_attn_indexes=gidx, window_size=14, out_chans=dim)
    pe = PromptEncoder(embed_dim=dim, image_embedding_size=(G, G), input_image_size=(img, img), mask_in_chans=16)
    md = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=dim, mlp_dim=mlp, num_heads=8),
                     transformer_dim=dim, iou_head_depth=3, iou_head_hidden_dim=dim, decoding_type="query", use_temp_objectness=True)
    fcs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(hidden, hidden), torch.nn.ReLU(inplace=True),
                                                   torch.nn.Linear(hidden, dim), torch.nn.Dropout(0.0))])
    esd = synth.synth_state_dict(synth.encoder_param_shapes(D, depth, heads, gidx, G), seed)
    dsd = synth.synth_state_dict(synth.decoder_param_shapes(dim, mlp), seed)
    fsd = synth.synth_state_dict(synth.text_fcs_shapes(hidden, dim), seed)
    assert not _load(enc, esd, "image_encoder.")
    _load(pe, {k: v for k, v in dsd.items() if k.startswith("prompt_encoder.")}, "prompt_encoder.")
    _load(md, {k: v for k, v in dsd.items() if k.startswith("mask_decoder.")}, "mask_decoder.")
    _load(fcs, fsd, "text_hidden_fcs.")
    for m in (enc, pe, md, fcs):
        m.double().train()
    self = NS(model=NS(text_hidden_fcs=fcs, grounding_encoder=NS(image_encoder=enc, prompt_encoder=pe, mask_decoder=md)),
              config=NS(num_frames=T, use_temp_objectness=True, temp_objectness_threshold=0.5),
              ce_loss_weight=1.0, giou_loss_weight=2.0, temp_objectness_loss_weight=2.0, det_token_idx=32005)
    ids = torch.full((1, L - 575), 7, dtype=torch.long)
    for p_ in synth.det_positions(L, P, seed):
        ids[0, p_ - 575 + 1] = 32005
    images = synth.synth_tensor(name + ".images", (1, 3, T, img, img), seed).double()
    hid = synth.synth_tensor(name + ".hidden", (1, L, hidden), seed).double().requires_grad_(True)
    rng = np.random.Generator(np.random.PCG64([seed, 77]))
    gt_b, gt_o = [[]], [[]]
    for f in range(T):
        o = (rng.uniform(size=P) < 0.5).astype(np.float64)
        if f == 0:
            o[0] = 1.0
        n = int(o.sum())
        gt_b[0].append(torch.from_numpy(np.concatenate([rng.uniform(0.3, 0.7, (n, 2)), rng.uniform(0.1, 0.4, (n, 2))], 1)))
        gt_o[0].append(torch.from_numpy(o))
    emb = C.get_grounding_encoder_embs(self, images)                       # GROVE.py:134-136 (no no_grad)
    mask = C._create_det_token_mask(self, ids)
    _, pred_list = C._process_hidden_states(self, [hid], mask, None)
    dense_pe = pe.get_dense_pe()
    tb, tl = C._generate_and_postprocess_masks(self, pred_list, emb, [(1280, 720)], dense_pe, infer=False)
    loss = C._compute_loss_components_video(self, tb, tl, gt_b, gt_o, NS(loss=torch.tensor(0.0, dtype=torch.float64)))
    loss["loss"].backward()
    out = {"losses": np.array([float(loss[k]) for k in ("loss", "ce_loss", "giou_loss", "l1_loss", "temp_objectness_loss")]),
           "gt_boxes": torch.cat(gt_b[0]).numpy(), "gt_obj": torch.cat(gt_o[0]).numpy(), "meta": np.array([D, depth, heads, img, dim, mlp, hidden, L, P, seed])}
    named = [("image_encoder." + k, v) for k, v in enc.named_parameters()] + [("mask_decoder." + k, v) for k, v in md.named_parameters()] + \
            [("text_hidden_fcs." + k, v) for k, v in fcs.named_parameters()] + [("hidden", hid)]
    keys = []
    for k, v in named:
        if v.grad is None:
            continue
        g = v.grad.detach().reshape(-1)
        step = max(g.numel() // 64, 1)
        out["gn:" + k] = np.array(float(g.norm()))
        out["gs:" + k] = g[::step][:64].numpy()
        keys.append(k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: float(v) for k, v in loss.items()}, len(keys), "gradient tensors")


def preprocess_case(name, seed):
    """ResizeLongestSide.apply_image (PIL bilinear) + grounding_enc_processor + .bfloat16() through the reference's own code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(REF, "model", "SAM", "utils", "transforms.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    gp = _functions_from_source(os.path.join(REF, "infer_iground.py"), {"grounding_enc_processor"}, {"torch": torch, "F": torch.nn.functional})
    rng = np.random.Generator(np.random.PCG64([seed, 3]))
    out = {}
    cases = [("down", 2, 90, 160, 64), ("up", 2, 20, 48, 64), ("tall", 1, 120, 45, 96), ("same", 1, 64, 64, 64), ("odd", 1, 77, 131, 80)]
    for tag, T, h, w, L in cases:
        yy, xx = np.mgrid[0:h, 0:w]
        base = (127 + 90 * np.sin(yy / 7.0)[..., None] * np.cos(xx / 5.0)[..., None] + rng.normal(0, 25, (T, h, w, 3))).clip(0, 255)
        frames = base.astype(np.uint8)
        resized = np.stack([tr.ResizeLongestSide(L).apply_image(f) for f in frames])
        out[f"{tag}.frames"] = frames
        out[f"{tag}.resized"] = resized
        out[f"{tag}.meta"] = np.array([T, h, w, L])
    # the normalise + pad + bf16 half has IMG_SIZE = 512 hard-coded (infer_iground.py:307): a small 40x64 clip, stored as bf16 bit patterns
    frames = rng.integers(0, 256, (2, 40, 64, 3), dtype=np.uint8)
    x = gp["grounding_enc_processor"](torch.from_numpy(frames).permute(3, 0, 1, 2).contiguous()).bfloat16()
    out["proc.frames"] = frames
    out["proc.bits_sub"] = x.view(torch.int16)[:, :, :48:1, :72:1].numpy()          # the image plus a rim of the zero padding
    out["proc.pad_nonzero"] = np.array(int((x[:, :, 40:, :] != 0).sum() + (x[:, :, :, 64:] != 0).sum()))
    out["proc.shape"] = np.array(x.shape)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("resized")})


def posembed_case(name, seed):
    """train.py:503-558 (resize_abs_pos_embedding / resize_rel_pos_embedding) executed from the reference's source"""
    fn = _functions_from_source(os.path.join(REF, "train.py"), {"resize_abs_pos_embedding", "resize_rel_pos_embedding"},
                                {"torch": torch, "F": torch.nn.functional, "nn": torch.nn})
    pos = synth.synth_tensor(name + ".pos", (1, 8, 8, 24), seed)
    rh, rw = synth.synth_tensor(name + ".rh", (15, 12), seed), synth.synth_tensor(name + ".rw", (15, 12), seed)
    out = {}
    for tgt in (64, 256):     # 8x8 -> 4x4 (down) and 16x16 (up)
        out[f"abs{tgt}"] = fn["resize_abs_pos_embedding"](pos, tgt, 16, 24).numpy()
        h, w = fn["resize_rel_pos_embedding"](rh, rw, tgt, 16, 12)
        out[f"relh{tgt}"], out[f"relw{tgt}"] = h.numpy(), w.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def box_eval_case(name, seed):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval_vidstg", os.path.join(REF, "eval_vidstg.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    ig = _functions_from_source(os.path.join(REF, "eval_iground.py"), {"compute_iou", "compute_iou_matrix"}, {"np": np})
    an = _functions_from_source(os.path.join(REF, "eval_anet.py"), {"bbox_overlaps_batch"}, {"np": np, "torch": torch})
    sl = _functions_from_source(os.path.join(REF, "infer_iground.py"), {"sliding_segment_with_mask"}, {})
    rng = np.random.Generator(np.random.PCG64([seed, 5]))

    def boxes(n, lo=0, hi=200, integer=False):
        xy = rng.uniform(lo, hi, (n, 2))
        wh = rng.uniform(0, 80, (n, 2))
        b = np.concatenate([xy, xy + wh], 1)
        return np.round(b) if integer else b
    b1, b2 = boxes(13), boxes(9)
    b1[3] = b2[2]                      # exact overlap
    b1[4] = [5, 5, 5, 5]               # zero-area
    b2[5] = [300, 300, 310, 310]       # disjoint
    iou64 = ev.np_box_iou(b1, b2)
    iou32 = ev.np_box_iou(b1.astype(np.float32), b2.astype(np.float32))
    p1, p2 = boxes(7, 0, 60, integer=True), boxes(6, 0, 60, integer=True)
    p1[2] = p2[1]
    mat = ig["compute_iou_matrix"](p1.tolist(), p2.tolist())
    # greedy matcher: the reference loop (eval_iground.py:85-96) re-executed on the reference's matrix
    sims = rng.uniform(0, 1, mat.shape)
    ious, ts, matches = mat.copy(), sims.copy(), []
    while ious.size > 0 and ts.size > 0:
        m = np.unravel_index(np.argmax(ious), ious.shape)
        if ious[m] < 0.3 or ts[m] < 0.2:
            break
        matches.append(m)
        ious[m[0], :] = 0; ious[:, m[1]] = 0; ts[m[0], :] = 0; ts[:, m[1]] = 0
    anc = np.concatenate([boxes(10, integer=True), np.arange(10)[:, None]], 1)[None].astype(np.float32)
    gtb = np.concatenate([boxes(4, integer=True), rng.integers(0, 10, (4, 1))], 1)[None].astype(np.float32)
    anc[0, 1, :4] = [7, 7, 7, 7]       # zero-area anchor -> -1
    gtb[0, 2, :4] = [9, 9, 9, 9]       # zero-area gt -> 0
    frm = (anc[0, :, 4][:, None] != gtb[0, :, 4][None, :]).astype(np.uint8)[None]
    ov = an["bbox_overlaps_batch"](torch.from_numpy(anc), torch.from_numpy(gtb), torch.from_numpy(frm)).numpy()
    ov_nomask = an["bbox_overlaps_batch"](torch.from_numpy(anc), torch.from_numpy(gtb), torch.from_numpy(np.zeros_like(frm))).numpy()
    seg = {str(n): sl["sliding_segment_with_mask"](n, 8) for n in (8, 48, 50, 61, 128)}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), b1=b1, b2=b2, iou64=iou64, iou32=iou32, p1=p1, p2=p2, mat=mat, sims=sims,
                        matches=np.array(matches, dtype=np.int64).reshape(-1, 2), anc=anc, gtb=gtb, frm=frm, ov=ov, ov_nomask=ov_nomask,
                        **{f"seg_idx_{k}": np.array(sum(v[0], [])) for k, v in seg.items()},
                        **{f"seg_mask_{k}": np.array(sum(v[1], [])) for k, v in seg.items()},
                        **{f"seg_len_{k}": np.array([len(r) for r in v[0]]) for k, v in seg.items()})
    print(name, "iou", iou64.shape, "matches", matches)



def decisions_case(name, seed):
    """The three decision utilities the round-1 product lacked, each executed from the reference's own source:
    centre-in-box accuracy (eval_youcookinteractions.py:8-51), video IoU + strict '>' recall flags (VidSTGiouEvaluator.evaluate,
    eval_vidstg.py:118-186) and the validation GIoU / objectness-accuracy sums (the `if logits_temp_objectness is not None:` block of
    validate_model_performance, train.py:821-840 -- the enclosing function cannot run (it references an undefined name), so the block
    is located in the AST and executed with its free variables bound)."""
    import importlib.util
    from torchvision.ops import generalized_box_iou_loss
    rng = np.random.Generator(np.random.PCG64([seed, 11]))
    out = {}
    # ---- centre-in-box
    spec = importlib.util.spec_from_file_location("ref_eval_yc", os.path.join(REF, "eval_youcookinteractions.py"))
    yc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(yc)
    n = 24
    gt = np.round(rng.uniform(0, 100, (n, 2)))
    gt = np.concatenate([gt, gt + np.round(rng.uniform(5, 60, (n, 2)))], 1)
    pred = gt + rng.uniform(-30, 30, (n, 4))
    pred[0] = [gt[0, 0], gt[0, 1], gt[0, 0], gt[0, 1]]                     # centre exactly on the top-left corner: inclusive -> correct
    pred[1] = [gt[1, 2] - 4, gt[1, 3] - 2, gt[1, 2] + 4, gt[1, 3] + 2]     # centre exactly on the bottom-right corner -> correct
    pred[2] = [gt[2, 2], 0, gt[2, 2] + 2, 2 * gt[2, 1]]                    # centre one pixel right of xbr -> wrong
    pred[3] = np.nan                                                        # NaN prediction: valid, not correct
    kinds = np.zeros(n, dtype=np.int64)                                     # 0 normal, 1 empty gt (skipped), 2 pred None, 3 NaN pred
    kinds[3] = 3; kinds[5] = 1; kinds[9] = 2; kinds[17] = 1
    gt_data, pred_dict = [], {}
    for c in range(3):                                                      # three clips of eight frames
        sl = slice(8 * c, 8 * c + 8)
        gt_boxes = [[] if k == 1 else tuple(float(v) for v in g) for g, k in zip(gt[sl], kinds[sl])]
        preds = [None if k == 2 else np.array([p]) for p, k in zip(pred[sl], kinds[sl])]
        gt_data.append({"video_id": f"v{c}", "segment_youcook_idx": c, "segment_bboxes": gt_boxes})
        pred_dict[f"v{c}_{c}"] = {"final_boxes": preds}
    acc, correct, valid = yc.evaluate_dataset_localization(pred_dict, gt_data, "youcook")
    flags = []
    for i in range(n):                                                      # per-pair decisions through the same function
        if kinds[i] in (1, 2):
            flags.append(0)
            continue
        _, c1, _ = yc.evaluate_dataset_localization({"a_0": {"final_boxes": [np.array([pred[i]])]}},
                                                    [{"video_id": "a", "segment_youcook_idx": 0, "segment_bboxes": [tuple(float(v) for v in gt[i])]}], "youcook")
        flags.append(c1)
    out.update(cib_pred=pred, cib_gt=gt, cib_kinds=kinds, cib_flags=np.array(flags, dtype=np.uint8), cib_result=np.array([acc, correct, valid], dtype=np.float64))
    # ---- video IoU
    spec = importlib.util.spec_from_file_location("ref_eval_vidstg2", os.path.join(REF, "eval_vidstg.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    E = object.__new__(ev.VidSTGiouEvaluator)
    E.iou_thresholds = [0.3, 0.5]
    nf = 10
    E.video_gt, preds = {}, {}
    vg, vp = [], []
    for v in range(5):
        g = rng.uniform(0, 100, (nf, 2)); g = np.concatenate([g, g + rng.uniform(10, 80, (nf, 2))], 1)
        jitter = [3, 12, 25, 60, 0][v]
        p_ = g + rng.uniform(-jitter, jitter, (nf, 4)) if jitter else g.copy()
        if v == 4:                                                          # every frame IoU exactly 0.5 -> vIoU == 0.5 -> '>' 0.5 is False
            g = np.tile(np.array([0.0, 0.0, 2.0, 1.0]), (nf, 1)); p_ = np.tile(np.array([0.0, 0.0, 1.0, 1.0]), (nf, 1))
        p_[2] = 0.0                                                         # all-zero prediction: IoU 0 without calling np_box_iou
        if v == 4:
            p_[2] = [0.0, 0.0, 1.0, 1.0]
        vg.append(g); vp.append(p_)
        E.video_gt[f"vid{v}"] = {"frame_ids": list(range(0, 2 * nf, 2)), "boxes": [list(map(float, r)) for r in g]}
        order = list(rng.permutation(nf))                                   # predictions arrive in a different frame order
        preds[f"vid{v}"] = {"qtype": "declarative", "frame_ids": [2 * int(i) for i in order], "boxes": [np.array([p_[int(i)]]) for i in order]}
    vm = E.evaluate(preds)
    out.update(viou_gt=np.stack(vg), viou_pred=np.stack(vp),
               viou_value=np.array([vm[f"vid{v}"]["gt_viou"] for v in range(5)]),
               viou_over=np.array([[vm[f"vid{v}"][f"gt_viou@{t}"] for t in E.iou_thresholds] for v in range(5)], dtype=np.uint8),
               viou_frame=np.array([[vm[f"vid{v}"]["img_metrics"][f"vid{v}_{2 * f}"]["iou"] for f in range(nf)] for v in range(5)]))
    # ---- validation sums: locate the block in train.py's AST and execute it
    tree = ast.parse(open(os.path.join(REF, "train.py")).read())
    fn = next(nd for nd in tree.body if isinstance(nd, ast.FunctionDef) and nd.name == "validate_model_performance")
    block = next(nd for nd in ast.walk(fn) if isinstance(nd, ast.If) and isinstance(nd.test, ast.Compare)
                 and isinstance(nd.test.left, ast.Name) and nd.test.left.id == "logits_temp_objectness")
    code = compile(ast.Module([block], []), "train.py", "exec")
    T, P = 8, 3
    pb = torch.from_numpy(rng.uniform(0.05, 0.95, (2, T, P, 4))).float()
    lg = torch.from_numpy(rng.normal(0, 2, (2, T, P))).float()
    lg[0, 0, 0] = 0.0                                                       # sigmoid(0) = 0.5 is NOT '> 0.5'
    go = torch.from_numpy((rng.uniform(size=(2, T, P)) < 0.5).astype(np.int32))
    gtb = torch.from_numpy(rng.uniform(0.05, 0.95, (2, T, P, 4))).float()
    for variant, gcast in (("f", lambda t: t), ("i", lambda t: t.int())):   # raw float ground truth, and the reference's own `.int()` cast
        scope = {"torch": torch, "F": torch.nn.functional, "generalized_box_iou_loss": generalized_box_iou_loss,
                 "pred_bboxes": [[pb[v, f] for f in range(T)] for v in range(2)],
                 "logits_temp_objectness": [[lg[v, f] for f in range(T)] for v in range(2)],
                 "gt_bboxes": [[gcast(gtb[v, f][go[v, f].bool()]) for f in range(T)] for v in range(2)],
                 "gt_temp_objectness": [[go[v, f] for f in range(T)] for v in range(2)]}
        exec(code, scope)
        out[f"val_{variant}"] = np.array([float(scope["giou_sum"]), float(scope["temp_objectness_sum"]), scope["num_bboxes"], scope["num_max_bboxes"]],
                                         dtype=np.float64)
    out.update(val_boxes=pb.numpy(), val_logits=lg.numpy(), val_obj=go.numpy(), val_gt=gtb.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "cib", out["cib_result"], "viou", out["viou_value"], out["viou_over"].tolist(), "val", out["val_f"], out["val_i"])


def infer_case(name, seed):
    """GROVEForCausalLM._generate_and_postprocess_masks(infer=True) (GROVE.py:297-331) at the real decoder width (256 / 2048, 16x16 grid) so the
    CUDA decoder can run the same case: un-normalise with orig_sizes, cxcywh -> xyxy, keep sigmoid(logit) > threshold.  The threshold is set to the
    median objectness of the case so that the keep decisions straddle it (the stock 0.5 keeps nothing for these random weights)."""
    _stub_mm()
    torch.Tensor.cuda = lambda self, *a, **k: self
    import model.GROVE as RG
    C = RG.GROVEForCausalLM
    dim, mlp, G, T = 256, 2048, 16, 8
    reps = [3] * T + [2] * T
    pe, md, sdict = decoder_case(name + "_dec", dim=dim, mlp=mlp, G=G, frames=2 * T, reps=reps, seed=seed)
    os.remove(os.path.join(OUT, name + "_dec.npz"))
    emb = synth.synth_tensor(name + "_dec.emb", (2 * T, dim, G, G), seed)
    txt = synth.synth_tensor(name + "_dec.txt", (sum(reps), 1, dim), seed)
    pred_list, s0 = [], 0
    for r in reps:
        pred_list.append(txt[s0:s0 + r, 0]); s0 += r
    sizes = [(1280, 720), (640, 360)]
    self = NS(model=NS(grounding_encoder=NS(prompt_encoder=pe, mask_decoder=md)),
              config=NS(num_frames=T, use_temp_objectness=True, temp_objectness_threshold=0.5))
    with torch.no_grad():
        tb, tl = C._generate_and_postprocess_masks(self, pred_list, emb, sizes, pe.get_dense_pe(), infer=False)
        logits = torch.cat([l for v in tl for l in v])
        sg = torch.sigmoid(logits).sort().values
        thr = float((sg[sg.numel() // 2 - 1] + sg[sg.numel() // 2]) / 2)      # midway between the two middle objectness values
        self.config.temp_objectness_threshold = thr
        ib, il = C._generate_and_postprocess_masks(self, pred_list, emb, sizes, pe.get_dense_pe(), infer=True)
    counts = np.array([b.shape[0] for v in ib for b in v])
    assert 0 < counts.sum() < sum(reps), counts
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array([dim, mlp, G, T, seed]), reps=np.array(reps), thr=np.array(thr),
                        sizes=np.array(sizes), train_boxes=torch.cat([b for v in tb for b in v]).numpy(), train_logits=logits.numpy(),
                        infer_boxes=torch.cat([b for v in ib for b in v]).numpy(), infer_counts=counts,
                        infer_logits=torch.cat([l for v in il for l in v]).numpy())
    print(name, "threshold %.4f keeps %d of %d boxes" % (thr, counts.sum(), sum(reps)), "margin", float((torch.sigmoid(logits) - thr).abs().min()))


def _classes_from_source(path, names, scope):
    """exec only the named top-level classes of a reference module (modeling_clip.py imports a transformers version that is not
    installed here); nothing is written to the repo."""
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), scope)
    return scope


def clip_case(name, seed):
    """the CLIP-side SpatioTemporalConvAdapter (modeling_clip.py:591-612) and AdaptiveAvgPooling3D (pooling.py:6-25),
    executed from the reference's own source in fp32 and fp64."""
    import importlib.util
    from einops import rearrange
    from oracle import synth
    enc_dir = os.path.join(REF, "model", "llava", "model", "multimodal_encoder")
    scope = _classes_from_source(os.path.join(enc_dir, "modeling_clip.py"), {"SpatioTemporalConvAdapter"},
                                 {"nn": torch.nn, "torch": torch, "Conv3d": torch.nn.Conv3d, "rearrange": rearrange})
    spec = importlib.util.spec_from_file_location("ref_pooling", os.path.join(enc_dir, "pooling.py"))
    pooling = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pooling)
    out = {}
    C, b = 32, 1
    x = synth.synth_tensor(name + ".adapter.x", (b * 8, 257, C), seed)
    w = synth.synth_tensor(name + ".adapter.w", (C, C, 3, 3, 3), seed) * (27 * C) ** -0.5
    bias = synth.synth_tensor(name + ".adapter.b", (C,), seed) * 0.1
    for dt, tag in ((torch.float64, "64"),):                                 # fp64 only: the fixture stays small
        ad = scope["SpatioTemporalConvAdapter"](C, C, (3, 3, 3)).to(dt)
        with torch.no_grad():
            ad.conv3d.weight.copy_(w.to(dt)); ad.conv3d.bias.copy_(bias.to(dt)); ad.alpha.fill_(0.5)
            out["adapter" + tag] = ad((x.to(dt),))[0].numpy()
        for pname, shape in (("pool_a", (1 * 8, 256, 16)), ("pool_b", (1 * 8, 576, 8))):
            xp = synth.synth_tensor(f"{name}.{pname}.x", shape, seed).to(dt)
            with torch.no_grad():
                out[pname + tag] = pooling.AdaptiveAvgPooling3D(num_frames=8)(xp).numpy()
    out["meta"] = np.array([C, b, seed])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    torch.manual_seed(0)
    # 512^2 = the reference's real operating point, reference adapter class VERBATIM (t=8,h=32 hard-coded)
    encoder_case("enc_tiny512_verbatim", embed_dim=64, depth=3, heads=2, global_idx=(1, 2), img=512, verbatim_adapter=True, seed=1)
    # 256^2 (G=16 -> windows padded 16->28) with the oracle-side generalised adapter
    encoder_case("enc_tiny256_padded", embed_dim=96, depth=2, heads=3, global_idx=(1,), img=256, verbatim_adapter=False, seed=2)
    # BASELINE config 1 at full size (64x64x256, 8 frames x 4 phrases)
    decoder_case("dec_cfg1_full", dim=256, mlp=2048, G=64, frames=8, reps=[4] * 8, seed=3)
    # ragged phrases incl. a frame with zero phrases
    decoder_case("dec_ragged", dim=256, mlp=2048, G=32, frames=4, reps=[2, 0, 3, 1], seed=4)
    glue_case("glue", seed=5)
    box_eval_case("box_eval", seed=6)
    train_case("train_tiny512", seed=7)
    preprocess_case("preprocess", seed=8)
    posembed_case("posembed", seed=9)
    clip_case("clip_adapters", seed=10)
    decisions_case("decisions", seed=12)
    infer_case("glue_infer", seed=13)


if __name__ == "__main__":
    main()
