"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header declares, the module
mirror keeps the reference's parameter names (checkpoint key contract), and the product path refuses to run without CUDA."""
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import synth


def test_library_exports_every_declared_symbol():
    from grove_b200._lib import SIGNATURES, lib
    hdr = open(os.path.join(ROOT, "include", "grove_b200.h")).read()
    declared = set(re.findall(r"\b(grove_[a-z0-9_]+)\s*\(", hdr)) - {"grove_gemm_epilogue", "grove_twoway_a_params", "grove_twoway_b_params"}
    l = lib()
    for name in declared:
        assert hasattr(l, name), f"{name} is declared in include/grove_b200.h but not exported by libgrove_b200.so"
    assert declared == set(SIGNATURES), (declared ^ set(SIGNATURES))
    assert l.grove_abi_version() == 5
    # the test-only cross-check kernels live in their own header / library and are NOT part of the product library
    from grove_b200._lib import LEGACY_SIGNATURES, lib_legacy
    lhdr = open(os.path.join(ROOT, "include", "grove_b200_legacy.h")).read()
    ldecl = set(re.findall(r"\b(grove_[a-z0-9_]+)\s*\(", lhdr))
    assert ldecl == set(LEGACY_SIGNATURES) - {"grove_last_error"}
    ll = lib_legacy()
    for name in ldecl:
        assert hasattr(ll, name) and not hasattr(l, name), name


@pytest.mark.parametrize("vit", ["vit_b", "vit_l", "vit_h"])
def test_state_dict_keys_match_reference_names(vit):
    """oracle/make_golden.py asserts synth's tables equal the REFERENCE modules' state_dict(); here the mirror must carry them too."""
    from oracle.grounding import VIT_CFG
    from grove_b200.modeling.build_sam import sam_model_registry
    cfg = VIT_CFG[vit]
    with torch.device("meta"):
        sam = sam_model_registry[vit]()
    sd = sam.state_dict()
    want = {**synth.encoder_param_shapes(cfg["embed_dim"], cfg["depth"], cfg["heads"], cfg["global_idx"], 64),
            **synth.decoder_param_shapes()}
    for k, shp in want.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == tuple(shp), (k, tuple(sd[k].shape), shp)


def test_grounding_branch_surface_and_no_cpu_fallback():
    from grove_b200.modeling.grounding import GroundingBranch
    gb = GroundingBranch(vit="vit_b", image_size=512)
    for name in ("get_grounding_encoder_embs", "_create_det_token_mask", "_process_hidden_states", "_generate_and_postprocess_masks",
                 "_compute_loss_components_video"):
        assert callable(getattr(gb, name))
    assert [n for n, _ in gb.text_hidden_fcs[0].named_parameters()] == ["0.weight", "0.bias", "2.weight", "2.bias"]
    # train.py:170-191 style surgery: attributes the trainer reads / replaces exist
    enc, dec = gb.grounding_encoder.image_encoder, gb.grounding_encoder.mask_decoder
    a = enc.adapters[0].conv3d
    assert (a.in_channels, a.out_channels, tuple(a.kernel_size)) == (768, 768, (3, 3, 3))
    assert dec.bbox_prediction_head[0].in_features == 256 and dec.temporal_objectness_head.out_features == 1
    ids = torch.full((1, 65), 7)
    ids[0, 10] = gb.det_token_idx
    m = gb._create_det_token_mask(ids)
    assert m.shape == (1, 575 + 64 + 1) and int(m[0].nonzero()) == 575 + 9
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(1, 3, 8, 512, 512))


def test_checkpoint_tooling_matches_reference_surgery():
    """grove_b200.checkpoint vs train.py:503-576 executed from the reference's source (tests/golden/posembed.npz), and the
    strict=False key mapping of GROVE-prefixed checkpoints."""
    import numpy as np
    from conftest import GOLDEN
    from grove_b200 import checkpoint as ck
    from grove_b200.modeling.build_sam import sam_model_registry
    g = np.load(os.path.join(GOLDEN, "posembed.npz"))
    pos = synth.synth_tensor("posembed.pos", (1, 8, 8, 24), 9)
    rh, rw = synth.synth_tensor("posembed.rh", (15, 12), 9), synth.synth_tensor("posembed.rw", (15, 12), 9)
    for tgt in (64, 256):
        np.testing.assert_allclose(ck.resize_abs_pos_embedding(pos, tgt).numpy(), g[f"abs{tgt}"], rtol=0, atol=1e-6)
        h, w = ck.resize_rel_pos_embedding(rh, rw, tgt)
        np.testing.assert_allclose(h.numpy(), g[f"relh{tgt}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g[f"relw{tgt}"], rtol=0, atol=1e-6)
    sam = sam_model_registry["vit_b"]()                       # reference default: 1024 encoder tables
    enc = sam.image_encoder
    ck.interpolate_positional_embeddings(enc, 512)
    assert tuple(enc.pos_embed.shape) == (1, 32, 32, 768) and enc.img_size == 512
    assert tuple(enc.blocks[2].attn.rel_pos_h.shape) == (63, 64) and tuple(enc.blocks[0].attn.rel_pos_h.shape) == (27, 64)
    sd = {"model.grounding_encoder." + k: torch.zeros_like(v) for k, v in sam.state_dict().items()}
    sd["model.layers.0.self_attn.q_proj.weight"] = torch.zeros(2, 2)
    sd["model.text_hidden_fcs.0.0.bias"] = torch.ones(4096)
    fcs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(4096, 4096), torch.nn.ReLU(), torch.nn.Linear(4096, 256), torch.nn.Dropout(0.0))])
    r = ck.load_sam_state_dict(sam, sd, fcs)
    assert r["skipped"] == ["model.layers.0.self_attn.q_proj.weight"] and not r["unexpected"]
    assert float(enc.pos_embed.abs().sum()) == 0.0 and float(fcs[0][0].bias.sum()) == 4096.0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_reference_trainer_surgery_leaves_a_runnable_module():
    """train.py:162-191 (initialize_custom_layers_in_model) and :561-576 (interpolate_positional_embeddings), executed FROM THE REFERENCE'S
    SOURCE on a grove_b200 model: the trainer replaces `adapters` with instances of the reference's own SpatioTemporalConvAdapter class and
    the heads with fresh nn.Linear stacks, and reassigns pos_embed / rel_pos / img_size.  The encoder must accept all of it."""
    import ast
    import sys
    from types import SimpleNamespace as NS
    import torch.nn as nn
    from grove_b200.modeling.grounding import GroundingBranch
    from grove_b200.modeling.image_encoder import is_conv_adapter
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    try:
        from model.SAM.modeling.image_encoder import SpatioTemporalConvAdapter as RefAdapter
    finally:
        sys.path.remove("/root/reference")
    src = open("/root/reference/train.py").read()
    scope = {"torch": torch, "nn": nn, "F": torch.nn.functional, "GroundingSpatioTemporalConvAdapter": RefAdapter}
    want = {"initialize_custom_layers_in_model", "interpolate_positional_embeddings", "resize_abs_pos_embedding", "resize_rel_pos_embedding"}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "train.py", "exec"), scope)
    with torch.device("meta"):
        gb = GroundingBranch(vit="vit_h")                       # GROVE builds ViT-H at the SAM default 1024 (GROVE.py:55)
    enc, dec = gb.grounding_encoder.image_encoder, gb.grounding_encoder.mask_decoder
    scope["initialize_custom_layers_in_model"](NS(get_model=lambda: gb, config=gb.config))
    assert len(enc.adapters) == 4 and all(type(a) is RefAdapter and is_conv_adapter(a) for a in enc.adapters)
    assert isinstance(dec.bbox_prediction_head[2], nn.Linear) and dec.temporal_objectness_head.out_features == 1
    scope["interpolate_positional_embeddings"](NS(model=gb))
    assert enc.img_size == 512 and tuple(enc.pos_embed.shape) == (1, 32, 32, 1280)
    assert tuple(enc.blocks[7].attn.rel_pos_h.shape) == (63, 80) and tuple(enc.blocks[0].attn.rel_pos_h.shape) == (27, 80)
    # what train.py does next (:279-296): unfreeze by .parameters() of the replaced containers
    n_train = sum(p.numel() for m in (enc.adapters, dec, gb.text_hidden_fcs) for p in m.parameters())
    assert n_train > 4 * 27 * 1280 * 1280
    # the encoder's own structural checks accept the replaced modules (the CUDA run of the same surgery is tests/test_gpu_model.py)
    c3 = enc.adapters[0].conv3d
    assert tuple(c3.kernel_size) == (3, 3, 3) and c3.in_channels == c3.out_channels == 1280
