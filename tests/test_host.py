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
    declared = set(re.findall(r"\b(grove_[a-z0-9_]+)\s*\(", hdr)) - {"grove_gemm_epilogue"}
    l = lib()
    for name in declared:
        assert hasattr(l, name), f"{name} is declared in include/grove_b200.h but not exported by libgrove_b200.so"
    assert declared == set(SIGNATURES), (declared ^ set(SIGNATURES))
    assert l.grove_abi_version() == 2


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
