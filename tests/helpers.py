"""Shared builders: grove_b200 modules loaded with the deterministic synthetic (reference-named) state dict."""
import torch

from oracle import synth


def encoder_with_weights(vit: str, img: int, seed: int, device="cuda", dtype=torch.float32):
    from oracle.grounding import VIT_CFG
    from grove_b200.modeling.build_sam import sam_model_registry
    cfg = VIT_CFG[vit]
    sam = sam_model_registry[vit](None, True, image_size=img)
    shapes = {**synth.encoder_param_shapes(cfg["embed_dim"], cfg["depth"], cfg["heads"], cfg["global_idx"], img // 16),
              **synth.decoder_param_shapes()}
    sd = synth.synth_state_dict(shapes, seed)
    missing, unexpected = sam.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if m in shapes]
    return sam.to(device=device, dtype=dtype), sd, cfg
