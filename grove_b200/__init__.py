"""grove_b200 — B200-native (sm_100a) implementation of GROVE's per-frame grounding path.

Host side: the reference's nn.Module API (`grove_b200.modeling`, mirroring model/SAM/modeling/*.py and the
grounding methods of model/GROVE.py) over a C-ABI CUDA library (`include/grove_b200.h`, `grove_b200/csrc`).
"""
from .modeling.build_sam import (build_sam_vit_b, build_sam_vit_h, build_sam_vit_l, sam_model_registry)  # noqa: F401
from .modeling.image_encoder import ImageEncoderViT, SpatioTemporalConvAdapter  # noqa: F401
from .modeling.mask_decoder import MaskDecoder  # noqa: F401
from .modeling.prompt_encoder import PromptEncoder  # noqa: F401
from .modeling.transformer import TwoWayTransformer  # noqa: F401
from .modeling.grounding import GroundingBranch  # noqa: F401
from .preprocess import ResizeLongestSide  # noqa: F401
