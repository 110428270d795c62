from .image_encoder import ImageEncoderViT  # noqa: F401
from .mask_decoder import MaskDecoder  # noqa: F401
from .prompt_encoder import PromptEncoder  # noqa: F401
from .transformer import TwoWayTransformer  # noqa: F401
