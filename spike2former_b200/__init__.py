"""spike2former_b200 -- B200-native spiking hot path of Spike2Former.

Host side: the reference's registered classes (`Spiking_vit_MetaFormer`,
`MaskFormerHead`, `DCNTransformerEncoderPixelDecoder`) with the reference's
parameter names.  Device side: hand-written sm_100a kernels behind the C-ABI
declared in include/s2f.h (csrc/ -> libs2f.so).  No CPU fallback.
"""
from . import configs  # noqa: F401
from .registry import MODELS, ConfigDict  # noqa: F401
from .neuron import Q_IFNode  # noqa: F401
from .models import (  # noqa: F401
    DCNTransformerEncoderPixelDecoder,
    EncoderDecoder,
    MaskFormerHead,
    Spiking_vit_MetaFormer,
    build_segmentor,
)

__all__ = ["MODELS", "ConfigDict", "Spiking_vit_MetaFormer", "MaskFormerHead",
           "DCNTransformerEncoderPixelDecoder", "EncoderDecoder", "build_segmentor", "configs", "Q_IFNode"]
