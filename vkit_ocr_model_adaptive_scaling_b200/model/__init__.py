"""Drop-in for ``vkit_open_model.model`` (reference model/__init__.py:12-20)."""
from .convnext import ConvNext, ConvNextBlock, ConvNextBlockLayer  # noqa: F401
from .upernext import PpmBlock, UperNextNeck, UperNextHead  # noqa: F401
from .fpn import FpnNeck, FpnHead  # noqa: F401
from .adaptive_scaling import (  # noqa: F401
    AdaptiveScalingSize,
    AdaptiveScalingNeckHeadType,
    AdaptiveScalingConfig,
    AdaptiveScaling,
)
