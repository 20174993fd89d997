"""Drop-in for ``vkit_open_model.loss_function`` (reference loss_function/__init__.py:12-24)."""
from .primitives import (  # noqa: F401
    WeightedBceWithLogitsLossFunction,
    CrossEntropyWithLogitsLossFunction,
    FocalWithLogitsLossFunction,
    L1LossFunction,
    L2LossFunction,
    WeightAdaptiveHeatmapRegressionLossFunction,
    DiceLossFunction,
)
from .adaptive_scaling import (  # noqa: F401
    Box,
    AdaptiveScalingRoughLossFunctionConifg,
    AdaptiveScalingRoughLossFunction,
    AdaptiveScalingPreciseLossFunctionConifg,
    AdaptiveScalingPreciseLossFunction,
)
