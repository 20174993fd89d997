"""Primitive loss callables — drop-ins for the seven classes of ``vkit_open_model.loss_function``
(weighted_bce_with_logits.py, focal_with_logits.py, dice.py, l1.py, l2.py, weight_adaptive_heatmap_regression.py,
cross_entropy_with_logits.py): same constructor arguments and ``(pred, gt, mask=None)`` call signature, each returning
a 0-dim differentiable tensor.  Every one is a single device-side reduction (``csrc/pointwise_loss.cu``); gradients
flow to ``pred`` only (the reference's ground truths never require grad).

Deviation, stated: the reference's dice / hard-negative BCE multiply ``gt`` by ``mask`` *in place* (dice.py:30,
weighted_bce_with_logits.py:33-34) — a side effect on the caller's tensor; here ``gt`` is left untouched.  Only
epsilon values of 1e-6 (the reference default) are supported.
"""
from typing import Optional

import torch

from .. import ops


def _check_eps(eps: float) -> None:
    if abs(eps - 1E-6) > 1e-12:
        raise NotImplementedError('vkocr_b200 loss kernels are built for the reference default eps=1e-6')


class WeightedBceWithLogitsLossFunction:

    def __init__(self, negative_ratio: float = 3.0, eps: float = 1E-6):
        self.negative_ratio = negative_ratio
        self.eps = eps

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.HardNegativeBceFn.apply(pred, gt, mask, float(self.negative_ratio), float(self.eps))


class FocalWithLogitsLossFunction:

    def __init__(self, alpha: float = 0.25, gamma: float = 2, eps: float = 1E-6):
        _check_eps(eps)
        self.alpha = alpha
        self.gamma = gamma
        self.eps = eps

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None):
        return ops.PointwiseLossFn.apply(pred, gt, mask, ops.FOCAL, float(self.alpha), float(self.gamma))


class DiceLossFunction:

    def __init__(self, eps: float = 1E-6):
        _check_eps(eps)
        self.eps = eps

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None):
        return ops.PointwiseLossFn.apply(pred, gt, mask, ops.DICE, 0.0, 0.0)


class L1LossFunction:

    def __init__(self, eps: float = 1E-6, smooth: bool = False, smooth_beta: float = 1.0):
        _check_eps(eps)
        self.smooth = smooth
        self.smooth_beta = smooth_beta
        self.eps = eps

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None):
        if self.smooth and self.smooth_beta > 0:
            return ops.PointwiseLossFn.apply(pred, gt, mask, ops.SMOOTH_L1, float(self.smooth_beta), 0.0)
        return ops.PointwiseLossFn.apply(pred, gt, mask, ops.L1, 0.0, 0.0)


class L2LossFunction:

    def __init__(self, eps: float = 1E-6):
        _check_eps(eps)
        self.eps = eps

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None):
        return ops.PointwiseLossFn.apply(pred, gt, mask, ops.L2, 0.0, 0.0)


class WeightAdaptiveHeatmapRegressionLossFunction:

    def __init__(self, gamma: float = 0.01):
        self.gamma = gamma

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor):
        # NOTE (as in the reference): `pred` should already be a probability (sigmoid applied by the caller).
        return ops.PointwiseLossFn.apply(pred, gt, None, ops.WAHR, float(self.gamma), 0.0)


class CrossEntropyWithLogitsLossFunction:

    def __call__(self, pred: torch.Tensor, gt: torch.Tensor):
        return ops.SoftCrossEntropyFn.apply(pred, gt)
