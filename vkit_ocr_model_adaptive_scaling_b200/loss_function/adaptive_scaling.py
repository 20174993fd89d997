"""Rough / precise adaptive-scaling losses on the B200 kernels — drop-ins for
``vkit_open_model.loss_function.adaptive_scaling`` (reference loss_function/adaptive_scaling.py:27-346): same config
classes (including the reference's ``Conifg`` spelling), constructors and keyword call signatures; each call returns a
0-dim differentiable fp32 tensor that stays on the device.

The terms that are active in the reference's default configuration run as ONE fused reduction per loss
(``csrc/loss.cu``): the prediction maps are read once inside the core box, the label points are gathered in the
kernel, and the backward writes the gradient maps directly.  Terms that are off by default (hard-negative BCE,
masked focal of the optional char-mask head, prob smooth-L1, WAHR) are added from the primitive-loss kernels.
"""
from typing import Optional, Tuple

import attrs
import torch

from .. import ops
from .primitives import (
    CrossEntropyWithLogitsLossFunction,
    DiceLossFunction,
    FocalWithLogitsLossFunction,
    L1LossFunction,
    L2LossFunction,
    WeightAdaptiveHeatmapRegressionLossFunction,
    WeightedBceWithLogitsLossFunction,
)


@attrs.define
class Box:
    """Inclusive pixel box.  The losses accept any object with ``up/down/left/right`` (the reference passes
    ``vkit.element.Box``; loss_function/adaptive_scaling.py:15,75-86)."""
    up: int
    down: int
    left: int
    right: int


def _crop(feature: torch.Tensor, box) -> torch.Tensor:
    """(B,1,H,W) -> (B,CH,CW) view of the core box."""
    return feature[:, 0, box.up:box.down + 1, box.left:box.right + 1]


def _require(cond: bool, what: str) -> None:
    if not cond:
        raise NotImplementedError(
            'vkocr_b200 fused loss kernels are built for the reference\'s hyper-parameters; ' + what +
            ' (changing the primitive objects on the composite loss has no effect on the fused path)')


def _check_box(box, shape: Tuple[int, int], gt: torch.Tensor) -> Tuple[int, int]:
    ch, cw = box.down - box.up + 1, box.right - box.left + 1
    assert 0 <= box.up and box.down < shape[0] and 0 <= box.left and box.right < shape[1], 'core box outside the map'
    assert tuple(gt.shape[1:]) == (ch, cw), f'ground truth {tuple(gt.shape)} does not match the core box ({ch}, {cw})'
    return ch, cw


@attrs.define
class AdaptiveScalingRoughLossFunctionConifg:
    bce_negative_ratio: float = 3.0
    bce_factor: float = 0.0
    focal_factor: float = 5.0
    dice_factor: float = 1.0
    l1_factor: float = 1.0
    downsampled_score_map_min: float = 1.1
    char_height_feature_min: float = 1.1


class AdaptiveScalingRoughLossFunction:

    def __init__(self, config: AdaptiveScalingRoughLossFunctionConifg):
        self.config = config
        self.weighted_bce_with_logits = WeightedBceWithLogitsLossFunction(negative_ratio=config.bce_negative_ratio)
        self.focal_with_logits = FocalWithLogitsLossFunction()
        self.dice = DiceLossFunction()
        self.l1 = L1LossFunction(smooth=True)

    def __call__(
        self,
        rough_char_mask_feature: torch.Tensor,    # (B, 1, H, W)
        rough_char_height_feature: torch.Tensor,  # (B, 1, H, W)
        downsampled_mask: torch.Tensor,           # (B, CH, CW)
        downsampled_score_map: torch.Tensor,      # (B, CH, CW)
        downsampled_shape: Tuple[int, int],
        downsampled_core_box,
    ) -> torch.Tensor:
        assert rough_char_mask_feature.shape == rough_char_height_feature.shape
        assert tuple(rough_char_mask_feature.shape[1:]) == (1, *downsampled_shape)
        box = downsampled_core_box
        _check_box(box, tuple(downsampled_shape), downsampled_mask)
        cfg = self.config
        assert downsampled_mask.shape == downsampled_score_map.shape
        assert downsampled_mask.shape[0] == rough_char_mask_feature.shape[0], 'ground-truth batch differs from the prediction batch'
        # the fused kernel (csrc/loss.cu) compiles in the hyper-parameters of the reference's primitives (:42-51)
        fo, l1 = self.focal_with_logits, self.l1
        _require(abs(fo.alpha - 0.25) < 1e-12 and abs(fo.gamma - 2) < 1e-12, f'focal alpha {fo.alpha} gamma {fo.gamma} != 0.25 / 2')
        _require(l1.smooth and abs(l1.smooth_beta - 1.0) < 1e-12, f'rough smooth-L1 beta {l1.smooth_beta} != 1.0')
        loss = ops.RoughLossFn.apply(
            rough_char_mask_feature, rough_char_height_feature, downsampled_mask, downsampled_score_map,
            int(box.up), int(box.left), float(cfg.char_height_feature_min), float(cfg.downsampled_score_map_min),
            float(cfg.focal_factor), float(cfg.dice_factor), float(cfg.l1_factor))
        if cfg.bce_factor > 0.0:
            loss = loss + cfg.bce_factor * self.weighted_bce_with_logits(
                pred=_crop(rough_char_mask_feature, box), gt=downsampled_mask)
        return loss


@attrs.define
class AdaptiveScalingPreciseLossFunctionConifg:
    char_mask_focal_factor: float = 0.0
    char_prob_l1_factor: float = 0.0
    char_prob_pos_l2_factor: float = 2.0
    char_prob_neg_l2_factor: float = 1.0
    char_prob_wahr_factor: float = 0.0
    char_up_left_offset_l1_factor: float = 1.0
    char_up_left_distance_regulation_l1_factor: float = 1.0
    char_corner_angle_cross_entropy_factor: float = 5.0
    char_corner_distance_l1_factor: float = 1.0
    loss_factor: float = 0.15


class AdaptiveScalingPreciseLossFunction:

    LABEL_POINT_SMOOTH_BETA = 2.5  # reference :160-165

    def __init__(self, config: AdaptiveScalingPreciseLossFunctionConifg):
        self.config = config
        self.char_mask_focal_with_logits = FocalWithLogitsLossFunction()
        self.char_prob_l1 = L1LossFunction(smooth=True, smooth_beta=0.25)
        self.char_prob_l2 = L2LossFunction()
        self.char_prob_wahr = WeightAdaptiveHeatmapRegressionLossFunction()
        self.char_up_left_offset_l1 = L1LossFunction(smooth=True, smooth_beta=self.LABEL_POINT_SMOOTH_BETA)
        self.char_up_left_distance_regulation_l1 = L1LossFunction(smooth=True, smooth_beta=self.LABEL_POINT_SMOOTH_BETA)
        self.char_corner_angle_cross_entropy = CrossEntropyWithLogitsLossFunction()
        self.char_corner_distance_l1 = L1LossFunction(smooth=True, smooth_beta=self.LABEL_POINT_SMOOTH_BETA)

    @classmethod
    def get_label_point_feature(cls, feature: torch.Tensor, label_point_y: torch.Tensor, label_point_x: torch.Tensor):
        """(B,C,H,W) gathered at (B,P) points -> (B,P,C) (reference :167-179).  A host-side convenience for callers; the
        loss itself gathers inside the fused kernel."""
        batch_size = feature.shape[0]
        assert batch_size == label_point_y.shape[0] == label_point_x.shape[0]
        return feature[torch.arange(batch_size, device=feature.device)[:, None], :, label_point_y, label_point_x]

    def __call__(
        self,
        precise_char_mask_feature: Optional[torch.Tensor],             # (B, 1, H, W) or None
        precise_char_prob_feature: torch.Tensor,                       # (B, 1, H, W)
        precise_char_up_left_corner_offset_feature: torch.Tensor,      # (B, 2, H, W)
        precise_char_corner_angle_feature: torch.Tensor,               # (B, 4, H, W)
        precise_char_corner_distance_feature: torch.Tensor,            # (B, 4, H, W)
        downsampled_char_prob_score_map: torch.Tensor,                 # (B, CH, CW)
        downsampled_char_mask: torch.Tensor,                           # (B, CH, CW)
        downsampled_shape: Tuple[int, int],
        downsampled_core_box,
        downsampled_label_point_y: torch.Tensor,                       # (B, P) int64
        downsampled_label_point_x: torch.Tensor,                       # (B, P) int64
        char_up_left_offsets: torch.Tensor,                            # (B, P, 2)
        char_corner_angles: torch.Tensor,                              # (B, P, 4)
        char_corner_distances: torch.Tensor,                           # (B, P, 3)
    ) -> torch.Tensor:
        if precise_char_mask_feature is not None:
            assert precise_char_mask_feature.shape == precise_char_prob_feature.shape
        assert tuple(precise_char_prob_feature.shape[1:]) == (1, *downsampled_shape)
        box = downsampled_core_box
        _check_box(box, tuple(downsampled_shape), downsampled_char_mask)
        cfg = self.config
        assert precise_char_up_left_corner_offset_feature.shape[1] == 2
        assert precise_char_corner_angle_feature.shape[1] == 4 and precise_char_corner_distance_feature.shape[1] == 4

        for prim in (self.char_up_left_offset_l1, self.char_up_left_distance_regulation_l1, self.char_corner_distance_l1):
            _require(prim.smooth and abs(prim.smooth_beta - self.LABEL_POINT_SMOOTH_BETA) < 1e-12,
                     f'label-point smooth-L1 beta {prim.smooth_beta} != {self.LABEL_POINT_SMOOTH_BETA}')
        factors = (cfg.char_prob_pos_l2_factor, cfg.char_prob_neg_l2_factor, cfg.char_up_left_offset_l1_factor,
                   cfg.char_up_left_distance_regulation_l1_factor, cfg.char_corner_angle_cross_entropy_factor,
                   cfg.char_corner_distance_l1_factor, cfg.loss_factor)
        loss = ops.PreciseLossFn.apply(
            precise_char_prob_feature, precise_char_up_left_corner_offset_feature, precise_char_corner_angle_feature,
            precise_char_corner_distance_feature, downsampled_char_prob_score_map, downsampled_char_mask,
            int(box.up), int(box.left), downsampled_label_point_y, downsampled_label_point_x, char_up_left_offsets,
            char_corner_angles, char_corner_distances, self.LABEL_POINT_SMOOTH_BETA, factors)

        # off-by-default terms (reference :272-307), each scaled by loss_factor like the rest (:344)
        extra = None
        if cfg.char_mask_focal_factor > 0:
            assert precise_char_mask_feature is not None
            extra = cfg.char_mask_focal_factor * self.char_mask_focal_with_logits(
                pred=_crop(precise_char_mask_feature, box), gt=downsampled_char_mask)
        if cfg.char_prob_l1_factor > 0:
            term = cfg.char_prob_l1_factor * ops.PointwiseLossFn.apply(
                _crop(precise_char_prob_feature, box), downsampled_char_prob_score_map, downsampled_char_mask,
                ops.SMOOTH_L1, 0.25, 0.0, True)
            extra = term if extra is None else extra + term
        if cfg.char_prob_wahr_factor > 0:
            term = cfg.char_prob_wahr_factor * ops.PointwiseLossFn.apply(
                _crop(precise_char_prob_feature, box), downsampled_char_prob_score_map, None,
                ops.WAHR, float(self.char_prob_wahr.gamma), 0.0, True)
            extra = term if extra is None else extra + term
        if extra is not None:
            loss = loss + extra * cfg.loss_factor
        return loss
