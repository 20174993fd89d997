"""The adaptive-scaling text-detection network on the B200 kernels — drop-in for
``vkit_open_model.model.adaptive_scaling`` (reference model/adaptive_scaling.py:27-237): same config classes / enum
values, module attribute names (hence ``state_dict`` keys), ``forward_rough`` / ``forward_precise`` signatures and
return tuples; no ``forward`` (the reference has none either).

All heads that read one neck tensor run as ONE fused group: the x2 up-sampled 384-channel operand is produced once,
one implicit-GEMM 3x3 convolution computes every head's inner channels (weights concatenated along N), and a
per-head tail kernel applies LayerNorm + GELU + the 1x1 projection (+ Softplus) and writes the fp32 NCHW maps.
"""
from enum import Enum, unique
import logging
from typing import Dict, List, Mapping, Sequence, Tuple

import attrs
import torch
from torch import nn

from .. import ops
from .. import runtime
from . import _holders as H
from .convnext import ConvNext
from .fpn import FpnHead, FpnNeck
from .upernext import UperNextHead, UperNextNeck

logger = logging.getLogger(__name__)


@unique
class AdaptiveScalingSize(Enum):
    TINY = 'tiny'
    SMALL = 'small'
    BASE = 'base'
    LARGE = 'large'


@unique
class AdaptiveScalingNeckHeadType(Enum):
    FPN = 'fpn'
    UPERNEXT = 'upernext'


@attrs.define
class AdaptiveScalingConfig:
    size: AdaptiveScalingSize = AdaptiveScalingSize.SMALL
    neck_head_type: AdaptiveScalingNeckHeadType = AdaptiveScalingNeckHeadType.FPN
    rough_upsampling_factor: int = 2
    rough_init_char_height_output_bias: float = 8.0
    precise_upsampling_factor: int = 2
    precise_enable_char_mask_head: bool = False


class SoftplusHead(nn.Sequential):
    """``nn.Sequential(head, nn.Softplus())`` of the reference (adaptive_scaling.py:93-102,133-141) with the Softplus
    fused into the head's tail kernel; child ``0`` is the head, child ``1`` is parameter-free."""

    def __init__(self, head: nn.Module) -> None:
        super().__init__(head, H.Slot('softplus'))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore
        head = self[0]
        x = ops.to_nhwc(x, runtime.compute_dtype())
        return ops.HeadGroupFn.apply(x, head.upsampling_factor, head.resample_mode, (True,), *head.head_params())[0]


def _run_head_group(neck_feature: torch.Tensor, heads: Sequence[nn.Module], label_points=None) -> Tuple[torch.Tensor, ...]:
    softplus = tuple(isinstance(h, SoftplusHead) for h in heads)
    cores = [h[0] if isinstance(h, SoftplusHead) else h for h in heads]
    factor, mode = cores[0].upsampling_factor, cores[0].resample_mode
    params: List[nn.Parameter] = []
    for core in cores:
        assert core.upsampling_factor == factor
        params.extend(core.head_params())
    if label_points is not None:
        per_head = [params[6 * i:6 * i + 6] for i in range(len(cores))]
        if ops.HeadGroupPointsFn.supported(neck_feature, factor, per_head):
            return ops.HeadGroupPointsFn.apply(neck_feature, factor, mode, softplus, label_points[0], label_points[1], *params)
    return ops.HeadGroupFn.apply(neck_feature, factor, mode, softplus, *params)


class AdaptiveScaling(nn.Module):

    def __init__(self, config: AdaptiveScalingConfig):
        super().__init__()
        creators = {
            AdaptiveScalingSize.TINY: ConvNext.create_tiny,
            AdaptiveScalingSize.SMALL: ConvNext.create_small,
            AdaptiveScalingSize.BASE: ConvNext.create_base,
            AdaptiveScalingSize.LARGE: ConvNext.create_large,
        }
        if config.size not in creators:
            raise NotImplementedError()
        # Module construction order == reference order (it fixes both the state_dict order and the RNG stream of
        # the initialisers): backbone, rough neck, rough heads, precise neck, precise heads.
        self.backbone = creators[config.size]()

        if config.neck_head_type == AdaptiveScalingNeckHeadType.FPN:
            neck_creator, head_creator = FpnNeck, FpnHead
        elif config.neck_head_type == AdaptiveScalingNeckHeadType.UPERNEXT:
            neck_creator, head_creator = UperNextNeck, UperNextHead
        else:
            raise NotImplementedError()

        neck_out_channels = self.backbone.in_channels_group[-2]

        self.rough_neck = neck_creator(in_channels_group=self.backbone.in_channels_group, out_channels=neck_out_channels)
        self.rough_char_mask_head = head_creator(
            in_channels=neck_out_channels, out_channels=1, upsampling_factor=config.rough_upsampling_factor)
        self.rough_char_height_head = SoftplusHead(head_creator(
            in_channels=neck_out_channels, out_channels=1, upsampling_factor=config.rough_upsampling_factor,
            init_output_bias=config.rough_init_char_height_output_bias))

        self.precise_neck = neck_creator(in_channels_group=self.backbone.in_channels_group, out_channels=neck_out_channels)
        self.precise_char_mask_head = None
        if config.precise_enable_char_mask_head:
            self.precise_char_mask_head = head_creator(
                in_channels=neck_out_channels, out_channels=1, upsampling_factor=config.precise_upsampling_factor)
        self.precise_char_prob_head = head_creator(
            in_channels=neck_out_channels, out_channels=1, upsampling_factor=config.precise_upsampling_factor)
        self.precise_char_up_left_corner_offset_head = head_creator(
            in_channels=neck_out_channels, out_channels=2, upsampling_factor=config.precise_upsampling_factor)
        self.precise_char_corner_angle_head = head_creator(
            in_channels=neck_out_channels, out_channels=4, upsampling_factor=config.precise_upsampling_factor)
        self.precise_char_corner_distance_head = SoftplusHead(head_creator(
            in_channels=neck_out_channels, out_channels=4, upsampling_factor=config.precise_upsampling_factor))

    def forward_rough(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B,3,H,W) fp32 image in 0..255 -> (mask logits, softplus char height), each (B,1,H/2,W/2) fp32
        (reference adaptive_scaling.py:143-154)."""
        feature = self.backbone(x)
        rough_neck_feature = self.rough_neck(feature)
        mask, height = _run_head_group(rough_neck_feature, (self.rough_char_mask_head, self.rough_char_height_head))
        return mask, height

    def forward_precise(self, x: torch.Tensor, label_points=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (prob logits (B,1,.), up-left offset (B,2,.), corner-angle logits (B,4,.), softplus corner distance (B,4,.))
        at (H/2, W/2), fp32 (reference adaptive_scaling.py:156-177).

        ``label_points`` (an extension, off by default): ``(downsampled_label_point_y, downsampled_label_point_x)``, (B, P)
        int64.  The precise loss reads the offset / angle / distance maps at these points only
        (loss_function/adaptive_scaling.py:235-260); given them, those three heads are evaluated at the label pixels alone
        (``ops.HeadGroupPointsFn``) and their maps are ZERO elsewhere -- loss and gradients are those of the dense call."""
        feature = self.backbone(x)
        precise_neck_feature = self.precise_neck(feature)
        if label_points is not None:
            precise_neck_feature = ops.to_nhwc(precise_neck_feature, runtime.compute_dtype())
            (prob,) = _run_head_group(precise_neck_feature, (self.precise_char_prob_head,))
            offset, angle, distance = _run_head_group(precise_neck_feature, (
                self.precise_char_up_left_corner_offset_head, self.precise_char_corner_angle_head,
                self.precise_char_corner_distance_head), label_points=label_points)
            return prob, offset, angle, distance
        prob, offset, angle, distance = _run_head_group(precise_neck_feature, (
            self.precise_char_prob_head,
            self.precise_char_up_left_corner_offset_head,
            self.precise_char_corner_angle_head,
            self.precise_char_corner_distance_head,
        ))
        return prob, offset, angle, distance

    # ---- gradient-inspection helpers of the reference (adaptive_scaling.py:179-237); plain host-side Python ----
    @classmethod
    def debug_get_rough_name_to_grad(cls, model: torch.nn.Module):
        rough_name_to_grad: Dict[str, torch.Tensor] = {}
        for name, parameter in model.named_parameters():
            if parameter.grad is None:
                continue
            assert name not in rough_name_to_grad
            rough_name_to_grad[name] = parameter.grad.cpu().clone()
        return rough_name_to_grad

    @classmethod
    def debug_get_precise_name_to_grad(cls, model: torch.nn.Module, rough_name_to_grad: Mapping[str, torch.Tensor]):
        precise_name_to_grad: Dict[str, torch.Tensor] = {}
        for name, parameter in model.named_parameters():
            if parameter.grad is None or name not in rough_name_to_grad:
                continue
            assert name not in precise_name_to_grad
            precise_name_to_grad[name] = parameter.grad.cpu() - rough_name_to_grad[name]
        return precise_name_to_grad

    @classmethod
    def debug_inspect_name_to_grad(cls, rough_name_to_grad: Mapping[str, torch.Tensor],
                                   precise_name_to_grad: Mapping[str, torch.Tensor]):
        names = sorted(set(rough_name_to_grad) & set(precise_name_to_grad))
        stats = {}
        for tag, table in (('rough', rough_name_to_grad), ('precise', precise_name_to_grad)):
            flat = torch.abs(torch.cat([table[name].view(-1) for name in names]))
            stats[tag] = (float(torch.mean(flat)), float(torch.std(flat)))
            logger.info(f'{tag}_abs_grads_mean = {stats[tag][0]}, {tag}_abs_grads_std = {stats[tag][1]}')
        logger.info('rough_abs_grads_mean / precise_abs_grads_mean = '
                    f'{stats["rough"][0] / (stats["precise"][0] + 1E-15)}')
