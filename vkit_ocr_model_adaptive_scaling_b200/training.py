"""The training step of the adaptive-scaling model as the reference's loop runs it
(experiment/adaptive_scaling/train.py:397-478, minus data loading and the optimizer):

    forward_rough(rough images)   -> rough loss / 2   -> backward
    forward_precise(precise imgs) -> precise loss / 2 -> backward   (gradients accumulate on the shared backbone)
    [data parallel: bucketed gradient all-reduce overlapped with the backward passes]

Losses stay on the device (the reference's ``float(loss)`` host syncs, train.py:415,453, are left to the caller).
"""
from typing import Dict, Optional, Tuple

import torch

from .loss_function import AdaptiveScalingPreciseLossFunction, AdaptiveScalingRoughLossFunction
from .parallel import DataParallel

Tensor = torch.Tensor

ROUGH_KEYS = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
PRECISE_KEYS = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
                'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
                'char_corner_distances')


def batch_to_device(batch: Dict[str, object], device, non_blocking: bool = True) -> Dict[str, object]:
    """training/opt.py:17-27 of the reference: tensors move, everything else passes through."""
    return {k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def train_step(model, rough_loss_function: AdaptiveScalingRoughLossFunction,
               precise_loss_function: AdaptiveScalingPreciseLossFunction, rough_batch: Dict[str, object],
               precise_batch: Dict[str, object], dp: Optional[DataParallel] = None) -> Tuple[Tensor, Tensor]:
    """One two-pass step; returns the (un-halved) rough and precise losses as device tensors."""
    scale = 0.5 * (dp.loss_scale if dp is not None else 1.0)
    if dp is not None:
        dp.begin_step()
        dp.begin_pass(final=('rough',))
    mask, height = model.forward_rough(rough_batch['image'])
    rough_loss = rough_loss_function(
        rough_char_mask_feature=mask, rough_char_height_feature=height, **{k: rough_batch[k] for k in ROUGH_KEYS})
    (rough_loss * scale).backward()
    del mask, height
    if dp is not None:
        dp.begin_pass(final=None)
    prob, offset, angle, distance = model.forward_precise(precise_batch['image'])
    precise_loss = precise_loss_function(
        precise_char_mask_feature=None, precise_char_prob_feature=prob,
        precise_char_up_left_corner_offset_feature=offset, precise_char_corner_angle_feature=angle,
        precise_char_corner_distance_feature=distance, **{k: precise_batch[k] for k in PRECISE_KEYS})
    (precise_loss * scale).backward()
    if dp is not None:
        dp.finish_step()
    return rough_loss.detach(), precise_loss.detach()
