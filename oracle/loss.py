"""Oracle (test infrastructure): fp32 restatement of ``vkit_open_model.loss_function``.

Plain functions over tensors; formulas are written out (no call into torchvision) so that each term is the
arithmetic the CUDA loss kernels must reproduce.  Citations are ``file:line`` into ``/root/reference``.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EPS = 1e-6


@dataclass
class Box:
    """Inclusive pixel box; the only thing the losses use of ``vkit.element.Box`` (loss_function/adaptive_scaling.py:75-86)."""
    up: int
    down: int
    left: int
    right: int


def _masked_mean(loss: Tensor, mask: Optional[Tensor]) -> Tensor:
    # l1.py:42-47, l2.py:29-34, focal_with_logits.py:43-47
    if mask is None:
        return loss.mean()
    return (loss * mask).sum() / (mask.sum() + EPS)


def focal_with_logits(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None, alpha: float = 0.25,
                      gamma: float = 2.0) -> Tensor:
    """focal_with_logits.py:18-47 -> torchvision.ops.sigmoid_focal_loss (torchvision 0.26 ops/focal_loss.py)."""
    p = torch.sigmoid(pred)
    ce = F.binary_cross_entropy_with_logits(pred, gt, reduction='none')
    p_t = p * gt + (1 - p) * (1 - gt)
    loss = ce * (1 - p_t) ** gamma
    loss = (alpha * gt + (1 - alpha) * (1 - gt)) * loss
    return _masked_mean(loss, mask)


def dice(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """dice.py:17-35: one global ratio over the batch."""
    if mask is not None:
        pred = pred * mask
        gt = gt * mask
    inter = (pred * gt).sum()
    union = pred.sum() + gt.sum() + EPS
    return 1 - 2.0 * inter / union


def l1(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None, smooth: bool = False, smooth_beta: float = 1.0) -> Tensor:
    """l1.py:19-47 (F.l1_loss / F.smooth_l1_loss; integer targets are promoted to fp32)."""
    d = pred - gt.to(pred.dtype)
    a = d.abs()
    if smooth:
        loss = torch.where(a < smooth_beta, 0.5 * d * d / smooth_beta, a - 0.5 * smooth_beta)
    else:
        loss = a
    return _masked_mean(loss, mask)


def l2(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """l2.py:18-34."""
    return _masked_mean((pred - gt) ** 2, mask)


def cross_entropy_with_logits(pred: Tensor, gt: Tensor) -> Tensor:
    """cross_entropy_with_logits.py:16-19: class dim 1, probability targets, mean over the other dims."""
    logp = torch.log_softmax(pred, dim=1)
    return -(gt * logp).sum(dim=1).mean()


def weight_adaptive_heatmap_regression(pred: Tensor, gt: Tensor, gamma: float = 0.01) -> Tensor:
    """weight_adaptive_heatmap_regression.py:18-33."""
    soft = gt ** gamma
    weight = soft * (1 - pred) + (1 - soft) * pred
    return (weight * (pred - gt) ** 2).mean()


def weighted_bce_with_logits(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None, negative_ratio: float = 3.0) -> Tensor:
    """weighted_bce_with_logits.py:18-54: all positives + the top-k hardest negatives, k = min(round(3*#pos), #neg)."""
    pos = gt
    neg = 1 - gt
    if mask is not None:
        pos = pos * mask
        neg = neg * mask
    pos = pos.byte()
    n_pos = int(pos.long().sum())
    pos = pos.float()
    neg = neg.byte()
    n_neg = min(round(n_pos * negative_ratio), int(neg.long().sum()))
    neg = neg.float()
    loss = F.binary_cross_entropy_with_logits(pred, gt, reduction='none')
    neg_loss, _ = torch.topk((loss * neg).reshape(-1), n_neg)
    return ((loss * pos).sum() + neg_loss.sum()) / (n_pos + n_neg + EPS)


def _crop(x: Tensor, box: Box) -> Tensor:
    return x[:, box.up:box.down + 1, box.left:box.right + 1]


def rough_loss(
    rough_char_mask_feature: Tensor,
    rough_char_height_feature: Tensor,
    downsampled_mask: Tensor,
    downsampled_score_map: Tensor,
    downsampled_shape: Tuple[int, int],
    downsampled_core_box: Box,
    bce_factor: float = 0.0,
    focal_factor: float = 5.0,
    dice_factor: float = 1.0,
    l1_factor: float = 1.0,
    score_map_min: float = 1.1,
    height_min: float = 1.1,
) -> Tensor:
    """AdaptiveScalingRoughLossFunction.__call__ (loss_function/adaptive_scaling.py:53-131)."""
    assert rough_char_mask_feature.shape == rough_char_height_feature.shape
    assert tuple(rough_char_mask_feature.shape[1:]) == (1, *downsampled_shape)
    m = _crop(rough_char_mask_feature.squeeze(1), downsampled_core_box)
    h = _crop(rough_char_height_feature.squeeze(1), downsampled_core_box)
    loss = 0.0
    if bce_factor > 0.0:
        loss = loss + bce_factor * weighted_bce_with_logits(m, downsampled_mask)
    if focal_factor > 0.0:
        loss = loss + focal_factor * focal_with_logits(m, downsampled_mask)
    if dice_factor > 0.0:
        loss = loss + dice_factor * dice(torch.sigmoid(m), downsampled_mask)
    if l1_factor > 0.0:
        l1_mask = ((h > height_min) & (downsampled_score_map > score_map_min) & downsampled_mask.bool()).float()
        hc = torch.clamp(h, min=height_min)
        sc = torch.clamp(downsampled_score_map, min=score_map_min)
        loss = loss + l1_factor * l1(torch.log(hc), torch.log(sc), mask=l1_mask, smooth=True)
    return loss


def get_label_point_feature(feature: Tensor, y: Tensor, x: Tensor) -> Tensor:
    """loss_function/adaptive_scaling.py:167-179: (B,C,H,W) gathered at (B,P) points -> (B,P,C)."""
    b = feature.shape[0]
    return feature[torch.arange(b, device=feature.device)[:, None], :, y, x]


def precise_loss(
    precise_char_mask_feature: Optional[Tensor],
    precise_char_prob_feature: Tensor,
    precise_char_up_left_corner_offset_feature: Tensor,
    precise_char_corner_angle_feature: Tensor,
    precise_char_corner_distance_feature: Tensor,
    downsampled_char_prob_score_map: Tensor,
    downsampled_char_mask: Tensor,
    downsampled_shape: Tuple[int, int],
    downsampled_core_box: Box,
    downsampled_label_point_y: Tensor,
    downsampled_label_point_x: Tensor,
    char_up_left_offsets: Tensor,
    char_corner_angles: Tensor,
    char_corner_distances: Tensor,
    char_mask_focal_factor: float = 0.0,
    char_prob_l1_factor: float = 0.0,
    char_prob_pos_l2_factor: float = 2.0,
    char_prob_neg_l2_factor: float = 1.0,
    char_prob_wahr_factor: float = 0.0,
    char_up_left_offset_l1_factor: float = 1.0,
    char_up_left_distance_regulation_l1_factor: float = 1.0,
    char_corner_angle_cross_entropy_factor: float = 5.0,
    char_corner_distance_l1_factor: float = 1.0,
    loss_factor: float = 0.15,
) -> Tensor:
    """AdaptiveScalingPreciseLossFunction.__call__ (loss_function/adaptive_scaling.py:181-346)."""
    assert tuple(precise_char_prob_feature.shape[1:]) == (1, *downsampled_shape)
    box = downsampled_core_box
    prob = _crop(precise_char_prob_feature.squeeze(1), box)
    y, x = downsampled_label_point_y, downsampled_label_point_x
    off = get_label_point_feature(precise_char_up_left_corner_offset_feature, y, x)   # (B,P,2)
    ang = get_label_point_feature(precise_char_corner_angle_feature, y, x)            # (B,P,4)
    dist = get_label_point_feature(precise_char_corner_distance_feature, y, x)        # (B,P,4)
    loss = 0.0
    if char_mask_focal_factor > 0:
        assert precise_char_mask_feature is not None
        mk = _crop(precise_char_mask_feature.squeeze(1), box)
        loss = loss + char_mask_focal_factor * focal_with_logits(mk, downsampled_char_mask)
    if char_prob_l1_factor > 0 or char_prob_pos_l2_factor > 0 or char_prob_neg_l2_factor > 0 or char_prob_wahr_factor > 0:
        ps = torch.sigmoid(prob)
        if char_prob_l1_factor > 0:
            loss = loss + char_prob_l1_factor * l1(ps, downsampled_char_prob_score_map, downsampled_char_mask, True, 0.25)
        if char_prob_pos_l2_factor > 0:
            loss = loss + char_prob_pos_l2_factor * l2(ps, downsampled_char_prob_score_map, downsampled_char_mask)
        if char_prob_neg_l2_factor > 0:
            loss = loss + char_prob_neg_l2_factor * l2(ps, downsampled_char_prob_score_map, 1 - downsampled_char_mask)
        if char_prob_wahr_factor > 0:
            loss = loss + char_prob_wahr_factor * weight_adaptive_heatmap_regression(ps, downsampled_char_prob_score_map)
    if char_up_left_offset_l1_factor > 0:
        loss = loss + char_up_left_offset_l1_factor * l1(off, char_up_left_offsets, None, True, 2.5)
    if char_up_left_distance_regulation_l1_factor > 0:
        loss = loss + char_up_left_distance_regulation_l1_factor * l1(
            torch.linalg.norm(off, dim=2), dist[:, :, 0], None, True, 2.5)
    if char_corner_angle_cross_entropy_factor > 0:
        loss = loss + char_corner_angle_cross_entropy_factor * cross_entropy_with_logits(
            ang.transpose(1, 2), char_corner_angles.transpose(1, 2))
    if char_corner_distance_l1_factor > 0:
        loss = loss + char_corner_distance_l1_factor * l1(dist[:, :, 1:], char_corner_distances, None, True, 2.5)
    return loss * loss_factor
