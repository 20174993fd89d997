"""Tensor side of the reference's inference passes on the device.  Rough (``AdaptiveScalingInferencing.rough_infer``,
vkit_open_model/inferencing/adaptive_scaling.py:92-188): uint8 page image in, uint8 text mask and fp32 character-height
map out.  Precise (``precise_infer`` :295-396 and the peak picking of ``precise_build_grouped_polygons`` :477-491): uint8
image in, char-prob score map, NHWC offset / angle-distribution / distance arrays and the uint8 peak mask out.  The image
resize and the polygon post-processing of the reference (cv2 / scipy / vkit) are out of scope; the config field names are the reference's (``AdaptiveScalingInferencingConfig``, :40-60)."""
import math
from typing import Tuple

import attrs
import torch

from . import _lib as L

Tensor = torch.Tensor


@attrs.define
class RoughInferConfig:
    backbone_downsampling_factor: int = 32          # inferencing/adaptive_scaling.py:45
    rough_head_upsampling_factor: int = 2           # :46
    rough_char_mask_positive_thr: float = 0.5       # :48
    rough_valid_char_height_min: float = 3.0        # :49


def pad_length_to_make_divisible(length: int, downsampling_factor: int) -> Tuple[int, int]:
    """inferencing/opt.py:16-18."""
    padded = math.ceil(length / downsampling_factor) * downsampling_factor
    return padded, padded - length


def ingest_images(images_u8: Tensor, downsampling_factor: int = 32) -> Tensor:
    """(B, H, W, 3) or (H, W, 3) uint8 CUDA tensor -> (B, 3, Hp, Wp) fp32, zero-padded at the bottom / right to a multiple
    of ``downsampling_factor`` (inferencing/opt.py:21-41 + adaptive_scaling.py:116-121)."""
    L.require_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.shape[-1] != 3:
        raise L.VkocrError('ingest_images expects uint8 (.., H, W, 3) images')
    if images_u8.dim() == 3:
        images_u8 = images_u8.unsqueeze(0)
    images_u8 = images_u8.contiguous()
    B, H, W, _ = images_u8.shape
    Hp, _ = pad_length_to_make_divisible(H, downsampling_factor)
    Wp, _ = pad_length_to_make_divisible(W, downsampling_factor)
    out = torch.empty((B, 3, Hp, Wp), dtype=torch.float32, device=images_u8.device)
    L.check(L.LIB.vkocr_ingest_image_u8(L.ptr(images_u8), B, H, W, L.ptr(out), Hp, Wp, L.stream_ptr()), 'ingest_image_u8')
    return out


def rough_infer_tensors(model, images_u8: Tensor, config: RoughInferConfig = RoughInferConfig()):
    """uint8 page image(s) -> (rough_char_mask uint8 (B, h, w), rough_char_height_score_map fp32 (B, h, w), resized_shape).
    ``resized_shape`` = ceil(H / FDF), ceil(W / FDF) with FDF = 4 // rough_head_upsampling_factor (adaptive_scaling.py:135,
    176-180); rows / columns beyond it are padding and come back as zeros."""
    x = ingest_images(images_u8, config.backbone_downsampling_factor)
    H, W = (images_u8.shape[-3], images_u8.shape[-2])
    with torch.no_grad():
        logit, height = model.forward_rough(x)
    fdf = 4 // config.rough_head_upsampling_factor
    B, _, h, w = logit.shape
    assert (h, w) == (x.shape[2] // fdf, x.shape[3] // fdf), ((h, w), tuple(x.shape))
    valid_h, valid_w = math.ceil(H / fdf), math.ceil(W / fdf)
    mask = torch.empty((B, h, w), dtype=torch.uint8, device=x.device)
    hmap = torch.empty((B, h, w), dtype=torch.float32, device=x.device)
    L.check(L.LIB.vkocr_rough_postprocess(L.ptr(logit.contiguous()), L.ptr(height.contiguous()), B, h, w, valid_h, valid_w,
                                          config.rough_char_mask_positive_thr, config.rough_valid_char_height_min, L.ptr(mask),
                                          L.ptr(hmap), L.stream_ptr()), 'rough_postprocess')
    return mask, hmap, (valid_h, valid_w)


@attrs.define
class PreciseInferConfig:
    backbone_downsampling_factor: int = 32                      # inferencing/adaptive_scaling.py:45
    precise_head_upsampling_factor: int = 2                     # :51
    precise_build_polygons_positive_char_prob_thr: float = 0.7  # :59
    precise_build_polygons_maximum_filter_size: int = 5         # :60


def precise_infer_tensors(model, images_u8: Tensor, config: PreciseInferConfig = PreciseInferConfig()):
    """uint8 image(s) -> (char_prob_score_map fp32 (B, h, w), up_left_corner_offset fp32 (B, h, w, 2),
    corner_angle_distribution fp32 (B, h, w, 4), corner_distance fp32 (B, h, w, D)) as the reference's
    ``AdaptiveScalingInferencingPresiceInferResult`` holds them (inferencing/adaptive_scaling.py:326-396); rows / columns of
    the score map beyond ceil(H / FDF), ceil(W / FDF) are padding and come back as zeros."""
    x = ingest_images(images_u8, config.backbone_downsampling_factor)
    H, W = (images_u8.shape[-3], images_u8.shape[-2])
    with torch.no_grad():
        prob, offset, angle, distance = model.forward_precise(x)
    fdf = 4 // config.precise_head_upsampling_factor
    B, _, h, w = prob.shape
    D = int(distance.shape[1])
    if offset.shape[1] != 2 or angle.shape[1] != 4:
        raise L.VkocrError('precise_infer_tensors expects 2 offset and 4 angle channels')
    dev = x.device
    prob_map = torch.empty((B, h, w), dtype=torch.float32, device=dev)
    offsets = torch.empty((B, h, w, 2), dtype=torch.float32, device=dev)
    angles = torch.empty((B, h, w, 4), dtype=torch.float32, device=dev)
    distances = torch.empty((B, h, w, D), dtype=torch.float32, device=dev)
    L.check(L.LIB.vkocr_precise_postprocess(L.ptr(prob.contiguous()), L.ptr(offset.contiguous()), L.ptr(angle.contiguous()),
                                            L.ptr(distance.contiguous()), B, h, w, D, math.ceil(H / fdf), math.ceil(W / fdf),
                                            L.ptr(prob_map), L.ptr(offsets), L.ptr(angles), L.ptr(distances), L.stream_ptr()),
            'precise_postprocess')
    return prob_map, offsets, angles, distances


def find_peaks(char_prob_score_map: Tensor, char_mask: Tensor = None, config: PreciseInferConfig = PreciseInferConfig()) -> Tensor:
    """uint8 (B, h, w) mask of the character peaks: local maxima of the (masked) char-prob map under the maximum filter that
    reach the positive threshold (inferencing/adaptive_scaling.py:477-491)."""
    L.require_cuda(char_prob_score_map)
    pm = char_prob_score_map.contiguous()
    if pm.dim() == 2:
        pm = pm.unsqueeze(0)
    B, h, w = pm.shape
    cm = None
    if char_mask is not None:
        cm = char_mask.reshape(B, h, w).to(torch.uint8).contiguous()
    peaks = torch.empty((B, h, w), dtype=torch.uint8, device=pm.device)
    L.check(L.LIB.vkocr_peak_mask(L.ptr(pm), L.ptr(cm), B, h, w, int(config.precise_build_polygons_maximum_filter_size),
                                  float(config.precise_build_polygons_positive_char_prob_thr), L.ptr(peaks), L.stream_ptr()), 'peak_mask')
    return peaks
