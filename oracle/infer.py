"""Oracle (test infrastructure): numpy/torch restatement of the tensor-side operations of the reference's inference passes.
Citations are ``file:line`` into ``/root/reference``.  Pinned: ``tests/golden/inference_tensor_ops.npz`` holds the outputs of
the UNMODIFIED reference (``inferencing/opt.py`` and ``AdaptiveScalingInferencing.rough_infer`` / ``.precise_infer``, run by
``oracle/make_golden.py`` with import-only stand-ins for the absent ``vkit`` / ``iolite`` packages and a stand-in model
that returns seeded network outputs) together with the reference's own test vectors (tests/test_evaluation.py:15-22);
``tests/test_oracle_golden.py`` holds these functions to them bit for bit.  ``peak_mask`` restates :477-491, whose only
arithmetic is the reference's own dependency ``scipy.ndimage.maximum_filter`` (called here as the reference calls it)."""
import math

import numpy as np
import torch


def pad_length_to_make_divisible(length: int, downsampling_factor: int):
    """vkit_open_model/inferencing/opt.py:16-18."""
    padded_length = math.ceil(length / downsampling_factor) * downsampling_factor
    return padded_length, padded_length - length


def pad_mat_to_make_divisible(mat: np.ndarray, downsampling_factor: int) -> np.ndarray:
    """vkit_open_model/inferencing/opt.py:16-41."""
    height, width = mat.shape[:2]
    ph = math.ceil(height / downsampling_factor) * downsampling_factor
    pw = math.ceil(width / downsampling_factor) * downsampling_factor
    if ph == height and pw == width:
        return mat
    shape = list(mat.shape)
    shape[0], shape[1] = ph, pw
    out = np.zeros(shape, dtype=mat.dtype)
    out[:height, :width] = mat
    return out


def network_input(image_hwc_u8: np.ndarray, downsampling_factor: int = 32) -> torch.Tensor:
    """vkit_open_model/inferencing/adaptive_scaling.py:109-121."""
    mat = pad_mat_to_make_divisible(image_hwc_u8, downsampling_factor)
    mat = np.transpose(mat, axes=(2, 0, 1)).astype(np.float32)
    return torch.from_numpy(mat).unsqueeze(0)


def rough_postprocess(mask_feature: torch.Tensor, height_feature: torch.Tensor, image_height: int, image_width: int,
                      padded_height: int, padded_width: int, upsampling_factor: int = 2, positive_thr: float = 0.5,
                      height_min: float = 3.0):
    """vkit_open_model/inferencing/adaptive_scaling.py:131-180 for one image: (h, w) logits / heights -> uint8 mask, fp32 map."""
    fdf = 4 // upsampling_factor
    m = torch.sigmoid(mask_feature.clone())
    mask = torch.greater_equal(m, positive_thr).numpy().astype(np.uint8)
    hmap = height_feature.numpy().astype(np.float32).copy()
    if image_height < padded_height:
        begin = math.ceil(image_height / fdf)
        if begin < mask.shape[0]:
            mask[begin:] = 0
            hmap[begin:] = 0.0
    if image_width < padded_width:
        begin = math.ceil(image_width / fdf)
        if begin < mask.shape[1]:
            mask[:, begin:] = 0
            hmap[:, begin:] = 0.0
    hmap[hmap < height_min] = 0.0
    return mask, hmap, (math.ceil(image_height / fdf), math.ceil(image_width / fdf))


def precise_postprocess(prob_feature: torch.Tensor, offset_feature: torch.Tensor, angle_feature: torch.Tensor,
                        distance_feature: torch.Tensor, image_height: int, image_width: int, padded_height: int,
                        padded_width: int, upsampling_factor: int = 2):
    """vkit_open_model/inferencing/adaptive_scaling.py:322-386 for one image: (1, C, h, w) network outputs -> score map (h, w),
    offsets (h, w, 2), softmax angle distribution (h, w, 4), distances (h, w, D), all float32 numpy."""
    prob = torch.sigmoid(prob_feature[0][0].clone()).numpy().astype(np.float32)
    offsets = torch.permute(offset_feature[0], [1, 2, 0]).numpy().astype(np.float32)
    angles = torch.softmax(torch.permute(angle_feature[0], [1, 2, 0]), dim=-1).numpy().astype(np.float32)
    distances = torch.permute(distance_feature[0], [1, 2, 0]).numpy().astype(np.float32)
    fdf = 4 // upsampling_factor
    if image_height < padded_height:
        begin = math.ceil(image_height / fdf)
        if begin < prob.shape[0]:
            prob[begin:] = 0.0
    if image_width < padded_width:
        begin = math.ceil(image_width / fdf)
        if begin < prob.shape[1]:
            prob[:, begin:] = 0.0
    return prob, offsets, angles, distances


def peak_mask(score_map: np.ndarray, char_mask: np.ndarray = None, size: int = 5, positive_thr: float = 0.7) -> np.ndarray:
    """vkit_open_model/inferencing/adaptive_scaling.py:477-491: maximum_filter peaks of the (masked) char-prob map."""
    from scipy.ndimage import maximum_filter
    mat = score_map.copy()
    if char_mask is not None:
        mat[~char_mask.astype(bool)] = 0
    np_local_maximum = maximum_filter(mat, size=size)
    np_mask = (np_local_maximum == mat)
    np_mask[mat < positive_thr] = 0
    return np_mask.astype(np.uint8)
