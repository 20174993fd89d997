"""Oracle (test infrastructure): numpy/torch restatement of the tensor-side operations of the reference's rough inference
pass.  Citations are ``file:line`` into ``/root/reference``.  Parity unpinned by golden vectors of the reference itself:
``vkit_open_model.inferencing`` imports ``vkit`` / ``iolite`` (absent), so these few lines are restated from source."""
import math

import numpy as np
import torch


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
