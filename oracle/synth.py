"""Oracle (test infrastructure): deterministic synthetic weights and batches.

``synth_state_dict`` writes a reference-layout ``state_dict`` from a seeded CPU generator, so
that the same weights can be rebuilt on any box without the reference; the value distributions are chosen so that
every sub-module matters numerically (``block_scale`` ~ U(0.5, 1) instead of the 1e-6 init, convnext.py:38).
``synth_rough_batch`` / ``synth_precise_batch`` follow the input recipes of the reference's integration test
(tests/test_adaptive_scaling.py:126-169) and the collate schema (dataset/adaptive_scaling.py:282-368).
"""
from collections import OrderedDict
from typing import Dict, Sequence, Tuple

import torch

from .loss import Box

Tensor = torch.Tensor

SIZES = {
    # convnext.py:175-225
    'tiny': ((96, 192, 384, 768), (3, 3, 9, 3)),
    'small': ((96, 192, 384, 768), (3, 3, 27, 3)),
    'base': ((128, 256, 512, 1024), (3, 3, 27, 3)),
    'large': ((192, 384, 768, 1536), (3, 3, 27, 3)),
}


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)
        self.sd: 'OrderedDict[str, Tensor]' = OrderedDict()

    def normal(self, key: str, shape: Sequence[int], std: float) -> None:
        self.sd[key] = torch.randn(tuple(shape), generator=self.g) * std

    def uniform(self, key: str, shape: Sequence[int], lo: float, hi: float) -> None:
        self.sd[key] = torch.rand(tuple(shape), generator=self.g) * (hi - lo) + lo

    def weight(self, key: str, shape: Sequence[int]) -> None:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        self.normal(key, shape, (1.0 / fan_in) ** 0.5)

    def bias(self, key: str, n: int) -> None:
        self.normal(key, (n,), 0.1)

    def ln(self, prefix: str, n: int) -> None:
        self.uniform(prefix + 'weight', (n,), 0.5, 1.5)
        self.normal(prefix + 'bias', (n,), 0.1)


def backbone_state_dict(g: _Gen, channels: Sequence[int], depths: Sequence[int], prefix: str = 'backbone.',
                        stem_patch: int = 4, stem_in: int = 3) -> None:
    c0 = channels[0]
    # raw 0..255 pixels come in (dataset/adaptive_scaling.py:296): keep the stem output O(1)
    g.normal(prefix + 'stem.0.weight', (c0, stem_in, stem_patch, stem_patch), 0.02 / 8)
    g.bias(prefix + 'stem.0.bias', c0)
    g.ln(prefix + 'stem.2.', c0)
    for s, (c, d) in enumerate(zip(channels, depths)):
        for l in range(d):
            p = f'{prefix}blocks.{s}.layers.{l}.'
            g.uniform(p + 'block_scale', (c, 1, 1), 0.5, 1.0)
            g.weight(p + 'block.0.weight', (c, 1, 7, 7))
            g.bias(p + 'block.0.bias', c)
            g.ln(p + 'block.2.', c)
            g.weight(p + 'block.3.weight', (4 * c, c))
            g.bias(p + 'block.3.bias', 4 * c)
            g.weight(p + 'block.5.weight', (c, 4 * c))
            g.bias(p + 'block.5.bias', c)
        g.ln(f'{prefix}blocks.{s}.ln.1.', c)
        if s + 1 < len(channels):
            g.weight(f'{prefix}blocks.{s}.pconv2x2.weight', (channels[s + 1], c, 2, 2))
            g.bias(f'{prefix}blocks.{s}.pconv2x2.bias', channels[s + 1])


def neck_state_dict(g: _Gen, prefix: str, neck_type: str, in_channels: Sequence[int], out_channels: int,
                    ppm_scales: Sequence[int] = (1, 2, 3, 6)) -> None:
    n = len(in_channels)
    inner = out_channels // n
    if neck_type == 'upernext':  # upernext.py:135-161
        for i in range(n - 1):
            p = f'{prefix}step1_conv_blocks.{i}.'
            g.weight(p + '1.weight', (inner, in_channels[i]))
            g.bias(p + '1.bias', inner)
            g.ln(p + '2.', inner)
        p = f'{prefix}step1_conv_blocks.{n - 1}.'
        for k in range(len(ppm_scales)):
            q = f'{p}ap_conv_blocks.{k}.1.'
            g.weight(q + '1.weight', (inner, in_channels[-1]))
            g.bias(q + '1.bias', inner)
            g.ln(q + '2.', inner)
        cat = in_channels[-1] + len(ppm_scales) * inner
        g.weight(p + 'final_conv_block.0.weight', (inner, cat, 3, 3))
        g.bias(p + 'final_conv_block.0.bias', inner)
        g.ln(p + 'final_conv_block.2.', inner)
        for i in range(n - 1):
            p = f'{prefix}step2_conv_blocks.{i}.'
            g.weight(p + '0.weight', (inner, inner, 3, 3))
            g.bias(p + '0.bias', inner)
            g.ln(p + '2.', inner)
    elif neck_type == 'fpn':  # fpn.py:87-108
        for i in range(n):
            p = f'{prefix}step1_conv_blocks.{i}.'
            g.weight(p + '1.weight', (out_channels, in_channels[i]))
            g.bias(p + '1.bias', out_channels)
            g.ln(p + '2.', out_channels)
        for i in range(n):
            p = f'{prefix}step2_conv_blocks.{i}.'
            g.weight(p + '0.weight', (inner, out_channels, 3, 3))
            g.bias(p + '0.bias', inner)
            g.ln(p + '2.', inner)
    else:
        raise ValueError(neck_type)


def head_state_dict(g: _Gen, prefix: str, neck_type: str, in_channels: int, out_channels: int, ksize: int = 3,
                    out_bias: float = 0.0) -> None:
    inner = (in_channels + out_channels) // 2  # upernext.py:213, fpn.py:163
    k1, k2 = ('step1_conv3x3.', 'step2_conv1x1.') if neck_type == 'upernext' else ('step1_conv.', 'step2_conv.')
    g.weight(prefix + k1 + '0.weight', (inner, in_channels, ksize, ksize))
    g.bias(prefix + k1 + '0.bias', inner)
    g.ln(prefix + k1 + '2.', inner)
    g.weight(prefix + k2 + '1.weight', (out_channels, inner))
    g.bias(prefix + k2 + '1.bias', out_channels)
    g.sd[prefix + k2 + '1.bias'] += out_bias


def synth_state_dict(size: str = 'tiny', neck_type: str = 'upernext', seed: int = 133,
                     rough_height_bias: float = 8.0) -> Dict[str, Tensor]:
    """Reference-layout weights of ``AdaptiveScaling(AdaptiveScalingConfig(size, neck_type))`` (adaptive_scaling.py:51-141)."""
    channels, depths = SIZES[size]
    g = _Gen(seed)
    backbone_state_dict(g, channels, depths)
    neck_out = channels[-2]  # adaptive_scaling.py:79
    neck_state_dict(g, 'rough_neck.', neck_type, channels, neck_out)
    head_state_dict(g, 'rough_char_mask_head.', neck_type, neck_out, 1)
    head_state_dict(g, 'rough_char_height_head.0.', neck_type, neck_out, 1, out_bias=rough_height_bias)
    neck_state_dict(g, 'precise_neck.', neck_type, channels, neck_out)
    head_state_dict(g, 'precise_char_prob_head.', neck_type, neck_out, 1)
    head_state_dict(g, 'precise_char_up_left_corner_offset_head.', neck_type, neck_out, 2)
    head_state_dict(g, 'precise_char_corner_angle_head.', neck_type, neck_out, 4)
    head_state_dict(g, 'precise_char_corner_distance_head.0.', neck_type, neck_out, 4)
    return g.sd


def synth_image(batch: int, height: int, width: int, seed: int = 133) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, 3, height, width), generator=g).float()


def synth_rough_batch(batch: int, height: int, width: int, seed: int = 133, inset: int = 10) -> Dict[str, object]:
    g = torch.Generator().manual_seed(seed + 1)
    dh, dw = height // 2, width // 2
    box = Box(up=inset, down=dh - inset - 1, left=inset, right=dw - inset - 1)
    ch, cw = box.down - box.up + 1, box.right - box.left + 1
    return {
        'image': synth_image(batch, height, width, seed),
        'downsampled_mask': (torch.rand(batch, ch, cw, generator=g) > 0.5).float(),
        'downsampled_score_map': torch.rand(batch, ch, cw, generator=g) * 12.0 + 0.5,
        'downsampled_shape': (dh, dw),
        'downsampled_core_box': box,
    }


def synth_precise_batch(batch: int, height: int, width: int, points: int = 200, seed: int = 133,
                        inset: int = 10) -> Dict[str, object]:
    g = torch.Generator().manual_seed(seed + 2)
    dh, dw = height // 2, width // 2
    box = Box(up=inset, down=dh - inset - 1, left=inset, right=dw - inset - 1)
    ch, cw = box.down - box.up + 1, box.right - box.left + 1
    return {
        'image': synth_image(batch, height, width, seed + 7),
        'downsampled_char_prob_score_map': torch.rand(batch, ch, cw, generator=g),
        'downsampled_char_mask': (torch.rand(batch, ch, cw, generator=g) > 0.5).float(),
        'downsampled_shape': (dh, dw),
        'downsampled_core_box': box,
        'downsampled_label_point_y': torch.randint(inset, dh - inset, (batch, points), generator=g),
        'downsampled_label_point_x': torch.randint(inset, dw - inset, (batch, points), generator=g),
        'char_up_left_offsets': torch.randint(-20, 21, (batch, points, 2), generator=g),
        'char_corner_angles': torch.softmax(torch.rand(batch, points, 4, generator=g), dim=-1),
        'char_corner_distances': torch.rand(batch, points, 3, generator=g) * 20.0,
    }
