"""Oracle (test infrastructure): functional fp32 restatement of ``vkit_open_model.model``.

Every function takes the network parameters as a flat ``state_dict``-style mapping whose keys are the
reference's own ``state_dict()`` keys, so a reference checkpoint can be evaluated
without instantiating any ``nn.Module``.  The structure (depths, channels, neck type) is *derived from the
keys*.  All math is stock ``torch.nn.functional`` in the tensors' own dtype/device (fp32 on CPU in tests).

Reference citations are ``file:line`` into ``/root/reference``.
"""
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6  # vkit_open_model/model/helper.py:96-97


@dataclass
class ModelSpec:
    """Structure recovered from a state_dict (no reference code involved)."""
    channels: Tuple[int, ...]
    depths: Tuple[int, ...]
    neck_type: str  # 'upernext' | 'fpn'
    stem_patch: int

    @classmethod
    def from_state_dict(cls, sd: Mapping[str, Tensor], backbone_prefix: str = 'backbone.') -> 'ModelSpec':
        channels: List[int] = []
        depths: List[int] = []
        stage = 0
        while f'{backbone_prefix}blocks.{stage}.ln.1.weight' in sd:
            channels.append(int(sd[f'{backbone_prefix}blocks.{stage}.ln.1.weight'].shape[0]))
            depth = 0
            while f'{backbone_prefix}blocks.{stage}.layers.{depth}.block_scale' in sd:
                depth += 1
            depths.append(depth)
            stage += 1
        neck_type = 'none'
        for key in sd:
            if key.endswith('step1_conv_blocks.3.final_conv_block.0.weight') or '.ap_conv_blocks.' in key:
                neck_type = 'upernext'
                break
            if key.endswith('_neck.step2_conv_blocks.3.0.weight'):
                neck_type = 'fpn'
        stem_patch = int(sd[f'{backbone_prefix}stem.0.weight'].shape[-1])
        return cls(tuple(channels), tuple(depths), neck_type, stem_patch)


def _ln_nchw(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    # helper.py:86-97 -- permute to BHWC, LayerNorm over C (biased variance, eps 1e-6), permute back.
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, LN_EPS)
    return y.permute(0, 3, 1, 2)


def _linear_nchw(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    # helper.py:18-22 -- nn.Linear applied on the channel axis of a BHWC view == 1x1 conv.
    return F.linear(x.permute(0, 2, 3, 1), w, b).permute(0, 3, 1, 2)


def stochastic_depth_probs(depths: Sequence[int]) -> List[List[float]]:
    """prob_bypass of every layer: 0.1 * global_layer_idx / (num_layers_total - 1) (convnext.py:75-76,133-135)."""
    total = sum(depths)
    end = total - 1
    out: List[List[float]] = []
    begin = 0
    for d in depths:
        out.append([0.1 * (begin + i) / end for i in range(d)])
        begin += d
    return out


def convnext_layer(sd: Mapping[str, Tensor], p: str, x: Tensor, drop_mask: Optional[Tensor] = None) -> Tensor:
    """One ConvNeXt residual layer (convnext.py:29-59).

    block = dw7x7(pad 3) -> LN(C) -> Linear(C,4C) -> exact GELU -> Linear(4C,C); out = x + mask * scale * block(x).
    ``drop_mask`` is the already-scaled per-sample stochastic-depth multiplier of shape (B,1,1,1), or None.
    """
    c = x.shape[1]
    y = F.conv2d(x, sd[p + 'block.0.weight'], sd[p + 'block.0.bias'], padding=3, groups=c)
    y = _ln_nchw(y, sd[p + 'block.2.weight'], sd[p + 'block.2.bias'])
    y = _linear_nchw(y, sd[p + 'block.3.weight'], sd[p + 'block.3.bias'])
    y = F.gelu(y)  # exact erf GELU, helper.py:100-101
    y = _linear_nchw(y, sd[p + 'block.5.weight'], sd[p + 'block.5.bias'])
    y = sd[p + 'block_scale'] * y
    if drop_mask is not None:
        y = drop_mask * y
    return y + x


def convnext_forward(
    sd: Mapping[str, Tensor],
    x: Tensor,
    prefix: str = 'backbone.',
    drop_masks: Optional[Mapping[Tuple[int, int], Tensor]] = None,
) -> List[Tensor]:
    """ConvNext.forward (convnext.py:227-235): stem, then per stage [layers -> LN -> (feature, pconv2x2)]."""
    spec = ModelSpec.from_state_dict(sd, prefix)
    w = sd[prefix + 'stem.0.weight']
    x = F.conv2d(x, w, sd[prefix + 'stem.0.bias'], stride=w.shape[-1])  # convnext.py:106-123
    x = _ln_nchw(x, sd[prefix + 'stem.2.weight'], sd[prefix + 'stem.2.bias'])
    features: List[Tensor] = []
    for s, depth in enumerate(spec.depths):
        for l in range(depth):
            mask = None if drop_masks is None else drop_masks.get((s, l))
            x = convnext_layer(sd, f'{prefix}blocks.{s}.layers.{l}.', x, mask)
        x = _ln_nchw(x, sd[f'{prefix}blocks.{s}.ln.1.weight'], sd[f'{prefix}blocks.{s}.ln.1.bias'])  # :83-87
        features.append(x)
        key = f'{prefix}blocks.{s}.pconv2x2.weight'
        if key in sd:  # convnext.py:89-99
            x = F.conv2d(x, sd[key], sd[f'{prefix}blocks.{s}.pconv2x2.bias'], stride=2)
    return features


def _conv1x1_block(sd: Mapping[str, Tensor], p: str, x: Tensor) -> Tensor:
    # upernext.py:21-35 / fpn.py:21-28: permute -> Linear(.1) -> LN(.2) -> permute -> GELU
    y = _linear_nchw(x, sd[p + '1.weight'], sd[p + '1.bias'])
    y = _ln_nchw(y, sd[p + '2.weight'], sd[p + '2.bias'])
    return F.gelu(y)


def _convkxk_block(sd: Mapping[str, Tensor], p: str, x: Tensor) -> Tensor:
    # upernext.py:38-45 / fpn.py:31-48: conv kxk 'same' (.0) -> LN (.2) -> GELU
    w = sd[p + '0.weight']
    y = F.conv2d(x, w, sd[p + '0.bias'], padding=w.shape[-1] // 2)
    y = _ln_nchw(y, sd[p + '2.weight'], sd[p + '2.bias'])
    return F.gelu(y)


def _ppm(sd: Mapping[str, Tensor], p: str, x: Tensor, scales: Sequence[int]) -> Tensor:
    # upernext.py:73-84: [x] + [bilinear(conv1x1_block(adaptive_avg_pool(x, s)))] -> cat -> conv3x3 block
    size = (x.shape[-2], x.shape[-1])
    feats = [x]
    for i, s in enumerate(scales):
        f = F.adaptive_avg_pool2d(x, s)
        f = _conv1x1_block(sd, f'{p}ap_conv_blocks.{i}.1.', f)
        feats.append(F.interpolate(f, size=size, mode='bilinear'))
    return _convkxk_block(sd, p + 'final_conv_block.', torch.cat(feats, dim=1))


def neck_forward(
    sd: Mapping[str, Tensor],
    prefix: str,
    features: Sequence[Tensor],
    neck_type: str,
    ppm_scales: Sequence[int] = (1, 2, 3, 6),
) -> Tensor:
    """UperNextNeck.forward (upernext.py:163-198) / FpnNeck.forward (fpn.py:110-146)."""
    n = len(features)
    mode = 'bilinear' if neck_type == 'upernext' else 'nearest'
    outs: List[Tensor] = []
    for i in range(n):
        p = f'{prefix}step1_conv_blocks.{i}.'
        if neck_type == 'upernext' and i == n - 1:
            outs.append(_ppm(sd, p, features[i], ppm_scales))
        else:
            outs.append(_conv1x1_block(sd, p, features[i]))
    # top-down, cumulative (upernext.py:174-182, fpn.py:121-129)
    for i in range(n - 1, 0, -1):
        size = (outs[i - 1].shape[-2], outs[i - 1].shape[-1])
        outs[i - 1] = outs[i - 1] + F.interpolate(outs[i], size=size, mode=mode)
    # step 2: UperNeXt skips the last level (upernext.py:126,185-186); FPN applies to all (fpn.py:78,132-133)
    n2 = n - 1 if neck_type == 'upernext' else n
    for i in range(n2):
        outs[i] = _convkxk_block(sd, f'{prefix}step2_conv_blocks.{i}.', outs[i])
    size0 = (features[0].shape[-2], features[0].shape[-1])
    for i in range(1, n):
        outs[i] = F.interpolate(outs[i], size=size0, mode=mode)
    return torch.cat(outs, dim=1)


def head_forward(
    sd: Mapping[str, Tensor],
    prefix: str,
    x: Tensor,
    neck_type: str,
    upsampling_factor: int,
    softplus: bool = False,
) -> Tensor:
    """UperNextHead.forward (upernext.py:233-248) / FpnHead.forward (fpn.py:193-208) (+ nn.Softplus, adaptive_scaling.py:101,140)."""
    if neck_type == 'upernext':
        k1, k2, mode = 'step1_conv3x3.', 'step2_conv1x1.', 'bilinear'
    else:
        k1, k2, mode = 'step1_conv.', 'step2_conv.', 'nearest'
    if upsampling_factor > 1:
        x = F.interpolate(
            x, size=(x.shape[-2] * upsampling_factor, x.shape[-1] * upsampling_factor), mode=mode)
    x = _convkxk_block(sd, prefix + k1, x)
    x = _linear_nchw(x, sd[prefix + k2 + '1.weight'], sd[prefix + k2 + '1.bias'])
    if softplus:
        x = F.softplus(x)
    return x


def forward_rough(
    sd: Mapping[str, Tensor],
    x: Tensor,
    upsampling_factor: int = 2,
    drop_masks: Optional[Mapping[Tuple[int, int], Tensor]] = None,
) -> Tuple[Tensor, Tensor]:
    """AdaptiveScaling.forward_rough (adaptive_scaling.py:143-154)."""
    spec = ModelSpec.from_state_dict(sd)
    feats = convnext_forward(sd, x, 'backbone.', drop_masks)
    neck = neck_forward(sd, 'rough_neck.', feats, spec.neck_type)
    mask = head_forward(sd, 'rough_char_mask_head.', neck, spec.neck_type, upsampling_factor)
    height = head_forward(sd, 'rough_char_height_head.0.', neck, spec.neck_type, upsampling_factor, softplus=True)
    return mask, height


def forward_precise(
    sd: Mapping[str, Tensor],
    x: Tensor,
    upsampling_factor: int = 2,
    drop_masks: Optional[Mapping[Tuple[int, int], Tensor]] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """AdaptiveScaling.forward_precise (adaptive_scaling.py:156-177)."""
    spec = ModelSpec.from_state_dict(sd)
    feats = convnext_forward(sd, x, 'backbone.', drop_masks)
    neck = neck_forward(sd, 'precise_neck.', feats, spec.neck_type)
    t = spec.neck_type
    prob = head_forward(sd, 'precise_char_prob_head.', neck, t, upsampling_factor)
    offset = head_forward(sd, 'precise_char_up_left_corner_offset_head.', neck, t, upsampling_factor)
    angle = head_forward(sd, 'precise_char_corner_angle_head.', neck, t, upsampling_factor)
    dist = head_forward(sd, 'precise_char_corner_distance_head.0.', neck, t, upsampling_factor, softplus=True)
    return prob, offset, angle, dist
