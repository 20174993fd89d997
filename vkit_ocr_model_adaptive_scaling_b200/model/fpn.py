"""FPN neck / head on the B200 kernels — drop-in for ``vkit_open_model.model.fpn`` (reference model/fpn.py:21-208).

Same structure as the UperNeXt variant with nearest-neighbour resampling, 1x1 laterals on all four levels, 3x3
smoothing on all four levels, kaiming-normal initialisation and a 5x5 head conv for up-sampling factors in (2, 4].
"""
from typing import List, Sequence

import torch
from torch import nn

from .. import ops
from .. import runtime
from . import _holders as H

MODE = ops.NEAREST


def build_conv1x1_block(in_channels: int, out_channels: int):
    return H.pointwise_ln_gelu(in_channels, out_channels)


def build_conv3x3_block(in_channels: int, out_channels: int):
    return H.conv_ln_gelu(in_channels, out_channels, 3)


def build_conv5x5_block(in_channels: int, out_channels: int):
    return H.conv_ln_gelu(in_channels, out_channels, 5)


def _init_kaiming(root: nn.Module) -> None:
    for module in root.modules():  # fpn.py:104-108, 185-189
        if isinstance(module, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(module.weight)
            if module.bias is not None:
                nn.init.zeros_(module.bias)


class FpnNeck(nn.Module):

    @classmethod
    def build_step1_conv_blocks(cls, in_channels_group: Sequence[int], out_channels: int):
        return nn.ModuleList([build_conv1x1_block(c, out_channels) for c in in_channels_group])

    @classmethod
    def build_step2_conv_blocks(cls, in_channels_group: Sequence[int], out_channels: int):
        assert out_channels % len(in_channels_group) == 0
        inner_channels = out_channels // len(in_channels_group)
        return nn.ModuleList([build_conv3x3_block(out_channels, inner_channels) for _ in in_channels_group])

    def __init__(self, in_channels_group: Sequence[int], out_channels: int) -> None:
        super().__init__()
        assert len(in_channels_group) > 1
        self.step1_conv_blocks = self.build_step1_conv_blocks(in_channels_group, out_channels)
        self.step2_conv_blocks = self.build_step2_conv_blocks(in_channels_group, out_channels)
        _init_kaiming(self)

    def forward(self, features: List[torch.Tensor]) -> torch.Tensor:  # type: ignore
        num_features = len(features)
        assert num_features == len(self.step1_conv_blocks)
        dtype = runtime.compute_dtype()
        features = [ops.to_nhwc(f, dtype) for f in features]
        outputs = [
            ops.ConvLnGeluFn.apply(features[idx], *H.pointwise_ln_gelu_params(block))
            for idx, block in enumerate(self.step1_conv_blocks)
        ]
        for idx in range(num_features - 1, 0, -1):  # fpn.py:121-129, cumulative
            outputs[idx - 1] = ops.UpsampleAddFn.apply(outputs[idx - 1], outputs[idx], MODE)
        for idx, block in enumerate(self.step2_conv_blocks):
            outputs[idx] = ops.ConvLnGeluFn.apply(outputs[idx], *H.conv_ln_gelu_params(block))
        height, width = int(features[0].shape[-2]), int(features[0].shape[-1])
        return ops.UpsampleConcatFn.apply(MODE, height, width, *outputs)


class FpnHead(nn.Module):

    def __init__(self, in_channels: int, out_channels: int, upsampling_factor: int = 1, init_output_bias: float = 0.0):
        super().__init__()
        self.upsampling_factor = upsampling_factor
        inner_channels = (in_channels + out_channels) // 2
        if 1 <= self.upsampling_factor <= 2:
            self.step1_conv = build_conv3x3_block(in_channels, inner_channels)
        elif 2 < self.upsampling_factor <= 4:
            self.step1_conv = build_conv5x5_block(in_channels, inner_channels)
        else:
            raise NotImplementedError()
        self.step2_conv = H.projection(inner_channels, out_channels)
        _init_kaiming(self)
        nn.init.constant_(self.step2_conv[1].bias, init_output_bias)

    resample_mode = MODE

    def head_params(self) -> List[nn.Parameter]:
        return H.conv_ln_gelu_params(self.step1_conv) + [self.step2_conv[1].weight, self.step2_conv[1].bias]

    def forward(self, fpn_neck_feature: torch.Tensor) -> torch.Tensor:  # type: ignore
        x = ops.to_nhwc(fpn_neck_feature, runtime.compute_dtype())
        return ops.HeadGroupFn.apply(x, self.upsampling_factor, MODE, (False,), *self.head_params())[0]
