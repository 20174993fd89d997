"""UperNeXt neck / head on the B200 kernels — drop-in for ``vkit_open_model.model.upernext``
(reference model/upernext.py:21-248): same constructors, ``forward`` signatures and ``state_dict`` layout.

Every ``Linear/Conv -> LayerNorm -> GELU`` block is one tensor-core (implicit-)GEMM plus one fused LayerNorm+GELU pass;
``F.interpolate(mode='bilinear')`` + ``+=`` / ``torch.cat`` become resampling kernels that write straight into the
destination (channel slice of the concat buffer); the head fuses LayerNorm, GELU, the 1x1 projection and Softplus.
"""
from typing import List, Sequence

import torch
from torch import nn

from .. import ops
from .. import runtime
from . import _holders as H

MODE = ops.BILINEAR


def build_conv1x1_block(in_channels: int, out_channels: int, no_ln: bool = False):
    if no_ln:
        raise NotImplementedError('vkocr_b200: build_conv1x1_block(no_ln=True) is never used by the reference model')
    return H.pointwise_ln_gelu(in_channels, out_channels)


def build_conv3x3_block(in_channels: int, out_channels: int):
    return H.conv_ln_gelu(in_channels, out_channels, 3)


def _init_trunc_normal(root: nn.Module) -> None:
    for module in root.modules():  # upernext.py:157-161, 225-229
        if isinstance(module, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(module.weight, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)


class PpmBlock(nn.Module):
    """Pyramid pooling: [x] + [bilinear(conv1x1_block(adaptive_avg_pool(x, s))) for s in scales] -> concat ->
    conv3x3 block (reference upernext.py:48-84)."""

    def __init__(self, ppm_scales: Sequence[int], in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.ppm_scales = tuple(int(s) for s in ppm_scales)
        self.ap_conv_blocks = nn.ModuleList([
            H.Chain(H.Slot(f'adaptive_avg_pool2d({s})'), build_conv1x1_block(in_channels, out_channels))
            for s in self.ppm_scales
        ])
        self.final_conv_block = build_conv3x3_block(in_channels + len(self.ppm_scales) * out_channels, out_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore
        x = ops.to_nhwc(x, runtime.compute_dtype())
        height, width = int(x.shape[-2]), int(x.shape[-1])
        pyramid = [x]
        for scale, chain in zip(self.ppm_scales, self.ap_conv_blocks):
            pooled = ops.AvgPoolFn.apply(x, scale)
            pyramid.append(ops.ConvLnGeluFn.apply(pooled, *H.pointwise_ln_gelu_params(chain[1])))
        cat = ops.UpsampleConcatFn.apply(MODE, height, width, *pyramid)
        return ops.ConvLnGeluFn.apply(cat, *H.conv_ln_gelu_params(self.final_conv_block))


class UperNextNeck(nn.Module):
    """Laterals (+PPM on the coarsest level) -> cumulative top-down bilinear add -> 3x3 smoothing on all but the coarsest
    level -> bilinear to level-0 size -> concat (reference upernext.py:87-198)."""

    @classmethod
    def build_step1_conv_blocks(cls, in_channels_group: Sequence[int], ppm_scales: Sequence[int], inner_channels: int):
        blocks: List[nn.Module] = [build_conv1x1_block(c, inner_channels) for c in in_channels_group[:-1]]
        blocks.append(PpmBlock(ppm_scales=ppm_scales, in_channels=in_channels_group[-1], out_channels=inner_channels))
        return nn.ModuleList(blocks)

    @classmethod
    def build_step2_conv_blocks(cls, num_step1_conv_blocks: int, inner_channels: int):
        # the coarsest level already went through the PPM's 3x3 conv (upernext.py:126)
        return nn.ModuleList([build_conv3x3_block(inner_channels, inner_channels) for _ in range(num_step1_conv_blocks - 1)])

    def __init__(self, in_channels_group: Sequence[int], out_channels: int, ppm_scales: Sequence[int] = (1, 2, 3, 6)) -> None:
        super().__init__()
        assert len(in_channels_group) > 1
        assert out_channels % len(in_channels_group) == 0
        inner_channels = out_channels // len(in_channels_group)
        self.step1_conv_blocks = self.build_step1_conv_blocks(in_channels_group, ppm_scales, inner_channels)
        self.step2_conv_blocks = self.build_step2_conv_blocks(len(self.step1_conv_blocks), inner_channels)
        _init_trunc_normal(self)

    def forward(self, features: List[torch.Tensor]) -> torch.Tensor:  # type: ignore
        num_features = len(features)
        assert num_features == len(self.step1_conv_blocks)
        dtype = runtime.compute_dtype()
        features = [ops.to_nhwc(f, dtype) for f in features]
        outputs: List[torch.Tensor] = []
        for idx, block in enumerate(self.step1_conv_blocks):
            if isinstance(block, PpmBlock):
                outputs.append(block(features[idx]))
            else:
                outputs.append(ops.ConvLnGeluFn.apply(features[idx], *H.pointwise_ln_gelu_params(block)))
        for idx in range(num_features - 1, 0, -1):  # cumulative: uses the already-accumulated coarser level
            outputs[idx - 1] = ops.UpsampleAddFn.apply(outputs[idx - 1], outputs[idx], MODE)
        for idx, block in enumerate(self.step2_conv_blocks):
            outputs[idx] = ops.ConvLnGeluFn.apply(outputs[idx], *H.conv_ln_gelu_params(block))
        height, width = int(features[0].shape[-2]), int(features[0].shape[-1])
        return ops.UpsampleConcatFn.apply(MODE, height, width, *outputs)


class UperNextHead(nn.Module):
    """x`upsampling_factor` bilinear -> conv3x3 -> LN -> GELU -> Linear(inner -> out) (reference upernext.py:201-248)."""

    def __init__(self, in_channels: int, out_channels: int, upsampling_factor: int = 1, init_output_bias: float = 0.0):
        super().__init__()
        self.upsampling_factor = upsampling_factor
        inner_channels = (in_channels + out_channels) // 2
        self.step1_conv3x3 = build_conv3x3_block(in_channels, inner_channels)
        self.step2_conv1x1 = H.projection(inner_channels, out_channels)
        _init_trunc_normal(self)
        nn.init.constant_(self.step2_conv1x1[1].bias, init_output_bias)

    # -- fused-head protocol used by AdaptiveScaling (all heads of one neck share the up-sampled operand) --
    resample_mode = MODE

    def head_params(self) -> List[nn.Parameter]:
        return H.conv_ln_gelu_params(self.step1_conv3x3) + [self.step2_conv1x1[1].weight, self.step2_conv1x1[1].bias]

    def forward(self, fpn_neck_feature: torch.Tensor) -> torch.Tensor:  # type: ignore
        x = ops.to_nhwc(fpn_neck_feature, runtime.compute_dtype())
        return ops.HeadGroupFn.apply(x, self.upsampling_factor, MODE, (False,), *self.head_params())[0]
