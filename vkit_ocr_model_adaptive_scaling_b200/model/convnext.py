"""ConvNeXt backbone on the B200 kernels — drop-in for ``vkit_open_model.model.convnext`` (same constructors,
``forward`` signatures, attribute names and ``state_dict`` layout; reference model/convnext.py:20-235).

Activations travel between layers as NHWC buffers (logical shape stays (B, C, H, W)), so the reference's
permute / permute-back pairs around every LayerNorm and Linear (helper.py:76-93) disappear.  Per residual layer:

    dwconv7x7+bias            vkocr_dwconv7_fwd           (HBM-bound)
    LayerNorm                 vkocr_layernorm_fwd         (HBM-bound)
    Linear C->4C +bias +GELU  vkocr_gemm_nt (tcgen05)     epilogue keeps the pre-activation for backward
    Linear 4C->C +bias, x layer-scale, x stochastic-depth mask, + residual   vkocr_gemm_nt epilogue
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from .. import ops
from .. import runtime
from . import _holders as H


class ConvNextBlockLayer(nn.Module):
    """x + drop_path(block_scale * MLP(LN(dw7x7(x))))  (reference convnext.py:20-59)."""

    def __init__(self, in_channels: int, prob_bypass: float = 0.0) -> None:
        super().__init__()
        self.block = H.Chain(
            nn.Conv2d(in_channels, in_channels, kernel_size=7, padding=3, groups=in_channels),  # helper.dconv7x7
            H.Slot('bchw->bhwc'),
            H.layer_norm(in_channels),
            nn.Linear(in_channels, 4 * in_channels),
            H.Slot('gelu'),
            nn.Linear(4 * in_channels, in_channels),
            H.Slot('bhwc->bchw'),
        )
        self.block_scale = nn.parameter.Parameter(torch.ones(in_channels, 1, 1) * 1E-6)
        self.prob_bypass = prob_bypass

    def stochastic_depth_mask(self, batch: int, device) -> Optional[torch.Tensor]:
        """Per-sample keep mask / p_keep, drawn exactly like the reference (convnext.py:41-53): same torch calls on the
        device generator, in layer order, only in training mode and only for layers with prob_bypass > 0."""
        if not self.training or self.prob_bypass == 0.0:
            return None
        mask = torch.empty([batch, 1, 1, 1], dtype=torch.float32, device=device)
        prob_keep = 1.0 - self.prob_bypass
        mask.bernoulli_(prob_keep)
        if prob_keep > 0.0:
            mask.div_(prob_keep)
        return mask.reshape(batch)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore
        x = ops.to_nhwc(x, runtime.compute_dtype())
        b = self.block
        return ops.ConvNextLayerFn.apply(x, b[0].weight, b[0].bias, b[2].weight, b[2].bias, b[3].weight, b[3].bias,
                                         b[5].weight, b[5].bias, self.block_scale,
                                         self.stochastic_depth_mask(x.shape[0], x.device),
                                         1.0 / (1.0 - self.prob_bypass) if self.prob_bypass < 1.0 else 1.0)


class ConvNextBlock(nn.Module):
    """One stage: N residual layers -> LayerNorm (= the returned feature) -> optional 2x2/s2 down-sampling conv
    (reference convnext.py:62-101)."""

    def __init__(self, layer_idx_begin: int, layer_idx_end: int, in_channels: int, num_layers: int,
                 out_channels: Optional[int]) -> None:
        super().__init__()
        self.layers = nn.Sequential(*[
            ConvNextBlockLayer(in_channels=in_channels, prob_bypass=0.1 * (layer_idx_begin + idx) / layer_idx_end)
            for idx in range(num_layers)
        ])
        self.ln = H.Chain(H.Slot('bchw->bhwc'), H.layer_norm(in_channels), H.Slot('bhwc->bchw'))
        self.pconv2x2: Optional[nn.Module] = None
        if out_channels:
            self.pconv2x2 = nn.Conv2d(in_channels, out_channels, kernel_size=2, stride=2)  # helper.pconv2x2

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:  # type: ignore
        x = ops.to_nhwc(x, runtime.compute_dtype())
        x = self.layers(x)
        x = ops.LayerNormFn.apply(x, self.ln[1].weight, self.ln[1].bias, 0)
        feature = x
        if self.pconv2x2 is not None:
            x = ops.PatchConvFn.apply(x, self.pconv2x2.weight, self.pconv2x2.bias)
        return feature, x


class ConvNext(nn.Module):
    """Stem (4x4/s4 or 2x2/s2 patchify conv + LN) and four stages; returns the four stage features
    (reference convnext.py:104-235)."""

    @classmethod
    def build_stem(cls, stem_in_channels: int, block_in_channels: int, use_pconv2x2: bool):
        patch = 2 if use_pconv2x2 else 4
        return H.Chain(
            nn.Conv2d(stem_in_channels, block_in_channels, kernel_size=patch, stride=patch),
            H.Slot('bchw->bhwc'),
            H.layer_norm(block_in_channels),
            H.Slot('bhwc->bchw'),
        )

    @classmethod
    def build_blocks(cls, block_in_channels_and_num_layers: Sequence[Tuple[int, int]]):
        total = sum(n for _, n in block_in_channels_and_num_layers)
        blocks: List[ConvNextBlock] = []
        widths: List[int] = []
        begin = 0
        for idx, (channels, depth) in enumerate(block_in_channels_and_num_layers):
            last = idx + 1 == len(block_in_channels_and_num_layers)
            blocks.append(ConvNextBlock(
                layer_idx_begin=begin,
                layer_idx_end=total - 1,
                in_channels=channels,
                num_layers=depth,
                out_channels=None if last else block_in_channels_and_num_layers[idx + 1][0],
            ))
            widths.append(channels)
            begin += depth
        return nn.ModuleList(blocks), widths

    def __init__(self, stem_in_channels: int, block_in_channels_and_num_layers: Sequence[Tuple[int, int]],
                 stem_use_pconv2x2: bool):
        super().__init__()
        self.stem = self.build_stem(stem_in_channels, block_in_channels_and_num_layers[0][0], stem_use_pconv2x2)
        self.blocks, self.in_channels_group = self.build_blocks(block_in_channels_and_num_layers)
        for module in self.modules():  # reference init, convnext.py:169-173
            if isinstance(module, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(module.weight, std=0.02)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)

    @classmethod
    def _create(cls, widths: Sequence[int], depths: Sequence[int], stem_use_pconv2x2: bool):
        return ConvNext(stem_in_channels=3, block_in_channels_and_num_layers=tuple(zip(widths, depths)),
                        stem_use_pconv2x2=stem_use_pconv2x2)

    @classmethod
    def create_tiny(cls, stem_use_pconv2x2: bool = False):
        return cls._create((96, 192, 384, 768), (3, 3, 9, 3), stem_use_pconv2x2)

    @classmethod
    def create_small(cls, stem_use_pconv2x2: bool = False):
        return cls._create((96, 192, 384, 768), (3, 3, 27, 3), stem_use_pconv2x2)

    @classmethod
    def create_base(cls, stem_use_pconv2x2: bool = False):
        return cls._create((128, 256, 512, 1024), (3, 3, 27, 3), stem_use_pconv2x2)

    @classmethod
    def create_large(cls, stem_use_pconv2x2: bool = False):
        return cls._create((192, 384, 768, 1536), (3, 3, 27, 3), stem_use_pconv2x2)

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:  # type: ignore
        s = self.stem
        x = ops.StemFn.apply(x, s[0].weight, s[0].bias, s[2].weight, s[2].bias, runtime.compute_dtype())
        features: List[torch.Tensor] = []
        for block in self.blocks:
            feature, x = block(x)
            features.append(feature)
        return features
