"""Parameter containers that reproduce the reference's ``state_dict`` key layout.

The reference builds its blocks as ``nn.Sequential`` chains of stock layers interleaved with parameter-free
``Permutation`` / ``GELU`` modules (model/helper.py:76-101), and the *positions* inside those chains are part of the
checkpoint format (``block.0``, ``block.2``, ``block.3``, ``block.5`` ...).  Here the chains are only containers: the
stock ``nn.Conv2d`` / ``nn.Linear`` / ``nn.LayerNorm`` objects own the fp32 master parameters (so construction order,
initialisation and RNG consumption equal the reference's), ``Slot`` marks the parameter-free positions, and the
enclosing modules' ``forward`` hand the parameters to the sm_100a kernels (``..ops``).  Nothing in here computes.
"""
from typing import List

from torch import nn

LN_EPS = 1e-6  # helper.ln (model/helper.py:96-97)


class Slot(nn.Module):
    """A parameter-free position of a reference chain (permute / GELU / pooling / Softplus)."""

    def __init__(self, what: str) -> None:
        super().__init__()
        self.what = what

    def extra_repr(self) -> str:
        return self.what

    def forward(self, *args, **kwargs):  # pragma: no cover - containers are never executed
        raise RuntimeError('vkocr_b200 parameter containers are not executable; call the enclosing module')


class Chain(nn.Sequential):
    """``nn.Sequential`` used purely for its child naming."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError('vkocr_b200 parameter containers are not executable; call the enclosing module')


def layer_norm(channels: int) -> nn.LayerNorm:
    return nn.LayerNorm(channels, eps=LN_EPS)


def pointwise_ln_gelu(in_channels: int, out_channels: int) -> Chain:
    """[permute, Linear, LN, permute, GELU]  (upernext.build_conv1x1_block :21-35, fpn.build_conv1x1_block :21-28)."""
    return Chain(Slot('bchw->bhwc'), nn.Linear(in_channels, out_channels), layer_norm(out_channels), Slot('bhwc->bchw'),
                 Slot('gelu'))


def conv_ln_gelu(in_channels: int, out_channels: int, kernel_size: int) -> Chain:
    """[Conv2d k x k 'same', permute, LN, permute, GELU]  (upernext.py:38-45, fpn.py:31-48)."""
    return Chain(nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2),
                 Slot('bchw->bhwc'), layer_norm(out_channels), Slot('bhwc->bchw'), Slot('gelu'))


def projection(in_channels: int, out_channels: int) -> Chain:
    """[permute, Linear, permute]  (head step 2: upernext.py:219-223, fpn.py:179-183)."""
    return Chain(Slot('bchw->bhwc'), nn.Linear(in_channels, out_channels), Slot('bhwc->bchw'))


def conv_ln_gelu_params(chain: Chain) -> List[nn.Parameter]:
    return [chain[0].weight, chain[0].bias, chain[2].weight, chain[2].bias]


def pointwise_ln_gelu_params(chain: Chain) -> List[nn.Parameter]:
    return [chain[1].weight, chain[1].bias, chain[2].weight, chain[2].bias]
