"""Process-wide settings of the B200 path: the storage/compute dtype of activations.

``bfloat16`` (default): activations are stored in bf16, every contraction runs on the tcgen05 tensor cores with fp32
accumulation in tensor memory, and all LayerNorm / GELU / loss math is fp32 in registers (north_star "bf16 mode",
rel 2e-2).  ``float32``: activations are fp32 and the contractions use the fp32 SIMT kernel (north_star "fp32 mode",
rel 1e-4).  Parameters are always the reference's fp32 master weights.
"""
import contextlib

import torch

_STATE = {'dtype': torch.bfloat16}


def set_compute_dtype(dtype: torch.dtype) -> None:
    if dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f'compute dtype must be torch.bfloat16 or torch.float32, got {dtype}')
    _STATE['dtype'] = dtype


def compute_dtype() -> torch.dtype:
    return _STATE['dtype']


@contextlib.contextmanager
def precision(dtype: torch.dtype):
    prev = _STATE['dtype']
    set_compute_dtype(dtype)
    try:
        yield
    finally:
        _STATE['dtype'] = prev
