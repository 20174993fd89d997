import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _exact_fp32_oracle():
    """The oracle must be true fp32 on the GPU: cuDNN convolutions default to TF32 (10-bit mantissa, ~3e-4)."""
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:   # pragma: no cover
        pass


_exact_fp32_oracle()


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
