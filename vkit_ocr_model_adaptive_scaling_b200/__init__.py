"""B200-native forward / backward / loss path of vkit_open_model's adaptive-scaling text-detection network.

``model`` and ``loss_function`` mirror ``vkit_open_model.model`` / ``vkit_open_model.loss_function`` (same class
names, constructors, call signatures and ``state_dict`` layout); underneath, every compute step is a hand-written
sm_100a kernel in ``libvkocr_b200.so`` reached through the C ABI declared in ``include/vkocr_b200.h``.
Importing the package loads (building it first if absent) that library and raises if that is impossible:
there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (loads libvkocr_b200.so; raises when unavailable)
from . import runtime  # noqa: F401
from .runtime import compute_dtype, precision, set_compute_dtype  # noqa: F401
from . import ops  # noqa: F401
from . import model  # noqa: F401
from . import loss_function  # noqa: F401
from . import parallel  # noqa: F401

__all__ = ['model', 'loss_function', 'ops', 'parallel', 'runtime', 'set_compute_dtype', 'compute_dtype', 'precision']
from . import training  # noqa: F401
from . import inferencing  # noqa: F401
from . import checkpoint  # noqa: F401
