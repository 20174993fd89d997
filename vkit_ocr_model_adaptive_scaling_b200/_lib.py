"""ctypes binding of ``libvkocr_b200.so`` (the C ABI declared in ``include/vkocr_b200.h``).

The product path has no CPU or PyTorch fallback: if the shared object cannot be built or loaded, importing this
module raises.
"""
import ctypes
import os
from typing import Optional

import torch

from . import _build

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float

F32 = 0
BF16 = 1


def dtype_tag(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise TypeError(f'vkocr_b200 kernels support float32 and bfloat16 storage, got {dtype}')


class ConvGeom(ctypes.Structure):
    _fields_ = [
        ('batch', c_int), ('H', c_int), ('W', c_int), ('ks', c_int), ('C', c_int),
        ('ld_x', c_ll), ('c_pad', c_int),
    ]


class Epilogue(ctypes.Structure):
    _fields_ = [
        ('out', c_void_p), ('ldo', c_ll), ('out_f32', c_int), ('accumulate', c_int),
        ('out_pre', c_void_p), ('ld_pre', c_ll),
        ('bias', c_void_p), ('act', c_int),
        ('col_scale', c_void_p), ('row_scale', c_void_p), ('rows_per_group', c_int),
        ('residual', c_void_p), ('ld_res', c_ll),
        ('aux', c_void_p), ('ld_aux', c_ll),
        ('tn_s_tap', c_ll), ('tn_s_i', c_ll), ('tn_s_j', c_ll),
    ]


MAX_HEADS = 4


class HeadTail(ctypes.Structure):
    _fields_ = [
        ('num_heads', c_int), ('slot', c_int), ('pixels_per_image', c_ll),
        ('gamma', c_void_p * MAX_HEADS), ('beta', c_void_p * MAX_HEADS), ('w2', c_void_p * MAX_HEADS),
        ('b2', c_void_p * MAX_HEADS), ('out', c_void_p * MAX_HEADS),
        ('inner', c_int * MAX_HEADS), ('out_channels', c_int * MAX_HEADS), ('softplus', c_int * MAX_HEADS),
    ]


_P = ctypes.POINTER

# name -> argtypes of every compute entry point declared in include/vkocr_b200.h (all return int status)
_SIGNATURES = {
    'vkocr_abi_version': [],
    'vkocr_device_check': [c_int],
    'vkocr_gemm_nt': [c_int, c_int, c_void_p, _P(ConvGeom), c_void_p, c_int, _P(Epilogue), c_void_p],
    'vkocr_gemm_nt_heads': [c_int, c_void_p, _P(ConvGeom), c_void_p, c_int, _P(Epilogue), _P(HeadTail), c_void_p],
    'vkocr_gemm_tn': [c_int, c_int, c_void_p, _P(ConvGeom), c_void_p, c_int, c_ll, _P(Epilogue), c_void_p],
    'vkocr_head_combine_fwd': [c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, _P(HeadTail), c_void_p,
                               c_ll, c_int, c_void_p],
    'vkocr_head_combine_bwd': [c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p],
    'vkocr_scatter_add_f32': [c_void_p, c_ll, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_ll, c_void_p],
    'vkocr_layernorm_fwd': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_float, c_int,
                            c_void_p, c_void_p, c_void_p],
    'vkocr_layernorm_bwd': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                            c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_colsum': [c_int, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p],
    'vkocr_scale_rows_colsum': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_int, c_void_p, c_void_p],
    'vkocr_dwconv7_fwd': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                          c_void_p, c_ll, c_void_p],
    'vkocr_dwconv7_wgrad': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    'vkocr_upsample_fwd': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_void_p],
    'vkocr_upsample_bwd': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_void_p],
    'vkocr_upsample_bwd_separable': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p],
    'vkocr_avgpool_fwd': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    'vkocr_avgpool_fwd_separable': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p],
    'vkocr_avgpool_bwd': [c_int, c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p],
    'vkocr_patchify_image': [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    'vkocr_space_to_depth2': [c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_void_p],
    'vkocr_copy_channels': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_int, c_void_p],
    'vkocr_head_tail_fwd': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_void_p, c_ll, c_ll, c_void_p],
    'vkocr_head_tail_bwd': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                            c_void_p, c_ll, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_head_tail_bwd_points': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_ll,
                                   c_void_p, c_ll, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_head_tail_fwd_points': [c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_ll,
                                   c_void_p, c_ll, c_void_p],
    'vkocr_points_claim': [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'vkocr_gather_up_taps': [c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    'vkocr_scatter_up_taps': [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_ll, c_void_p],
    'vkocr_rough_loss_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p],
    'vkocr_rough_loss_bwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_precise_loss_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_precise_loss_bwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_pack_weight': [c_void_p, c_ll, c_ll, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_ll, c_ll,
                          c_void_p],
    'vkocr_unpack_grad': [c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_ll, c_void_p],
    'vkocr_mlp2_grad_finalize': [c_void_p, c_ll, c_float, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p],
    'vkocr_accumulate_f32': [c_void_p, c_void_p, c_ll, c_void_p],
    'vkocr_set_columns': [c_int, c_void_p, c_ll, c_ll, c_int, c_int, c_float, c_void_p],
    'vkocr_zero': [c_void_p, c_ll, c_void_p],
    'vkocr_scale_rows': [c_int, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_int, c_void_p],
    'vkocr_ingest_image_u8': [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    'vkocr_rough_postprocess': [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p],
    'vkocr_precise_postprocess': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p],
    'vkocr_peak_mask': [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p],
    'vkocr_sumsq_f32': [c_void_p, c_ll, c_void_p, c_void_p],
    'vkocr_adamw_step': [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                         c_void_p, c_float, c_float, c_void_p],
    'vkocr_pointwise_loss_fwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_void_p, c_void_p, c_void_p],
    'vkocr_pointwise_loss_bwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                 c_void_p],
    'vkocr_soft_ce_fwd': [c_void_p, c_void_p, c_ll, c_int, c_ll, c_void_p, c_void_p, c_void_p],
    'vkocr_soft_ce_bwd': [c_void_p, c_void_p, c_ll, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p],
    'vkocr_hard_negative_bce_fwd': [c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p],
    'vkocr_hard_negative_bce_bwd': [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p],
}


def _load() -> ctypes.CDLL:
    path = _build.LIB
    if not os.path.exists(path) or os.environ.get('VKOCR_B200_REBUILD') == '1':
        path = _build.build()
    if os.environ.get('VKOCR_B200_LIB'):        # development: A/B a differently built copy of the same ABI
        path = os.environ['VKOCR_B200_LIB']
    lib = ctypes.CDLL(path)
    lib.vkocr_last_error.restype = ctypes.c_char_p
    lib.vkocr_launch_count.restype = ctypes.c_longlong
    lib.vkocr_launch_count.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = c_int
    return lib


class _Profile:
    """Optional CUDA-event bracketing of C-ABI calls on the current stream (bench.py / tools): per-call device time,
    grouped by a label the operator layer attaches (entry point + shape)."""

    def __init__(self) -> None:
        self.active = False
        self.only = None
        self.records = []     # (entry point, label, flops, bytes, start event, end event)
        self.pending = None   # (label, flops, bytes) noted by the operator layer for the next call

    def note(self, label: str, flops: float = 0.0, nbytes: float = 0.0) -> None:
        if self.active:
            self.pending = (label, flops, nbytes)

    def summary(self):
        """{label: dict(entry, calls, ms, flops, bytes)} — call after torch.cuda.synchronize()."""
        out = {}
        for entry, label, flops, nbytes, e0, e1 in self.records:
            row = out.setdefault(label, {'entry': entry, 'calls': 0, 'ms': 0.0, 'flops': 0.0, 'bytes': 0.0})
            row['calls'] += 1
            row['ms'] += e0.elapsed_time(e1)
            row['flops'] += flops
            row['bytes'] += nbytes
        return out


PROFILE = _Profile()


class _Library:
    """The loaded C ABI.  Entry points are plain attributes (raw ctypes functions); while profiling they are swapped for
    wrappers that bracket the call with CUDA events on the current stream."""

    def __init__(self, cdll: ctypes.CDLL) -> None:
        self._cdll = cdll
        self._raw = {name: getattr(cdll, name) for name in list(_SIGNATURES) + ['vkocr_last_error', 'vkocr_launch_count']}
        self._install(self._raw)

    def _install(self, table) -> None:
        for name, fn in table.items():
            setattr(self, name, fn)

    def start_profile(self, only=None) -> None:
        PROFILE.active, PROFILE.only, PROFILE.records, PROFILE.pending = True, only, [], None
        wrapped = {}
        for name, fn in self._raw.items():
            if name in ('vkocr_last_error', 'vkocr_launch_count', 'vkocr_abi_version', 'vkocr_device_check'):
                continue
            if only is not None and name not in only:
                continue
            wrapped[name] = self._wrap(name, fn)
        self._install(wrapped)

    def stop_profile(self):
        self._install(self._raw)
        PROFILE.active = False
        return PROFILE

    @staticmethod
    def _wrap(name, fn):
        def call(*args):
            note, PROFILE.pending = PROFILE.pending, None
            label, flops, nbytes = note if note is not None else (name, 0.0, 0.0)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            PROFILE.records.append((name, label, flops, nbytes, e0, e1))
            return rc
        return call


LIB = _Library(_load())


class VkocrError(RuntimeError):
    pass


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = LIB.vkocr_last_error().decode('utf-8', 'replace')
        raise VkocrError(f'vkocr_b200 {what} failed with status {rc}: {msg}')


def ptr(t: Optional[torch.Tensor]) -> c_void_p:
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device: Optional[torch.device] = None) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VkocrError(
                'vkocr_b200 operators run only on CUDA (sm_100a) tensors; there is no CPU fallback '
                f'(got a tensor on {t.device})')
