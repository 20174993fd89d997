"""Host-side operators of the adaptive-scaling hot path: thin Python over the C ABI of ``libvkocr_b200.so``.

Every compute step is a hand-written sm_100a kernel reached through ``_lib.LIB``; PyTorch here only owns device
memory, streams and the autograd tape (``torch.autograd.Function`` nodes whose forward/backward call the C ABI).

Activation layout: logical ``(B, C, H, W)`` tensors whose memory is NHWC ("channels last") with a pixel stride
``ld >= C`` that is a multiple of 8 elements, so that a channel vector is 16 bytes and TMA strides are legal.
Parameter gradients are accumulated by the kernels straight into ``param.grad`` (fp32), which is what the
data-parallel bucket views alias (see ``parallel.py``).
"""
import ctypes
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor
LN_EPS = 1e-6  # reference helper.ln: nn.LayerNorm(C, eps=1e-6) (model/helper.py:96-97)

BILINEAR = 0
NEAREST = 1


def _ceil_to(v: int, m: int) -> int:
    return (v + m - 1) // m * m


# ----------------------------------------------------------------------------------------------------- layout helpers
def alloc_nhwc(B: int, H: int, W: int, C: int, dtype: torch.dtype, device, zero: bool = False, ld: Optional[int] = None) -> Tensor:
    """(B, C, H, W)-shaped view of a fresh NHWC buffer with pixel stride ``ld`` (default: C rounded up to 8)."""
    ld = _ceil_to(C, 8) if ld is None else ld
    buf = (torch.zeros if zero else torch.empty)((B, H, W, ld), dtype=dtype, device=device)
    return buf[..., :C].permute(0, 3, 1, 2)


def is_nhwc(x: Tensor) -> bool:
    if x.dim() != 4:
        return False
    B, C, H, W = x.shape
    ld = x.stride(3)
    if C > 1 and x.stride(1) != 1:
        return False
    if ld < C or ld % 8 != 0:
        return False
    if W > 1 and False:
        return False
    ok = (H == 1 or x.stride(2) == W * ld) and (B == 1 or x.stride(0) == H * W * ld)
    return bool(ok) and x.data_ptr() % 16 == 0


def to_nhwc(x: Tensor, dtype: torch.dtype) -> Tensor:
    """Boundary conversion of a caller's (B,C,H,W) tensor to the internal layout / storage dtype."""
    L.require_cuda(x)
    if x.dtype == dtype and is_nhwc(x):
        return x
    B, C, H, W = x.shape
    out = alloc_nhwc(B, H, W, C, dtype, x.device)
    out.copy_(x)  # boundary plumbing only (layout/dtype change of caller-provided tensors)
    return out


def geom(x: Tensor) -> Tuple[int, int, int, int, int]:
    B, C, H, W = x.shape
    return B, H, W, C, x.stride(3)


_FROZEN_SCRATCH: Dict[Tuple, Tensor] = {}


def grad_buffer(p: Tensor) -> Tensor:
    """fp32 ``.grad`` of a parameter, created zeroed on first use; kernels accumulate into it.

    The backward nodes write parameter gradients here themselves and return ``None`` for them to autograd (that is what
    lets the gradients land in the flat data-parallel buckets without a copy), so ``torch.autograd.grad``, tensor hooks
    and ``create_graph`` do not see parameter gradients — use ``loss.backward()`` and read ``param.grad``.  A frozen
    parameter (``requires_grad_(False)``, e.g. backbone fine-tuning) gets no ``.grad``: its kernels accumulate into a
    shared scratch buffer that nothing reads."""
    if not p.requires_grad:
        key = (p.numel(), str(p.device))
        buf = _FROZEN_SCRATCH.get(key)
        if buf is None:
            buf = _FROZEN_SCRATCH[key] = torch.zeros(p.numel(), dtype=torch.float32, device=p.device)
        return buf.view(p.shape)
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
    g = p.grad
    if g.dtype != torch.float32 or not g.is_contiguous():
        raise L.VkocrError('vkocr_b200 accumulates parameter gradients into contiguous fp32 .grad buffers')
    return g


# Called with the parameters whose .grad has just been accumulated by a backward node (the kernels write .grad
# directly, so autograd's own post-accumulate hooks never fire for them).  ``parallel.DataParallel`` installs its
# bucket bookkeeping here to launch the gradient all-reduce while the rest of backward is still running.
_GRAD_READY_HOOK = None


def set_grad_ready_hook(fn) -> None:
    global _GRAD_READY_HOOK
    _GRAD_READY_HOOK = fn


def _ready(*params: Tensor) -> None:
    if _GRAD_READY_HOOK is not None:
        _GRAD_READY_HOOK(params)


def _s() -> ctypes.c_void_p:
    return L.stream_ptr()


def _tag(dtype: torch.dtype) -> int:
    return L.dtype_tag(dtype)


# ----------------------------------------------------------------------------------------------------- packed weights
class _PackCache:
    """Kernel-layout copies of the fp32 master weights, refreshed when the parameter's version counter moves."""

    def __init__(self) -> None:
        self.entries: Dict[Tuple, Tuple[Tuple, Tensor, Tuple]] = {}
        self.epoch = 0   # bumped by writers that update parameters outside torch (training.FusedAdamW): invalidates everything

    def bump(self) -> None:
        self.epoch += 1

    def get(self, key: Tuple, params: Sequence[Tensor], shape: Tuple[int, ...], dtype: torch.dtype, fill) -> Tensor:
        version = (self.epoch,) + tuple(int(p._version) for p in params) + tuple(int(p.data_ptr()) for p in params)
        ent = self.entries.get(key)
        # keys carry id(param); a dead parameter's id (and even its storage address) can be reused by a new one, so an
        # entry is only valid while the very same parameter objects are alive
        alive = ent is not None and all(r() is p for r, p in zip(ent[2], params))
        if alive and ent[0] == version and ent[1].device == params[0].device:
            return ent[1]
        if alive and ent[1].shape == shape and ent[1].dtype == dtype and ent[1].device == params[0].device:
            buf = ent[1]
        else:
            buf = torch.zeros(shape, dtype=dtype, device=params[0].device)
        fill(buf)
        if len(self.entries) > 4096:   # entries of parameters that no longer exist
            self.entries = {k: e for k, e in self.entries.items() if all(r() is not None for r in e[2])}
        self.entries[key] = (version, buf, tuple(weakref.ref(p) for p in params))
        return buf

    def clear(self) -> None:
        self.entries.clear()


PACK = _PackCache()


def _pack(w: Tensor, s_row: int, s_tap: int, s_col: int, rows: int, taps: int, cols: int, flip: int,
          col_scale: Optional[Tensor], out: Tensor, out_offset: int, o_row: int, o_tap: int) -> None:
    esz = out.element_size()
    L.check(L.LIB.vkocr_pack_weight(L.ptr(w), s_row, s_tap, s_col, rows, taps, cols, flip, L.ptr(col_scale),
                                    ctypes.c_void_p(out.data_ptr() + out_offset * esz), _tag(out.dtype), o_row, o_tap, _s()),
            'pack_weight')


def packed_linear_fwd(w: Tensor, dtype: torch.dtype) -> Tuple[Tensor, int]:
    """Linear weight (N, K) -> [N, c_pad] K-major operand."""
    N, K = w.shape
    c_pad = _ceil_to(K, 64)
    buf = PACK.get(('lin_fwd', id(w), dtype), [w], (N, c_pad), dtype,
                   lambda b: _pack(w.detach(), K, 0, 1, N, 1, K, 0, None, b, 0, c_pad, 0))
    return buf, c_pad


def packed_linear_dgrad(w: Tensor, dtype: torch.dtype, col_scale: Optional[Tensor] = None) -> Tuple[Tensor, int]:
    """Linear weight (N, K) -> [K, n_pad] operand of dX = dY . W (optionally with W rows scaled by col_scale[n])."""
    N, K = w.shape
    n_pad = _ceil_to(N, 64)
    # `col_scale` is the layer-scale PARAMETER itself (any shape with N elements): the cache entry lives as long as the
    # parameter objects do, so a per-call detached view would force a re-pack on every call
    params = [w] if col_scale is None else [w, col_scale]
    buf = PACK.get(('lin_dgrad', id(w), dtype, col_scale is not None), params, (K, n_pad), dtype,
                   lambda b: _pack(w.detach(), 1, 0, K, K, 1, N, 0,
                                   None if col_scale is None else col_scale.detach().reshape(-1), b, 0, n_pad, 0))
    return buf, n_pad


def packed_conv_fwd(ws: Sequence[Tensor], dtype: torch.dtype, n_slot: Optional[int] = None) -> Tuple[Tensor, int, int]:
    """Conv2d weights (N_h, C, k, k) of one or several heads -> [sum slots, k*k, c_pad]; each head gets ``n_slot`` rows."""
    C, k = ws[0].shape[1], ws[0].shape[2]
    T = k * k
    c_pad = _ceil_to(C, 64)
    slot = n_slot if n_slot is not None else ws[0].shape[0]
    rows = slot * len(ws)

    def fill(b: Tensor) -> None:
        for h, w in enumerate(ws):
            _pack(w.detach(), C * T, 1, T, w.shape[0], T, C, 0, None, b, h * slot * T * c_pad, T * c_pad, c_pad)

    buf = PACK.get(('conv_fwd', tuple(id(w) for w in ws), dtype, slot), list(ws), (rows, T * c_pad), dtype, fill)
    return buf, c_pad, rows


def packed_conv_dgrad(ws: Sequence[Tensor], dtype: torch.dtype, n_slot: Optional[int] = None) -> Tuple[Tensor, int]:
    """Conv2d weights -> [C, k*k (mirrored), n_pad] operand of the data gradient (a 'same' conv of dY with flipped taps)."""
    C, k = ws[0].shape[1], ws[0].shape[2]
    T = k * k
    slot = n_slot if n_slot is not None else ws[0].shape[0]
    n_pad = _ceil_to(slot * len(ws), 64)

    def fill(b: Tensor) -> None:
        for h, w in enumerate(ws):
            _pack(w.detach(), T, 1, C * T, C, T, w.shape[0], 1, None, b, h * slot, T * n_pad, n_pad)

    buf = PACK.get(('conv_dgrad', tuple(id(w) for w in ws), dtype, slot), list(ws), (C, T * n_pad), dtype, fill)
    return buf, n_pad


def packed_tapsplit_fwd(ws: Sequence[Tensor], dtype: torch.dtype, slot: int) -> Tuple[Tensor, int, int]:
    """Conv2d weights (N_h, C, k, k) of the heads of a group -> [k*k * ntot, c_pad] operand of Z = X . W_tap^T with rows
    ordered tap * ntot + head * slot + n (see csrc/head_combine.cu)."""
    C, k = ws[0].shape[1], ws[0].shape[2]
    T = k * k
    c_pad = _ceil_to(C, 64)
    ntot = slot * len(ws)

    def fill(b: Tensor) -> None:
        for h, w in enumerate(ws):
            _pack(w.detach(), C * T, 1, T, w.shape[0], T, C, 0, None, b, h * slot * c_pad, c_pad, ntot * c_pad)

    buf = PACK.get(('tapsplit_fwd', tuple(id(w) for w in ws), dtype, slot), list(ws), (T * ntot, c_pad), dtype, fill)
    return buf, c_pad, T * ntot


def packed_tapsplit_dgrad(ws: Sequence[Tensor], dtype: torch.dtype, slot: int) -> Tuple[Tensor, int]:
    """Conv2d weights -> [C, k_pad] operand of dX = dZ . W with columns ordered tap * ntot + head * slot + n."""
    C, k = ws[0].shape[1], ws[0].shape[2]
    T = k * k
    ntot = slot * len(ws)
    k_pad = _ceil_to(T * ntot, 64)

    def fill(b: Tensor) -> None:
        for h, w in enumerate(ws):
            _pack(w.detach(), T, 1, C * T, C, T, w.shape[0], 0, None, b, h * slot, k_pad, ntot)

    buf = PACK.get(('tapsplit_dgrad', tuple(id(w) for w in ws), dtype, slot), list(ws), (C, k_pad), dtype, fill)
    return buf, k_pad


def packed_tapsplit_points_dgrad(ws: Sequence[Tensor], dtype: torch.dtype, slot: int) -> Tuple[Tensor, int]:
    """Conv2d weights -> [k*k * C, k_pad] operand of U = G . W over the label-point rows (csrc/head_sparse.cu): row
    tap * C + c, column head * slot + n."""
    C, k = ws[0].shape[1], ws[0].shape[2]
    T = k * k
    k_pad = _ceil_to(slot * len(ws), 64)

    def fill(b: Tensor) -> None:
        for h, w in enumerate(ws):
            _pack(w.detach(), T, 1, C * T, C, T, w.shape[0], 0, None, b, h * slot, k_pad, C * k_pad)

    buf = PACK.get(('tapsplit_points_dgrad', tuple(id(w) for w in ws), dtype, slot), list(ws), (T * C, k_pad), dtype, fill)
    return buf, k_pad


# Upstream gradients that are zero outside a list of label points: data_ptr of the gradient map -> (label_y, label_x, (B, H, W)).
# PreciseLossFn.backward registers the maps it scatters into; HeadGroupFn.backward consumes the entry of a head whose incoming
# gradient is that very tensor (any accumulation or copy in between gives a new tensor and the dense path).
SPARSE_GRADS: Dict[int, Tuple[Tensor, Tensor, Tuple[int, int, int]]] = {}
SPARSE_BACKWARD = True     # tests set this to False to run the dense backward everywhere (cross-check)


def packed_patch_fwd(w: Tensor, dtype: torch.dtype) -> Tuple[Tensor, int]:
    """Patchify conv weight (N, C, p, p), stride p -> [N, c_pad] with k = (ky*p+kx)*C + c."""
    N, C, p, _ = w.shape
    T = p * p
    c_pad = _ceil_to(T * C, 64)
    buf = PACK.get(('patch_fwd', id(w), dtype), [w], (N, c_pad), dtype,
                   lambda b: _pack(w.detach(), C * T, 1, T, N, T, C, 0, None, b, 0, c_pad, C))
    return buf, c_pad


def packed_patch_dgrad(w: Tensor, dtype: torch.dtype) -> Tuple[Tensor, int]:
    """Patchify conv weight (N, C, p, p) -> [(t*C + c), n_pad] operand of dA = dY . W."""
    N, C, p, _ = w.shape
    T = p * p
    n_pad = _ceil_to(N, 64)
    buf = PACK.get(('patch_dgrad', id(w), dtype), [w], (T * C, n_pad), dtype,
                   lambda b: _pack(w.detach(), T, 1, C * T, C, T, N, 0, None, b, 0, n_pad, C * n_pad))
    return buf, n_pad


def packed_dwconv(w: Tensor, flip: bool) -> Tensor:
    """Depthwise weight (C,1,7,7) -> fp32 tap table [49][C] (mirrored for the data gradient)."""
    C = w.shape[0]
    return PACK.get(('dw', id(w), flip), [w], (49, C), torch.float32,
                    lambda b: _pack(w.detach(), 0, 1, 49, 1, 49, C, 1 if flip else 0, None, b, 0, 0, C))


# ----------------------------------------------------------------------------------------------------- raw kernel calls
def _epilogue(out: Tensor, ldo: int, *, out_f32: bool = False, accumulate: bool = False, out_pre: Optional[Tensor] = None,
              ld_pre: int = 0, bias: Optional[Tensor] = None, act: int = 0, col_scale: Optional[Tensor] = None,
              row_scale: Optional[Tensor] = None, rows_per_group: int = 1, residual: Optional[Tensor] = None, ld_res: int = 0,
              aux: Optional[Tensor] = None, ld_aux: int = 0, tn: Tuple[int, int, int] = (0, 0, 0)) -> L.Epilogue:
    return L.Epilogue(out.data_ptr(), ldo, int(out_f32), int(accumulate),
                      None if out_pre is None else out_pre.data_ptr(), ld_pre,
                      None if bias is None else bias.data_ptr(), act,
                      None if col_scale is None else col_scale.data_ptr(),
                      None if row_scale is None else row_scale.data_ptr(), rows_per_group,
                      None if residual is None else residual.data_ptr(), ld_res,
                      None if aux is None else aux.data_ptr(), ld_aux, tn[0], tn[1], tn[2])


TAPSPLIT = True     # heads with upsampling_factor >= 2: convolve on the low-resolution map, resample after (head_combine.cu)
COMBINE_ALGO = 0    # tests set this to 1 to force the generic gather kernels of head_combine.cu (cross-check of the tiled ones)
SIMT_BACKEND = 0  # tests set this to 1 to route bf16 GEMMs through the SIMT kernel (cross-check of the tcgen05 path)


def gemm_nt(x: Tensor, B: int, H: int, W: int, C: int, ld_x: int, ks: int, wp: Tensor, c_pad: int, N: int, ep: L.Epilogue,
            alg_kn: Optional[int] = None) -> None:
    """``alg_kn``: the algorithmic K*N of the contraction when C / N carry layout padding (profile bookkeeping only)."""
    if L.PROFILE.active:
        M = B * H * W
        L.PROFILE.note(f'gemm_nt ks{ks} M{M} K{ks * ks * C} N{N}', 2.0 * M * (ks * ks * C * N if alg_kn is None else alg_kn))
    g = L.ConvGeom(B, H, W, ks, C, ld_x, c_pad)
    L.check(L.LIB.vkocr_gemm_nt(_tag(x.dtype), SIMT_BACKEND, L.ptr(x), ctypes.byref(g), L.ptr(wp), N, ctypes.byref(ep), _s()), 'gemm_nt')


def gemm_tn(p: Tensor, B: int, H: int, W: int, I: int, ld_p: int, ks: int, q: Tensor, J: int, ld_q: int, ep: L.Epilogue,
            alg_ij: Optional[int] = None) -> None:
    if L.PROFILE.active:
        M = B * H * W
        L.PROFILE.note(f'gemm_tn ks{ks} M{M} I{I} J{J}', 2.0 * M * (ks * ks * I * J if alg_ij is None else alg_ij))
    g = L.ConvGeom(B, H, W, ks, I, ld_p, 0)
    L.check(L.LIB.vkocr_gemm_tn(_tag(p.dtype), SIMT_BACKEND, L.ptr(p), ctypes.byref(g), L.ptr(q), J, ld_q, ctypes.byref(ep), _s()), 'gemm_tn')


def layernorm_fwd(x: Tensor, ld_x: int, y: Tensor, ld_y: int, rows: int, C: int, gamma: Tensor, beta: Tensor, act: int,
                  mean: Optional[Tensor], rstd: Optional[Tensor]) -> None:
    if L.PROFILE.active:
        L.PROFILE.note(f'layernorm_fwd rows{rows} C{C} act{act}', 0.0, 2.0 * rows * C * x.element_size())
    L.check(L.LIB.vkocr_layernorm_fwd(_tag(x.dtype), L.ptr(x), ld_x, L.ptr(y), ld_y, rows, C, L.ptr(gamma), L.ptr(beta), LN_EPS, act,
                                      L.ptr(mean), L.ptr(rstd), _s()), 'layernorm_fwd')


def layernorm_bwd(dy: Tensor, ld_dy: int, x: Tensor, ld_x: int, mean: Tensor, rstd: Tensor, gamma: Tensor, beta: Tensor, act: int,
                  dx: Tensor, ld_dx: int, rows: int, C: int, dgamma: Optional[Tensor], dbeta: Optional[Tensor],
                  dxsum: Optional[Tensor]) -> None:
    if L.PROFILE.active:
        L.PROFILE.note(f'layernorm_bwd rows{rows} C{C} act{act}', 0.0, 3.0 * rows * C * x.element_size())
    L.check(L.LIB.vkocr_layernorm_bwd(_tag(x.dtype), L.ptr(dy), ld_dy, L.ptr(x), ld_x, L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                      L.ptr(beta), act, L.ptr(dx), ld_dx, rows, C, L.ptr(dgamma), L.ptr(dbeta), L.ptr(dxsum), _s()),
            'layernorm_bwd')


def colsum(x: Tensor, ld: int, rows: int, C: int, out: Tensor) -> None:
    if L.PROFILE.active:
        L.PROFILE.note(f'colsum rows{rows} C{C}', 0.0, 1.0 * rows * C * x.element_size())
    L.check(L.LIB.vkocr_colsum(_tag(x.dtype), L.ptr(x), ld, rows, C, L.ptr(out), _s()), 'colsum')


def dwconv7(x: Tensor, y: Tensor, wt: Tensor, bias: Optional[Tensor], add: Optional[Tensor]) -> None:
    B, H, W, C, ld_x = geom(x)
    if L.PROFILE.active:
        n = B * H * W * C
        L.PROFILE.note(f'dwconv7 {B}x{H}x{W}x{C}{" +add" if add is not None else ""}', 98.0 * n,
                       (3.0 if add is not None else 2.0) * n * x.element_size())
    L.check(L.LIB.vkocr_dwconv7_fwd(_tag(x.dtype), L.ptr(x), ld_x, L.ptr(y), y.stride(3), B, H, W, C, L.ptr(wt), L.ptr(bias),
                                    L.ptr(add), 0 if add is None else add.stride(3), _s()), 'dwconv7_fwd')


def dwconv7_wgrad(dy: Tensor, x: Tensor, dw: Tensor) -> None:
    B, H, W, C, ld_x = geom(x)
    if L.PROFILE.active:
        n = B * H * W * C
        L.PROFILE.note(f'dwconv7_wgrad {B}x{H}x{W}x{C}', 98.0 * n, 2.0 * n * x.element_size())
    L.check(L.LIB.vkocr_dwconv7_wgrad(_tag(x.dtype), L.ptr(dy), dy.stride(3), L.ptr(x), ld_x, B, H, W, C, L.ptr(dw), _s()),
            'dwconv7_wgrad')


def upsample_fwd(src: Tensor, dst: Tensor, C: int, mode: int, accumulate: bool) -> None:
    """dst[:, :C] (=|+=) resample(src[:, :C]); src/dst are NHWC views (possibly channel slices of wider buffers)."""
    B, _, h, w = src.shape
    _, _, H, W = dst.shape
    if L.PROFILE.active:
        L.PROFILE.note(f'upsample_fwd {B}x{h}x{w}->{H}x{W} C{C} mode{mode} acc{int(accumulate)}', 0.0,
                       (B * h * w + (2 if accumulate else 1) * B * H * W) * C * src.element_size())
    L.check(L.LIB.vkocr_upsample_fwd(_tag(src.dtype), L.ptr(src), src.stride(3), h, w, L.ptr(dst), dst.stride(3), H, W, B, C, mode,
                                     int(accumulate), _s()), 'upsample_fwd')


def upsample_bwd(ddst: Tensor, dsrc: Tensor, C: int, mode: int, accumulate: bool) -> None:
    B, _, h, w = dsrc.shape
    _, _, H, W = ddst.shape
    if L.PROFILE.active:
        L.PROFILE.note(f'upsample_bwd {B}x{H}x{W}->{h}x{w} C{C} mode{mode}', 0.0, (B * h * w + B * H * W) * C * ddst.element_size())
    V = 16 // ddst.element_size()
    if (H >= 3 * h or W >= 3 * w) and C % V == 0 and ddst.stride(3) % V == 0 and dsrc.stride(3) % V == 0 \
            and ddst.data_ptr() % 16 == 0 and dsrc.data_ptr() % 16 == 0:
        # large scale factors (PPM / top-down paths): two separable passes through an fp32 workspace [B, H, w, C]
        ws = torch.empty(B * H * w * C, dtype=torch.float32, device=ddst.device)
        L.check(L.LIB.vkocr_upsample_bwd_separable(_tag(ddst.dtype), L.ptr(ddst), ddst.stride(3), H, W, L.ptr(dsrc), dsrc.stride(3),
                                                   h, w, B, C, mode, int(accumulate), L.ptr(ws), _s()), 'upsample_bwd_separable')
        return
    L.check(L.LIB.vkocr_upsample_bwd(_tag(ddst.dtype), L.ptr(ddst), ddst.stride(3), H, W, L.ptr(dsrc), dsrc.stride(3), h, w, B, C,
                                     mode, int(accumulate), _s()), 'upsample_bwd')


def copy_channels(src: Tensor, dst: Tensor, C: int, accumulate: bool = False) -> None:
    B, _, H, W = src.shape
    L.check(L.LIB.vkocr_copy_channels(_tag(src.dtype), L.ptr(src), src.stride(3), L.ptr(dst), dst.stride(3), B * H * W, C,
                                      int(accumulate), _s()), 'copy_channels')


_ONES_TAIL = {}


def _ones_tail(dtype: torch.dtype, device) -> Tensor:
    """[1, 0, 0, 0, 0, 0, 0, 0] in the storage dtype (cached per device)."""
    key = (dtype, str(device))
    if key not in _ONES_TAIL:
        t = torch.zeros(8, dtype=dtype, device=device)
        t[0] = 1
        _ONES_TAIL[key] = t
    return _ONES_TAIL[key]


import os as _os
_TORCH_ZERO = _os.environ.get('VKOCR_TORCH_ZERO') == '1'   # A/B switch for experiments


def zero_(t: Tensor) -> Tensor:
    """Asynchronous zero fill on the current stream through the C ABI (no eager PyTorch kernel on the hot path)."""
    if _TORCH_ZERO:
        return t.zero_()
    L.check(L.LIB.vkocr_zero(L.ptr(t), t.numel() * t.element_size(), _s()), 'zero')
    return t


def _zeros_f32(n: int, device) -> Tensor:
    return zero_(torch.empty(n, dtype=torch.float32, device=device))


def _zeros(n: int, dtype: torch.dtype, device) -> Tensor:
    return zero_(torch.empty(n, dtype=dtype, device=device))


_GRAD_MODE = [True]      # grad mode of the caller of the innermost Function.apply in flight (forward itself always runs with it off)


class _Function(torch.autograd.Function):
    """``torch.autograd.Function`` that remembers the caller's grad mode: ``ctx.needs_input_grad`` only says which inputs
    require grad -- it is the same under ``torch.no_grad()`` -- and a forward that takes it for "a backward will follow"
    writes every training side channel (gelu' maps, LayerNorm statistics, conv outputs) during inference."""

    @classmethod
    def apply(cls, *args, **kwargs):
        _GRAD_MODE.append(torch.is_grad_enabled())
        try:
            return super().apply(*args, **kwargs)
        finally:
            _GRAD_MODE.pop()


def _needs_grad(ctx) -> bool:
    """True when autograd will call this node's backward: grad mode was on at ``apply`` and some input requires grad."""
    return _GRAD_MODE[-1] and any(ctx.needs_input_grad)


# ----------------------------------------------------------------------------------------------------- autograd nodes
class StemFn(_Function):
    """pconv(p x p, stride p) + bias -> LayerNorm, image NCHW fp32 -> NHWC (convnext.py:106-123, helper.py:43-58)."""

    @staticmethod
    def forward(ctx, image: Tensor, w: Tensor, b: Tensor, ln_w: Tensor, ln_b: Tensor, dtype: torch.dtype) -> Tensor:
        L.require_cuda(image, w)
        image = image.contiguous().float()
        B, Cin, H, W = image.shape
        N, _, p, _ = w.shape
        Ho, Wo = H // p, W // p
        M = B * Ho * Wo
        wp, c_pad = packed_patch_fwd(w, dtype)
        a = torch.empty((M, c_pad), dtype=dtype, device=image.device)
        L.check(L.LIB.vkocr_patchify_image(_tag(dtype), L.ptr(image), B, Cin, H, W, p, L.ptr(a), c_pad, _s()), 'patchify_image')
        pre = alloc_nhwc(B, Ho, Wo, N, dtype, image.device)
        gemm_nt(a, 1, 1, M, c_pad, c_pad, 1, wp, c_pad, N, _epilogue(pre, pre.stride(3), bias=b.detach()))
        y = alloc_nhwc(B, Ho, Wo, N, dtype, image.device)
        train = _needs_grad(ctx)
        mean = torch.empty(M, dtype=torch.float32, device=image.device) if train else None
        rstd = torch.empty(M, dtype=torch.float32, device=image.device) if train else None
        layernorm_fwd(pre, pre.stride(3), y, y.stride(3), M, N, ln_w.detach(), ln_b.detach(), 0, mean, rstd)
        if train:
            ctx.save_for_backward(a, pre, mean, rstd, w, b, ln_w, ln_b)
            ctx.meta = (M, N, c_pad, Cin, p)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        a, pre, mean, rstd, w, b, ln_w, ln_b = ctx.saved_tensors
        M, N, c_pad, Cin, p = ctx.meta
        dy = to_nhwc(dy, pre.dtype)
        dpre = torch.empty_like(pre)
        layernorm_bwd(dy, dy.stride(3), pre, pre.stride(3), mean, rstd, ln_w.detach(), ln_b.detach(), 0, dpre, dpre.stride(3), M, N,
                      grad_buffer(ln_w), grad_buffer(ln_b), grad_buffer(b))
        K = p * p * Cin
        g = _zeros_f32(N * K, dy.device)
        gemm_tn(dpre, 1, 1, M, N, dpre.stride(3), 1, a, K, c_pad, _epilogue(g, K, out_f32=True, accumulate=True, tn=(0, K, 1)))
        # GEMM order k = (t, c)  ->  parameter layout (n, c, t)
        L.check(L.LIB.vkocr_unpack_grad(L.ptr(g), N, p * p, Cin, L.ptr(grad_buffer(w)), Cin * p * p, 1, p * p, _s()), 'unpack_grad')
        _ready(w, b, ln_w, ln_b)
        return None, None, None, None, None, None


class LayerNormFn(_Function):
    """Channels-last LayerNorm, optionally followed by exact GELU (helper.py:96-101)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, act: int) -> Tensor:
        B, H, W, C, ld = geom(x)
        y = alloc_nhwc(B, H, W, C, x.dtype, x.device)
        train = _needs_grad(ctx)
        M = B * H * W
        mean = torch.empty(M, dtype=torch.float32, device=x.device) if train else None
        rstd = torch.empty(M, dtype=torch.float32, device=x.device) if train else None
        layernorm_fwd(x, ld, y, y.stride(3), M, C, w.detach(), b.detach(), act, mean, rstd)
        if train:
            ctx.save_for_backward(x, mean, rstd, w, b)
            ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, mean, rstd, w, b = ctx.saved_tensors
        B, H, W, C, ld = geom(x)
        dy = to_nhwc(dy, x.dtype)
        dx = alloc_nhwc(B, H, W, C, x.dtype, x.device)
        layernorm_bwd(dy, dy.stride(3), x, ld, mean, rstd, w.detach(), b.detach(), ctx.act, dx, dx.stride(3), B * H * W, C,
                      grad_buffer(w), grad_buffer(b), None)
        _ready(w, b)
        return dx, None, None, None


class ConvNextLayerFn(_Function):
    """x + mask * scale * Linear2(GELU(Linear1(LN(dwconv7x7(x)))))  (ConvNextBlockLayer, convnext.py:20-59)."""

    @staticmethod
    def forward(ctx, x: Tensor, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, scale, drop_mask: Optional[Tensor],
                inv_keep: float = 1.0) -> Tensor:
        B, H, W, C, ld = geom(x)
        M = B * H * W
        dt, dev = x.dtype, x.device
        train = _needs_grad(ctx)
        conv = alloc_nhwc(B, H, W, C, dt, dev)
        dwconv7(x, conv, packed_dwconv(dw_w, False), dw_b.detach(), None)
        if train:
            # 8 extra channels per pixel, [1, 0 .. 0]: the weight-gradient GEMM dH^T . [LN_out | 1] then delivers the bias
            # gradient of the up-projection (column sums of dH) as column C of its product, for 8 % more J instead of a
            # separate pass over the (M, 4C) gradient
            lnbuf = torch.empty((B, H, W, C + 8), dtype=dt, device=dev)
            L.check(L.LIB.vkocr_set_columns(_tag(dt), L.ptr(lnbuf), C + 8, M, C, 8, 1.0, _s()), 'set_columns')
            lnout = lnbuf[..., :C].permute(0, 3, 1, 2)
        else:
            lnout = alloc_nhwc(B, H, W, C, dt, dev)
        mean = torch.empty(M, dtype=torch.float32, device=dev) if train else None
        rstd = torch.empty(M, dtype=torch.float32, device=dev) if train else None
        layernorm_fwd(conv, conv.stride(3), lnout, lnout.stride(3), M, C, ln_w.detach(), ln_b.detach(), 0, mean, rstd)
        hid = 4 * C
        w1p, c1 = packed_linear_fwd(w1, dt)
        # With a stochastic-depth mask m_b the up-projection epilogue zeroes G for dropped samples and stores m_b * gelu' in
        # the side channel, so the backward runs on dY itself (no masked copy of the gradient).
        g = torch.empty((M, hid), dtype=dt, device=dev)
        ldg = g.stride(0)
        # training: the second output is (m_b *) gelu'(H_pre) (act 3) -- all the backward ever needs of the pre-activation,
        # and it shares the transcendental work with the GELU itself
        hpre = torch.empty((M, hid), dtype=dt, device=dev) if train else None
        gemm_nt(lnout, 1, 1, M, C, lnout.stride(3), 1, w1p, c1, hid,
                _epilogue(g, ldg, out_pre=hpre, ld_pre=hid, bias=b1.detach(), act=3 if train else 1,
                          row_scale=drop_mask if train else None, rows_per_group=H * W))
        w2p, c2 = packed_linear_fwd(w2, dt)
        y = alloc_nhwc(B, H, W, C, dt, dev)
        gamma = scale.detach().reshape(-1)
        gemm_nt(g, 1, 1, M, hid, ldg, 1, w2p, c2, C,
                _epilogue(y, y.stride(3), bias=b2.detach(), col_scale=gamma, row_scale=drop_mask, rows_per_group=H * W,
                          residual=x, ld_res=ld))
        if train:
            ctx.save_for_backward(x, conv, mean, rstd, lnout, hpre, g, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, scale,
                                  drop_mask if drop_mask is not None else torch.empty(0, device=dev))
            ctx.has_mask = drop_mask is not None
            ctx.inv_keep = float(inv_keep) if drop_mask is not None else 1.0
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, conv, mean, rstd, lnout, hpre, g, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, scale, mask = ctx.saved_tensors
        B, H, W, C, ld = geom(x)
        M = B * H * W
        hid = 4 * C
        dt, dev = x.dtype, x.device
        dy = to_nhwc(dy, dt)
        gamma = scale.detach().reshape(-1)
        # dH_pre = (dY . (gamma * W2)) * (m_b gelu'(H_pre))   (hpre holds m_b * gelu'(H_pre), written by the forward)
        w2d, n2 = packed_linear_dgrad(w2, dt, scale)
        dh = torch.empty((M, hid), dtype=dt, device=dev)
        gemm_nt(dy, 1, 1, M, C, dy.stride(3), 1, w2d, n2, hid, _epilogue(dh, hid, act=4, aux=hpre, ld_aux=hid))
        # sU[c] = sum_p m_b dY[p,c];  S[c,k] = sum_p dY[p,c] G[p,k] over the kept samples (x 1/p_keep in the finaliser)
        # one zero-fill for the three fp32 accumulators of this layer (sU, S and [dW1 | db1]): they are a few MB, the launches
        # are what costs
        ones_tail = lnout.stride(3) == C + 8
        acc = _zeros_f32(C + C * hid + (hid * (C + 8) if ones_tail else 0), dev)
        su = acc[:C]
        if ctx.has_mask:
            L.check(L.LIB.vkocr_scale_rows_colsum(_tag(dt), L.ptr(dy), dy.stride(3), None, 0, M, C, L.ptr(mask), H * W, L.ptr(su), _s()),
                    'scale_rows_colsum')
        else:
            colsum(dy, dy.stride(3), M, C, su)
        s = acc[C:C + C * hid]
        gemm_tn(dy, 1, 1, M, C, dy.stride(3), 1, g, hid, hid, _epilogue(s, hid, out_f32=True, accumulate=True, tn=(0, hid, 1)))
        L.check(L.LIB.vkocr_mlp2_grad_finalize(L.ptr(s), hid, ctx.inv_keep, L.ptr(su), 1, L.ptr(w2.detach()), L.ptr(b2.detach()),
                                               L.ptr(gamma), C, hid, L.ptr(grad_buffer(w2)), L.ptr(grad_buffer(scale)),
                                               L.ptr(grad_buffer(b2)), _s()),
                'mlp2_grad_finalize')
        del g
        ldl = lnout.stride(3)
        if ones_tail:             # LN_out carries the ones channel (see forward): dW1 and db1 from one GEMM
            gwb = acc[C + C * hid:]
            gemm_tn(dh, 1, 1, M, hid, hid, 1, lnout, C + 8, ldl, _epilogue(gwb, C + 8, out_f32=True, accumulate=True, tn=(0, C + 8, 1)))
            L.check(L.LIB.vkocr_scatter_add_f32(L.ptr(gwb), C + 8, 0, 1, hid, 1, C, L.ptr(grad_buffer(w1)), C, 0, 1, _s()), 'scatter_add_f32')
            L.check(L.LIB.vkocr_scatter_add_f32(ctypes.c_void_p(gwb.data_ptr() + 4 * C), C + 8, 0, 0, hid, 1, 1, L.ptr(grad_buffer(b1)), 1, 0, 0,
                                                _s()), 'scatter_add_f32')
        else:
            colsum(dh, hid, M, hid, grad_buffer(b1))
            gemm_tn(dh, 1, 1, M, hid, hid, 1, lnout, C, ldl, _epilogue(grad_buffer(w1), C, out_f32=True, accumulate=True, tn=(0, C, 1)))
        w1d, n1 = packed_linear_dgrad(w1, dt)
        dln = alloc_nhwc(B, H, W, C, dt, dev)
        gemm_nt(dh, 1, 1, M, hid, hid, 1, w1d, n1, C, _epilogue(dln, dln.stride(3)))
        del dh
        dconv = alloc_nhwc(B, H, W, C, dt, dev)
        layernorm_bwd(dln, dln.stride(3), conv, conv.stride(3), mean, rstd, ln_w.detach(), ln_b.detach(), 0, dconv, dconv.stride(3), M,
                      C, grad_buffer(ln_w), grad_buffer(ln_b), grad_buffer(dw_b))
        dwconv7_wgrad(dconv, x, grad_buffer(dw_w))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = alloc_nhwc(B, H, W, C, dt, dev)
            dwconv7(dconv, dx, packed_dwconv(dw_w, True), None, dy)   # + residual gradient
        _ready(dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, scale)
        return (dx,) + (None,) * 11


class PatchConvFn(_Function):
    """pconv2x2 (kernel 2, stride 2) on NHWC = space-to-depth gather + GEMM (helper.py:43-49, convnext.py:89-99)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor) -> Tensor:
        B, H, W, C, ld = geom(x)
        N = w.shape[0]
        if w.shape[2] != 2:
            raise L.VkocrError('vkocr_b200 PatchConvFn supports the 2x2/stride-2 down-sampling conv only')
        Ho, Wo = H // 2, W // 2
        M = B * Ho * Wo
        dt, dev = x.dtype, x.device
        wp, c_pad = packed_patch_fwd(w, dt)
        zero = c_pad != 4 * C
        a = (torch.zeros if zero else torch.empty)((M, c_pad), dtype=dt, device=dev)
        L.check(L.LIB.vkocr_space_to_depth2(_tag(dt), L.ptr(x), ld, B, H, W, C, L.ptr(a), c_pad, 0, 0, _s()), 'space_to_depth2')
        y = alloc_nhwc(B, Ho, Wo, N, dt, dev)
        gemm_nt(a, 1, 1, M, 4 * C, c_pad, 1, wp, c_pad, N, _epilogue(y, y.stride(3), bias=b.detach()))
        if _needs_grad(ctx):
            ctx.save_for_backward(a, w, b)
            ctx.meta = (B, H, W, C, N, c_pad)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        a, w, b = ctx.saved_tensors
        B, H, W, C, N, c_pad = ctx.meta
        dt, dev = a.dtype, a.device
        M = a.shape[0]
        dy = to_nhwc(dy, dt)
        colsum(dy, dy.stride(3), M, N, grad_buffer(b))
        K = 4 * C
        g = _zeros_f32(N * K, dev)
        gemm_tn(dy, 1, 1, M, N, dy.stride(3), 1, a, K, c_pad, _epilogue(g, K, out_f32=True, accumulate=True, tn=(0, K, 1)))
        L.check(L.LIB.vkocr_unpack_grad(L.ptr(g), N, 4, C, L.ptr(grad_buffer(w)), C * 4, 1, 4, _s()), 'unpack_grad')
        dx = None
        if ctx.needs_input_grad[0]:
            wd, n_pad = packed_patch_dgrad(w, dt)
            da = torch.empty((M, K), dtype=dt, device=dev)
            gemm_nt(dy, 1, 1, M, N, dy.stride(3), 1, wd, n_pad, K, _epilogue(da, K))
            dx = alloc_nhwc(B, H, W, C, dt, dev)
            L.check(L.LIB.vkocr_space_to_depth2(_tag(dt), L.ptr(dx), dx.stride(3), B, H, W, C, L.ptr(da), K, 1, 0, _s()),
                    'space_to_depth2(adjoint)')
        _ready(w, b)
        return dx, None, None


class ConvLnGeluFn(_Function):
    """k x k 'same' conv (k = 1 is the Linear of build_conv1x1_block) + bias -> LayerNorm -> GELU
    (upernext.py:21-45, fpn.py:21-48), as an implicit GEMM on the tensor cores."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, ln_w: Tensor, ln_b: Tensor) -> Tensor:
        B, H, W, C, ld = geom(x)
        dt, dev = x.dtype, x.device
        if w.dim() == 2:
            N, ks = w.shape[0], 1
            wp, c_pad = packed_linear_fwd(w, dt)
        else:
            N, ks = w.shape[0], w.shape[2]
            wp, c_pad, _ = packed_conv_fwd([w], dt)
        M = B * H * W
        pre = alloc_nhwc(B, H, W, N, dt, dev)
        if ks == 1:
            gemm_nt(x, 1, 1, M, C, ld, 1, wp, c_pad, N, _epilogue(pre, pre.stride(3), bias=b.detach()))
        else:
            gemm_nt(x, B, H, W, C, ld, ks, wp, c_pad, N, _epilogue(pre, pre.stride(3), bias=b.detach()))
        train = _needs_grad(ctx)
        mean = torch.empty(M, dtype=torch.float32, device=dev) if train else None
        rstd = torch.empty(M, dtype=torch.float32, device=dev) if train else None
        y = alloc_nhwc(B, H, W, N, dt, dev)
        layernorm_fwd(pre, pre.stride(3), y, y.stride(3), M, N, ln_w.detach(), ln_b.detach(), 1, mean, rstd)
        if train:
            ctx.save_for_backward(x, pre, mean, rstd, w, b, ln_w, ln_b)
            ctx.ks = ks
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, pre, mean, rstd, w, b, ln_w, ln_b = ctx.saved_tensors
        ks = ctx.ks
        B, H, W, C, ld = geom(x)
        N = w.shape[0]
        M = B * H * W
        dt, dev = x.dtype, x.device
        dy = to_nhwc(dy, dt)
        dpre = alloc_nhwc(B, H, W, N, dt, dev)
        layernorm_bwd(dy, dy.stride(3), pre, pre.stride(3), mean, rstd, ln_w.detach(), ln_b.detach(), 1, dpre, dpre.stride(3), M, N,
                      grad_buffer(ln_w), grad_buffer(ln_b), grad_buffer(b))
        gw = grad_buffer(w)
        T = ks * ks
        if ks == 1:
            gemm_tn(dpre, 1, 1, M, N, dpre.stride(3), 1, x, C, ld, _epilogue(gw, C, out_f32=True, accumulate=True, tn=(0, C, 1)))
        else:
            gemm_tn(dpre, B, H, W, N, dpre.stride(3), ks, x, C, ld,
                    _epilogue(gw, C, out_f32=True, accumulate=True, tn=(1, C * T, T)))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = alloc_nhwc(B, H, W, C, dt, dev)
            if ks == 1:
                wd, n_pad = packed_linear_dgrad(w, dt)
                gemm_nt(dpre, 1, 1, M, N, dpre.stride(3), 1, wd, n_pad, C, _epilogue(dx, dx.stride(3)))
            else:
                wd, n_pad = packed_conv_dgrad([w], dt)
                gemm_nt(dpre, B, H, W, N, dpre.stride(3), ks, wd, n_pad, C, _epilogue(dx, dx.stride(3)))
        _ready(w, b, ln_w, ln_b)
        return dx, None, None, None, None


class AvgPoolFn(_Function):
    """nn.AdaptiveAvgPool2d(S) on NHWC (upernext.py:59-65)."""

    @staticmethod
    def forward(ctx, x: Tensor, S: int) -> Tensor:
        B, H, W, C, ld = geom(x)
        y = alloc_nhwc(B, S, S, C, x.dtype, x.device)
        if H * W >= 64 * S * S:      # bins of >= 64 pixels: two separable passes instead of one serial sum per output
            ws = torch.empty(B * H * S * C, dtype=torch.float32, device=x.device)
            L.check(L.LIB.vkocr_avgpool_fwd_separable(_tag(x.dtype), L.ptr(x), ld, H, W, L.ptr(y), y.stride(3), S, B, C, L.ptr(ws),
                                                      _s()), 'avgpool_fwd_separable')
        else:
            L.check(L.LIB.vkocr_avgpool_fwd(_tag(x.dtype), L.ptr(x), ld, H, W, L.ptr(y), y.stride(3), S, B, C, _s()), 'avgpool_fwd')
        ctx.meta = (B, H, W, C, S)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        B, H, W, C, S = ctx.meta
        dy = to_nhwc(dy, dy.dtype)
        dx = alloc_nhwc(B, H, W, C, dy.dtype, dy.device)
        L.check(L.LIB.vkocr_avgpool_bwd(_tag(dy.dtype), L.ptr(dy), dy.stride(3), S, L.ptr(dx), dx.stride(3), H, W, B, C, 0, _s()),
                'avgpool_bwd')
        return dx, None


class UpsampleAddFn(_Function):
    """base + F.interpolate(src, size=base.shape[-2:], mode) — the top-down step (upernext.py:174-182, fpn.py:121-129)."""

    @staticmethod
    def forward(ctx, base: Tensor, src: Tensor, mode: int) -> Tensor:
        B, H, W, C, ld = geom(base)
        out = alloc_nhwc(B, H, W, C, base.dtype, base.device)
        copy_channels(base, out, C)
        upsample_fwd(src, out, C, mode, True)
        ctx.mode = mode
        ctx.src_shape = tuple(src.shape)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        dout = to_nhwc(dout, dout.dtype)
        B, C, h, w = ctx.src_shape
        dsrc = None
        if ctx.needs_input_grad[1]:
            dsrc = alloc_nhwc(B, h, w, C, dout.dtype, dout.device)
            upsample_bwd(dout, dsrc, C, ctx.mode, False)
        return (dout if ctx.needs_input_grad[0] else None), dsrc, None


class UpsampleConcatFn(_Function):
    """torch.cat([F.interpolate(t, size) or t ...], dim=1): every level is resampled straight into its channel slice of
    the NHWC concat buffer (upernext.py:76-82,189-197; fpn.py:136-144)."""

    @staticmethod
    def forward(ctx, mode: int, H: int, W: int, *xs: Tensor) -> Tensor:
        B = xs[0].shape[0]
        ctot = sum(int(t.shape[1]) for t in xs)
        out = alloc_nhwc(B, H, W, ctot, xs[0].dtype, xs[0].device)
        off = 0
        for t in xs:
            C = int(t.shape[1])
            dst = out[:, off:off + C]
            if t.shape[2] == H and t.shape[3] == W:
                copy_channels(t, dst, C)
            else:
                upsample_fwd(t, dst, C, mode, False)
            off += C
        ctx.mode = mode
        ctx.shapes = [tuple(t.shape) for t in xs]
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        dout = to_nhwc(dout, dout.dtype)
        H, W = dout.shape[2], dout.shape[3]
        grads: List[Optional[Tensor]] = []
        off = 0
        for i, (B, C, h, w) in enumerate(ctx.shapes):
            src = dout[:, off:off + C]
            off += C
            if not ctx.needs_input_grad[3 + i]:
                grads.append(None)
                continue
            if h == H and w == W:
                grads.append(src)   # a channel-slice view is a valid NHWC operand (ld = total width)
            else:
                d = alloc_nhwc(B, h, w, C, dout.dtype, dout.device)
                upsample_bwd(src, d, C, ctx.mode, False)
                grads.append(d)
        return (None, None, None, *grads)


class HeadGroupFn(_Function):
    """All heads that read one neck tensor, fused: x`factor` up-sample (once) -> one implicit-GEMM conv with the heads'
    k x k weights concatenated along N -> per-head LayerNorm+GELU+1x1(+Softplus) tail writing NCHW fp32 maps.
    (UperNextHead.forward upernext.py:233-248, FpnHead.forward fpn.py:193-208, Softplus adaptive_scaling.py:101,140.)

    ``params`` is a flat list of 6 tensors per head: conv_w, conv_b, ln_w, ln_b, lin_w, lin_b.
    """

    @staticmethod
    def forward(ctx, x: Tensor, factor: int, mode: int, softplus: Tuple[bool, ...], *params: Tensor):
        nh = len(softplus)
        heads = [params[6 * i:6 * i + 6] for i in range(nh)]
        B, h, w, C, ld = geom(x)
        dt, dev = x.dtype, x.device
        H, W = h * factor, w * factor
        inners = [int(hd[0].shape[0]) for hd in heads]
        ks = int(heads[0][0].shape[2])
        tapsplit = factor > 1 and TAPSPLIT and nh <= L.MAX_HEADS and max(inners) <= 256 and max(int(hd[4].shape[0]) for hd in heads) <= 4
        # columns per head: the convolve-first path only needs whole 16-byte vectors (4 x 200 instead of 4 x 208 columns for
        # the precise group: 4 % less Z / dZ traffic and GEMM work); the fused-epilogue GEMM needs a multiple of the UMMA N step
        slot = _ceil_to(max(inners), 8 if tapsplit else 16)
        ntot = slot * nh
        def fill_bias(buf: Tensor) -> None:
            for i, hd in enumerate(heads):
                _pack(hd[1].detach(), 0, 0, 1, 1, 1, inners[i], 0, None, buf, i * slot, 0, 0)
        # the heads' conv biases side by side in slot order: a cached, kernel-written staging vector (pad columns stay 0)
        bias = PACK.get(('head_bias', tuple(id(hd[1]) for hd in heads), slot), [hd[1] for hd in heads], (ntot,), torch.float32, fill_bias)
        train = _needs_grad(ctx)
        M = B * H * W
        outs = [torch.empty((B, int(hd[4].shape[0]), H, W), dtype=torch.float32, device=dev) for hd in heads]

        def head_tail() -> L.HeadTail:
            ht = L.HeadTail()
            ht.num_heads, ht.slot, ht.pixels_per_image = nh, slot, H * W
            for i, hd in enumerate(heads):
                ht.gamma[i], ht.beta[i] = hd[2].detach().data_ptr(), hd[3].detach().data_ptr()
                ht.w2[i], ht.b2[i] = hd[4].detach().data_ptr(), hd[5].detach().data_ptr()
                ht.out[i] = outs[i].data_ptr()
                ht.inner[i], ht.out_channels[i], ht.softplus[i] = inners[i], int(hd[4].shape[0]), int(softplus[i])
            return ht

        up = None
        if tapsplit:
            # convolve first, resample after (csrc/head_combine.cu): Z = X . W_tap^T on the LOW-resolution map -- factor^2
            # fewer FLOPs than the conv on the up-sampled map, which is never built -- then one pass that interpolates,
            # shifts and sums the taps and runs every head's LayerNorm/GELU/projection(/Softplus) on the fp32 sums
            T = ks * ks
            wz, c_pad, nz = packed_tapsplit_fwd([hd[0] for hd in heads], dt, slot)
            m_low = B * h * w
            z = torch.empty((m_low, nz), dtype=dt, device=dev)
            gemm_nt(x, 1, 1, m_low, C, ld, 1, wz, c_pad, nz, _epilogue(z, nz), alg_kn=C * T * sum(inners))
            conv = alloc_nhwc(B, H, W, ntot, dt, dev) if train else None
            ht = head_tail()
            if L.PROFILE.active:
                # algorithmic bytes: the heads' real channels of Z read once, the conv output written once, the fp32 maps
                L.PROFILE.note(f'head_combine_fwd {B}x{h}x{w} x{factor} ks{ks} N{ntot}', 0.0,
                               (m_low * T + (M if train else 0)) * sum(inners) * z.element_size()
                               + 4.0 * M * sum(int(hd[4].shape[0]) for hd in heads))
            L.check(L.LIB.vkocr_head_combine_fwd(_tag(dt), L.ptr(z), nz, B, h, w, factor, mode, ks, ntot, L.ptr(bias), ctypes.byref(ht),
                                                 L.ptr(conv), 0 if conv is None else conv.stride(3), COMBINE_ALGO, _s()),
                    'head_combine_fwd')
            del z
        else:
            if factor > 1:
                up = alloc_nhwc(B, H, W, C, dt, dev)
                upsample_fwd(x, up, C, mode, False)
            else:
                up = x
            wp, c_pad, rows = packed_conv_fwd([hd[0] for hd in heads], dt, slot)
            fused = dt == torch.bfloat16 and SIMT_BACKEND == 0 and nh <= L.MAX_HEADS and slot <= 256
            if fused:
                # conv + every head's LayerNorm/GELU/projection(/Softplus) in ONE tcgen05 GEMM: the tails run in the epilogue
                # on the fp32 accumulators; the bf16 conv output is only written when the backward will need it
                conv = alloc_nhwc(B, H, W, ntot, dt, dev) if train else None
                ht = head_tail()
                if L.PROFILE.active:
                    L.PROFILE.note(f'gemm_nt_heads ks{ks} M{M} K{ks * ks * C} N{ntot}', 2.0 * M * ks * ks * C * sum(inners))
                g = L.ConvGeom(B, H, W, ks, C, up.stride(3), c_pad)
                ep = L.Epilogue(None if conv is None else conv.data_ptr(), ntot if conv is None else conv.stride(3), 0, 0, None, 0,
                                bias.data_ptr(), 0, None, None, 1, None, 0, None, 0, 0, 0, 0)
                L.check(L.LIB.vkocr_gemm_nt_heads(_tag(dt), L.ptr(up), ctypes.byref(g), L.ptr(wp), ntot, ctypes.byref(ep),
                                                  ctypes.byref(ht), _s()), 'gemm_nt_heads')
            else:
                conv = alloc_nhwc(B, H, W, ntot, dt, dev)
                gemm_nt(up, B, H, W, C, up.stride(3), ks, wp, c_pad, ntot, _epilogue(conv, conv.stride(3), bias=bias))
                for i, hd in enumerate(heads):
                    O = int(hd[4].shape[0])
                    sl = conv[:, i * slot:(i + 1) * slot]
                    if L.PROFILE.active:
                        L.PROFILE.note(f'head_tail_fwd rows{M} inner{inners[i]} O{O}', 0.0, M * (slot * conv.element_size() + 4 * O))
                    L.check(L.LIB.vkocr_head_tail_fwd(_tag(dt), L.ptr(sl), conv.stride(3), inners[i], slot, L.ptr(hd[2].detach()),
                                                      L.ptr(hd[3].detach()), L.ptr(hd[4].detach()), L.ptr(hd[5].detach()), O,
                                                      int(softplus[i]), L.ptr(outs[i]), H * W, M, _s()), 'head_tail_fwd')
        if train:
            ctx.save_for_backward(x, x if up is None else up, conv, *outs, *params)
            ctx.meta = (nh, factor, mode, tuple(softplus), slot, ks, tapsplit)
        else:
            del conv
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts: Tensor):
        nh, factor, mode, softplus, slot, ks, tapsplit = ctx.meta
        saved = ctx.saved_tensors
        x, up, conv = saved[0], saved[1], saved[2]
        outs = saved[3:3 + nh]
        params = saved[3 + nh:]
        heads = [params[6 * i:6 * i + 6] for i in range(nh)]
        B, h, w, C, ld = geom(x)
        H, W = h * factor, w * factor
        M = B * H * W
        dt, dev = x.dtype, x.device
        ntot = slot * nh
        T = ks * ks
        dx = None
        # heads whose incoming gradient is zero outside the label points (registered by PreciseLossFn.backward) take the
        # label-point backward (csrc/head_sparse.cu): exact, and B * P rows of work instead of B * H * W
        points: Dict[int, Tuple[Tensor, Tensor]] = {}
        if tapsplit and SPARSE_BACKWARD:
            for i in range(nh):
                d = douts[i]
                info = SPARSE_GRADS.pop(d.data_ptr(), None) if d is not None else None
                if info is not None and d.dtype == torch.float32 and d.is_contiguous() and info[2] == (B, H, W) \
                        and info[0].shape[0] == B and 8 * info[0].numel() <= M:
                    points[i] = (info[0], info[1])
        dense_ids = [i for i in range(nh) if i not in points]

        def upstream(i: int) -> Tensor:
            dout = douts[i]
            if dout is None:
                dout = zero_(torch.empty_like(outs[i]))
            return dout.contiguous().float()

        n_d = slot * len(dense_ids) if tapsplit else ntot
        dconv = alloc_nhwc(B, H, W, n_d, dt, dev) if n_d > 0 else None
        for k, i in enumerate(dense_ids):
            hd = heads[i]
            inner, O = int(hd[0].shape[0]), int(hd[4].shape[0])
            dout = upstream(i)
            if L.PROFILE.active:
                L.PROFILE.note(f'head_tail_bwd rows{M} inner{inner} O{O}', 0.0, M * (2 * inner * conv.element_size() + 8 * O))
            L.check(L.LIB.vkocr_head_tail_bwd(_tag(dt), L.ptr(conv[:, i * slot:(i + 1) * slot]), conv.stride(3), inner, slot,
                                              L.ptr(hd[2].detach()), L.ptr(hd[3].detach()), L.ptr(hd[4].detach()), O,
                                              int(softplus[i]), L.ptr(outs[i]), L.ptr(dout), H * W, M,
                                              L.ptr(dconv[:, k * slot:(k + 1) * slot]), dconv.stride(3),
                                              L.ptr(grad_buffer(hd[2])), L.ptr(grad_buffer(hd[3])), L.ptr(grad_buffer(hd[4])),
                                              L.ptr(grad_buffer(hd[5])), L.ptr(grad_buffer(hd[1])), _s()), 'head_tail_bwd')
        if tapsplit:
            m_low = B * h * w
            if dense_ids:
                # adjoint of the interpolate + shift + sum, then two plain GEMMs on the low-resolution grid
                nz = T * n_d
                inner_sum = sum(int(heads[i][0].shape[0]) for i in dense_ids)
                dz = torch.empty((m_low, nz), dtype=dt, device=dev)
                if L.PROFILE.active:
                    L.PROFILE.note(f'head_combine_bwd {B}x{h}x{w} x{factor} ks{ks} N{n_d}', 0.0, (M + m_low * T) * inner_sum * dz.element_size())
                L.check(L.LIB.vkocr_head_combine_bwd(_tag(dt), L.ptr(dconv), dconv.stride(3), B, h, w, factor, mode, ks, n_d, L.ptr(dz), nz,
                                                     COMBINE_ALGO, _s()), 'head_combine_bwd')
                del dconv
                gw = _zeros_f32(nz * C, dev)
                gemm_tn(dz, 1, 1, m_low, nz, nz, 1, x, C, ld, _epilogue(gw, C, out_f32=True, accumulate=True, tn=(0, C, 1)),
                        alg_ij=C * T * inner_sum)
                for k, i in enumerate(dense_ids):
                    hd = heads[i]
                    # gw[(tap * n_d + k * slot + n) * C + c]  ->  grad[(n * C + c) * T + tap]
                    L.check(L.LIB.vkocr_scatter_add_f32(ctypes.c_void_p(gw.data_ptr() + 4 * k * slot * C), C, n_d * C, 1, int(hd[0].shape[0]),
                                                        T, C, L.ptr(grad_buffer(hd[0])), C * T, 1, T, _s()), 'scatter_add_f32')
                if ctx.needs_input_grad[0]:
                    wd, k_pad = packed_tapsplit_dgrad([heads[i][0] for i in dense_ids], dt, slot)
                    dx = alloc_nhwc(B, h, w, C, dt, dev)
                    gemm_nt(dz, 1, 1, m_low, nz, nz, 1, wd, k_pad, C, _epilogue(dx, dx.stride(3)), alg_kn=C * T * inner_sum)
                del dz
            elif ctx.needs_input_grad[0]:
                dx = alloc_nhwc(B, h, w, C, dt, dev)
                zero_(dx.permute(0, 2, 3, 1))
            # ---- label-point heads, grouped by point list
            groups: Dict[Tuple[int, int], List[int]] = {}
            for i, (py, px) in points.items():
                groups.setdefault((py.data_ptr(), px.data_ptr()), []).append(i)
            for ids in groups.values():
                py, px = points[ids[0]]
                P = int(py.shape[1])
                E = B * P
                n_s = slot * len(ids)
                owner = _zeros(B * H * W, torch.int32, dev)
                pix_index = torch.empty(E, dtype=torch.int32, device=dev)
                L.check(L.LIB.vkocr_points_claim(L.ptr(py), L.ptr(px), B, P, H, W, L.ptr(owner), L.ptr(pix_index), _s()), 'points_claim')
                g = torch.empty((E, n_s), dtype=dt, device=dev)
                for k, i in enumerate(ids):
                    hd = heads[i]
                    inner, O = int(hd[0].shape[0]), int(hd[4].shape[0])
                    dout = upstream(i)
                    if L.PROFILE.active:
                        L.PROFILE.note(f'head_tail_bwd_points entries{E} inner{inner} O{O}', 0.0, E * (2 * inner * conv.element_size() + 8 * O))
                    L.check(L.LIB.vkocr_head_tail_bwd_points(
                        _tag(dt), L.ptr(conv[:, i * slot:(i + 1) * slot]), conv.stride(3), inner, slot, L.ptr(hd[2].detach()),
                        L.ptr(hd[3].detach()), L.ptr(hd[4].detach()), O, int(softplus[i]), L.ptr(outs[i]), L.ptr(dout), H * W, L.ptr(pix_index), E, 0,
                        ctypes.c_void_p(g.data_ptr() + k * slot * g.element_size()), n_s, L.ptr(grad_buffer(hd[2])), L.ptr(grad_buffer(hd[3])),
                        L.ptr(grad_buffer(hd[4])), L.ptr(grad_buffer(hd[5])), L.ptr(grad_buffer(hd[1])), _s()), 'head_tail_bwd_points')
                # weight gradient: dW_tap = G^T . A_tap with A_tap the up-sampled, tap-shifted input rows of the label pixels
                a_pts = torch.empty((E, T * C), dtype=dt, device=dev)
                L.check(L.LIB.vkocr_gather_up_taps(_tag(dt), L.ptr(x), ld, B, h, w, C, factor, mode, ks, L.ptr(pix_index), E, L.ptr(a_pts), _s()),
                        'gather_up_taps')
                gw = _zeros_f32(n_s * T * C, dev)
                gemm_tn(g, 1, 1, E, n_s, n_s, 1, a_pts, T * C, T * C, _epilogue(gw, T * C, out_f32=True, accumulate=True, tn=(0, T * C, 1)),
                        alg_ij=T * C * sum(int(heads[i][0].shape[0]) for i in ids))
                for k, i in enumerate(ids):
                    hd = heads[i]
                    # gw[(k * slot + n) * T * C + tap * C + c]  ->  grad[(n * C + c) * T + tap]
                    L.check(L.LIB.vkocr_scatter_add_f32(ctypes.c_void_p(gw.data_ptr() + 4 * k * slot * T * C), T * C, C, 1, int(hd[0].shape[0]),
                                                        T, C, L.ptr(grad_buffer(hd[0])), C * T, 1, T, _s()), 'scatter_add_f32')
                if ctx.needs_input_grad[0]:
                    # data gradient: U = G . W per tap, then the adjoint of up-sample + shift scattered into dX
                    wp, k_pad = packed_tapsplit_points_dgrad([heads[i][0] for i in ids], dt, slot)
                    u = torch.empty((E, T * C), dtype=dt, device=dev)
                    gemm_nt(g, 1, 1, E, n_s, n_s, 1, wp, k_pad, T * C, _epilogue(u, T * C),
                            alg_kn=T * C * sum(int(heads[i][0].shape[0]) for i in ids))
                    L.check(L.LIB.vkocr_scatter_up_taps(_tag(dt), L.ptr(u), B, h, w, C, factor, mode, ks, L.ptr(pix_index), E, L.ptr(dx),
                                                        dx.stride(3), _s()), 'scatter_up_taps')
        else:
            # weight gradient of all heads in one pass over (dconv, up): G[n_total, C, T] in OIHW order, then per-head slices
            gw = _zeros_f32(ntot * C * T, dev)
            gemm_tn(dconv, B, H, W, ntot, dconv.stride(3), ks, up, C, up.stride(3),
                    _epilogue(gw, C, out_f32=True, accumulate=True, tn=(1, C * T, T)))
            for i, hd in enumerate(heads):
                inner = int(hd[0].shape[0])
                n = inner * C * T
                L.check(L.LIB.vkocr_accumulate_f32(ctypes.c_void_p(gw.data_ptr() + 4 * i * slot * C * T), L.ptr(grad_buffer(hd[0])), n, _s()),
                        'accumulate_f32')
            if ctx.needs_input_grad[0]:
                wd, n_pad = packed_conv_dgrad([hd[0] for hd in heads], dt, slot)
                dup = alloc_nhwc(B, H, W, C, dt, dev)
                gemm_nt(dconv, B, H, W, ntot, dconv.stride(3), ks, wd, n_pad, C, _epilogue(dup, dup.stride(3)))
                if factor > 1:
                    dx = alloc_nhwc(B, h, w, C, dt, dev)
                    upsample_bwd(dup, dx, C, mode, False)
                else:
                    dx = dup
        _ready(*params)
        return (dx, None, None, None) + (None,) * len(params)


class HeadGroupPointsFn(_Function):
    """Opt-in label-point evaluation of heads whose outputs the loss reads at the (B, P) label points only (the precise
    corner-offset / angle / distance heads, loss_function/adaptive_scaling.py:235-260): the conv outputs of those pixels are
    ONE small GEMM, conv[e, :] = bias + A[e, :] . W^T with A[e, tap, :] = up(x)[r_e + dy - k/2, s_e + dx - k/2, :]
    (csrc/head_sparse.cu), the LayerNorm -> GELU -> projection (-> Softplus) tail runs on those rows, and the returned NCHW
    maps hold the heads' values at the label pixels and ZERO elsewhere.  Loss and gradients equal the dense evaluation's
    (UperNextHead.forward upernext.py:233-248 / FpnHead.forward fpn.py:193-208); the maps do not, so this is never the default."""

    @staticmethod
    def supported(x: Tensor, factor: int, heads) -> bool:
        C = int(x.shape[1])
        inners = [int(hd[0].shape[0]) for hd in heads]
        return factor > 1 and C % 64 == 0 and len(heads) <= L.MAX_HEADS and max(inners) <= 256 \
            and max(int(hd[4].shape[0]) for hd in heads) <= 4

    @staticmethod
    def forward(ctx, x: Tensor, factor: int, mode: int, softplus: Tuple[bool, ...], py: Tensor, px: Tensor, *params: Tensor):
        nh = len(softplus)
        heads = [params[6 * i:6 * i + 6] for i in range(nh)]
        B, h, w, C, ld = geom(x)
        dt, dev = x.dtype, x.device
        H, W = h * factor, w * factor
        py, px = _i64c(py), _i64c(px)
        if tuple(py.shape) != tuple(px.shape) or py.dim() != 2 or py.shape[0] != B:
            raise L.VkocrError(f'label points {tuple(py.shape)} / {tuple(px.shape)} do not match the batch {B}')
        P = int(py.shape[1])
        E = B * P
        inners = [int(hd[0].shape[0]) for hd in heads]
        ks = int(heads[0][0].shape[2])
        T = ks * ks
        slot = _ceil_to(max(inners), 8)
        n_s = slot * nh
        owner = _zeros(B * H * W, torch.int32, dev)
        pix_index = torch.empty(E, dtype=torch.int32, device=dev)
        L.check(L.LIB.vkocr_points_claim(L.ptr(py), L.ptr(px), B, P, H, W, L.ptr(owner), L.ptr(pix_index), _s()), 'points_claim')
        a_pts = torch.empty((E, T * C), dtype=dt, device=dev)
        L.check(L.LIB.vkocr_gather_up_taps(_tag(dt), L.ptr(x), ld, B, h, w, C, factor, mode, ks, L.ptr(pix_index), E, L.ptr(a_pts), _s()),
                'gather_up_taps')
        wp, c_pad, _ = packed_conv_fwd([hd[0] for hd in heads], dt, slot)     # [n_s, T * c_pad], k = tap * c_pad + c (c_pad == C)

        def fill_bias(buf: Tensor) -> None:
            for i, hd in enumerate(heads):
                _pack(hd[1].detach(), 0, 0, 1, 1, 1, inners[i], 0, None, buf, i * slot, 0, 0)
        bias = PACK.get(('head_bias', tuple(id(hd[1]) for hd in heads), slot), [hd[1] for hd in heads], (n_s,), torch.float32, fill_bias)
        conv_pts = torch.empty((E, n_s), dtype=dt, device=dev)
        gemm_nt(a_pts, 1, 1, E, T * C, T * C, 1, wp, T * c_pad, n_s, _epilogue(conv_pts, n_s, bias=bias), alg_kn=T * C * sum(inners))
        outs = []
        for i, hd in enumerate(heads):
            O = int(hd[4].shape[0])
            out = zero_(torch.empty((B, O, H, W), dtype=torch.float32, device=dev))
            L.check(L.LIB.vkocr_head_tail_fwd_points(
                _tag(dt), ctypes.c_void_p(conv_pts.data_ptr() + i * slot * conv_pts.element_size()), n_s, inners[i], slot,
                L.ptr(hd[2].detach()), L.ptr(hd[3].detach()), L.ptr(hd[4].detach()), L.ptr(hd[5].detach()), O, int(softplus[i]), L.ptr(out),
                H * W, L.ptr(pix_index), E, _s()), 'head_tail_fwd_points')
            outs.append(out)
        if _needs_grad(ctx):
            ctx.save_for_backward(x, a_pts, conv_pts, pix_index, *outs, *params)
            ctx.meta = (nh, factor, mode, tuple(softplus), slot, ks)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts: Tensor):
        nh, factor, mode, softplus, slot, ks = ctx.meta
        saved = ctx.saved_tensors
        x, a_pts, conv_pts, pix_index = saved[:4]
        outs = saved[4:4 + nh]
        params = saved[4 + nh:]
        heads = [params[6 * i:6 * i + 6] for i in range(nh)]
        B, h, w, C, ld = geom(x)
        H, W = h * factor, w * factor
        dt, dev = x.dtype, x.device
        T = ks * ks
        E = int(pix_index.shape[0])
        n_s = slot * nh
        g = torch.empty((E, n_s), dtype=dt, device=dev)
        for i, hd in enumerate(heads):
            SPARSE_GRADS.pop(douts[i].data_ptr() if douts[i] is not None else 0, None)
            dout = douts[i] if douts[i] is not None else zero_(torch.empty_like(outs[i]))
            dout = dout.contiguous().float()
            inner, O = int(hd[0].shape[0]), int(hd[4].shape[0])
            L.check(L.LIB.vkocr_head_tail_bwd_points(
                _tag(dt), ctypes.c_void_p(conv_pts.data_ptr() + i * slot * conv_pts.element_size()), n_s, inner, slot, L.ptr(hd[2].detach()),
                L.ptr(hd[3].detach()), L.ptr(hd[4].detach()), O, int(softplus[i]), L.ptr(outs[i]), L.ptr(dout), H * W, L.ptr(pix_index), E, 1,
                ctypes.c_void_p(g.data_ptr() + i * slot * g.element_size()), n_s, L.ptr(grad_buffer(hd[2])), L.ptr(grad_buffer(hd[3])),
                L.ptr(grad_buffer(hd[4])), L.ptr(grad_buffer(hd[5])), L.ptr(grad_buffer(hd[1])), _s()), 'head_tail_bwd_points')
        inner_sum = sum(int(hd[0].shape[0]) for hd in heads)
        gw = _zeros_f32(n_s * T * C, dev)
        gemm_tn(g, 1, 1, E, n_s, n_s, 1, a_pts, T * C, T * C, _epilogue(gw, T * C, out_f32=True, accumulate=True, tn=(0, T * C, 1)),
                alg_ij=T * C * inner_sum)
        for i, hd in enumerate(heads):
            L.check(L.LIB.vkocr_scatter_add_f32(ctypes.c_void_p(gw.data_ptr() + 4 * i * slot * T * C), T * C, C, 1, int(hd[0].shape[0]), T, C,
                                                L.ptr(grad_buffer(hd[0])), C * T, 1, T, _s()), 'scatter_add_f32')
        dx = None
        if ctx.needs_input_grad[0]:
            wp, k_pad = packed_tapsplit_points_dgrad([hd[0] for hd in heads], dt, slot)
            u = torch.empty((E, T * C), dtype=dt, device=dev)
            gemm_nt(g, 1, 1, E, n_s, n_s, 1, wp, k_pad, T * C, _epilogue(u, T * C), alg_kn=T * C * inner_sum)
            dx = alloc_nhwc(B, h, w, C, dt, dev)
            zero_(dx.permute(0, 2, 3, 1))
            L.check(L.LIB.vkocr_scatter_up_taps(_tag(dt), L.ptr(u), B, h, w, C, factor, mode, ks, L.ptr(pix_index), E, L.ptr(dx), dx.stride(3),
                                                _s()), 'scatter_up_taps')
        _ready(*params)
        return (dx, None, None, None, None, None) + (None,) * len(params)


# ----------------------------------------------------------------------------------------------------- fused losses
def _f32c(t: Tensor) -> Tensor:
    """Caller-provided loss operand -> contiguous fp32 on its own device (no-op for the usual case)."""
    L.require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i64c(t: Tensor) -> Tensor:
    L.require_cuda(t)
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


class RoughLossFn(_Function):
    """focal + dice + masked log-space smooth-L1 of the rough maps inside the core box, one reduction pass
    (AdaptiveScalingRoughLossFunction.__call__, loss_function/adaptive_scaling.py:53-131)."""

    @staticmethod
    def forward(ctx, logit: Tensor, height: Tensor, gt_mask: Tensor, gt_score: Tensor, up: int, left: int,
                height_min: float, score_min: float, focal_factor: float, dice_factor: float, l1_factor: float) -> Tensor:
        logit_c, height_c = _f32c(logit), _f32c(height)
        gt_mask, gt_score = _f32c(gt_mask), _f32c(gt_score)
        B, _, H, W = logit_c.shape
        _, CH, CW = gt_mask.shape
        dev = logit_c.device
        sums = _zeros(6, torch.float64, dev)
        coef = torch.empty(6, dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_rough_loss_fwd(L.ptr(logit_c), L.ptr(height_c), L.ptr(gt_mask), L.ptr(gt_score), B, H, W, up, left, CH, CW,
                                           height_min, score_min, focal_factor, dice_factor, l1_factor, L.ptr(sums), L.ptr(coef), _s()),
                'rough_loss_fwd')
        ctx.save_for_backward(logit_c, height_c, gt_mask, gt_score)
        ctx.coef = coef
        ctx.meta = (B, H, W, up, left, CH, CW, height_min, score_min)
        return coef[0].clone()

    @staticmethod
    def backward(ctx, gout: Tensor):
        logit, height, gt_mask, gt_score = ctx.saved_tensors
        B, H, W, up, left, CH, CW, hmin, smin = ctx.meta
        gout = gout.contiguous().float()
        dlogit = torch.empty_like(logit)
        dheight = torch.empty_like(height)
        L.check(L.LIB.vkocr_rough_loss_bwd(L.ptr(logit), L.ptr(height), L.ptr(gt_mask), L.ptr(gt_score), B, H, W, up, left, CH, CW,
                                           hmin, smin, L.ptr(ctx.coef), L.ptr(gout), L.ptr(dlogit), L.ptr(dheight), _s()),
                'rough_loss_bwd')
        return (dlogit, dheight) + (None,) * 9


_FACTOR_CACHE: Dict[Tuple, Tensor] = {}


def _device_factors(values: Tuple[float, ...], device) -> Tensor:
    key = (values, str(device))
    t = _FACTOR_CACHE.get(key)
    if t is None:
        t = torch.tensor(values, dtype=torch.float32, device=device)
        _FACTOR_CACHE[key] = t
    return t


class PreciseLossFn(_Function):
    """Dense pos/neg L2 of sigmoid(prob) inside the core box + label-point gather terms (offset smooth-L1, distance
    regulation, soft-label CE of the corner angles, corner-distance smooth-L1), x loss_factor
    (AdaptiveScalingPreciseLossFunction.__call__, loss_function/adaptive_scaling.py:181-346)."""

    @staticmethod
    def forward(ctx, prob: Tensor, off: Tensor, ang: Tensor, dist: Tensor, gt_score: Tensor, gt_mask: Tensor, up: int, left: int,
                py: Tensor, px: Tensor, gt_off: Tensor, gt_ang: Tensor, gt_dist: Tensor, beta: float,
                factors: Tuple[float, ...]) -> Tensor:
        prob_c, off_c, ang_c, dist_c = _f32c(prob), _f32c(off), _f32c(ang), _f32c(dist)
        gt_score, gt_mask = _f32c(gt_score), _f32c(gt_mask)
        py, px = _i64c(py), _i64c(px)
        gt_off, gt_ang, gt_dist = _f32c(gt_off), _f32c(gt_ang), _f32c(gt_dist)   # int64 offsets (the dataset's) promote like F.smooth_l1_loss
        B, _, H, W = prob_c.shape
        P = int(py.shape[1]) if py.dim() == 2 else -1
        if not (py.shape == px.shape == (B, P) and tuple(gt_off.shape) == (B, P, 2) and tuple(gt_ang.shape) == (B, P, 4)
                and tuple(gt_dist.shape) == (B, P, 3) and off_c.shape[0] == ang_c.shape[0] == dist_c.shape[0] == gt_score.shape[0]
                == gt_mask.shape[0] == B):
            raise L.VkocrError(f'precise loss: label tensors do not match the prediction batch {B} / point count {P}: '
                               f'y {tuple(py.shape)} x {tuple(px.shape)} offsets {tuple(gt_off.shape)} angles {tuple(gt_ang.shape)} '
                               f'distances {tuple(gt_dist.shape)} score map {tuple(gt_score.shape)} mask {tuple(gt_mask.shape)}')
        B, _, H, W = prob_c.shape
        _, CH, CW = gt_mask.shape
        P = int(py.shape[1])
        dev = prob_c.device
        fac = _device_factors(tuple(float(f) for f in factors), dev)
        sums = _zeros(8, torch.float64, dev)
        coef = torch.empty(3, dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_precise_loss_fwd(L.ptr(prob_c), L.ptr(off_c), L.ptr(ang_c), L.ptr(dist_c), L.ptr(gt_score), L.ptr(gt_mask),
                                             B, H, W, up, left, CH, CW, L.ptr(py), L.ptr(px), L.ptr(gt_off), L.ptr(gt_ang),
                                             L.ptr(gt_dist), P, beta, L.ptr(fac), L.ptr(sums), L.ptr(coef), _s()), 'precise_loss_fwd')
        ctx.save_for_backward(prob_c, off_c, ang_c, dist_c, gt_score, gt_mask, py, px, gt_off, gt_ang, gt_dist)
        ctx.coef = coef
        ctx.fac = fac
        ctx.meta = (B, H, W, up, left, CH, CW, P, beta)
        return coef[0].clone()

    @staticmethod
    def backward(ctx, gout: Tensor):
        prob, off, ang, dist, gt_score, gt_mask, py, px, gt_off, gt_ang, gt_dist = ctx.saved_tensors
        B, H, W, up, left, CH, CW, P, beta = ctx.meta
        gout = gout.contiguous().float()
        dprob = torch.empty_like(prob)
        doff, dang, ddist = zero_(torch.empty_like(off)), zero_(torch.empty_like(ang)), zero_(torch.empty_like(dist))
        L.check(L.LIB.vkocr_precise_loss_bwd(L.ptr(prob), L.ptr(off), L.ptr(ang), L.ptr(dist), L.ptr(gt_score), L.ptr(gt_mask), B, H, W,
                                             up, left, CH, CW, L.ptr(py), L.ptr(px), L.ptr(gt_off), L.ptr(gt_ang), L.ptr(gt_dist), P,
                                             beta, L.ptr(ctx.fac), L.ptr(ctx.coef), L.ptr(gout), L.ptr(dprob), L.ptr(doff),
                                             L.ptr(dang), L.ptr(ddist), _s()), 'precise_loss_bwd')
        # the offset / angle / distance gradients are zero outside the label points (they were zero-filled and scattered into)
        SPARSE_GRADS.clear()
        for t in (doff, dang, ddist):
            SPARSE_GRADS[t.data_ptr()] = (py, px, (B, H, W))
        return (dprob, doff, dang, ddist) + (None,) * 11


# ----------------------------------------------------------------------------------------------------- primitive losses
FOCAL, DICE, L1, SMOOTH_L1, L2, WAHR = range(6)


class PointwiseLossFn(_Function):
    """One of the element-wise primitive losses with optional mask, reduced on the device
    (focal_with_logits.py, dice.py, l1.py, l2.py, weight_adaptive_heatmap_regression.py)."""

    @staticmethod
    def forward(ctx, pred: Tensor, gt: Tensor, mask: Optional[Tensor], kind: int, p0: float, p1: float,
                pre_sigmoid: bool = False) -> Tensor:
        pred_c, gt_c = _f32c(pred), _f32c(gt)
        if gt_c.shape != pred_c.shape:
            gt_c = gt_c.expand_as(pred_c).contiguous()
        mask_c = None
        if mask is not None:
            mask_c = _f32c(mask)
            if mask_c.shape != pred_c.shape:
                mask_c = mask_c.expand_as(pred_c).contiguous()
        n = pred_c.numel()
        dev = pred_c.device
        sums = torch.zeros(3, dtype=torch.float64, device=dev)
        coef = torch.empty(3, dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_pointwise_loss_fwd(kind, int(pre_sigmoid), L.ptr(pred_c), L.ptr(gt_c), L.ptr(mask_c), n, p0, p1, L.ptr(sums), L.ptr(coef),
                                               _s()), 'pointwise_loss_fwd')
        ctx.save_for_backward(pred_c, gt_c, mask_c if mask_c is not None else torch.empty(0, device=dev))
        ctx.coef = coef
        ctx.meta = (kind, p0, p1, mask_c is not None, tuple(pred.shape), int(pre_sigmoid))
        return coef[0].clone()

    @staticmethod
    def backward(ctx, gout: Tensor):
        pred, gt, mask = ctx.saved_tensors
        kind, p0, p1, has_mask, shape, pre_sigmoid = ctx.meta
        gout = gout.contiguous().float()
        dpred = torch.empty_like(pred)
        L.check(L.LIB.vkocr_pointwise_loss_bwd(kind, pre_sigmoid, L.ptr(pred), L.ptr(gt), L.ptr(mask if has_mask else None), pred.numel(), p0, p1,
                                               L.ptr(ctx.coef), L.ptr(gout), L.ptr(dpred), _s()), 'pointwise_loss_bwd')
        return dpred.reshape(shape), None, None, None, None, None, None


class SoftCrossEntropyFn(_Function):
    """F.cross_entropy(pred, gt) with probability targets, class axis 1 (cross_entropy_with_logits.py:16-19)."""

    @staticmethod
    def forward(ctx, pred: Tensor, gt: Tensor) -> Tensor:
        pred_c, gt_c = _f32c(pred), _f32c(gt)
        if pred_c.dim() < 2 or pred_c.shape != gt_c.shape:
            raise L.VkocrError('soft cross entropy expects pred and gt of equal shape (N, C, ...)')
        outer, C = int(pred_c.shape[0]), int(pred_c.shape[1])
        inner = pred_c.numel() // max(outer * C, 1)
        dev = pred_c.device
        sums = torch.zeros(1, dtype=torch.float64, device=dev)
        coef = torch.empty(2, dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_soft_ce_fwd(L.ptr(pred_c), L.ptr(gt_c), outer, C, inner, L.ptr(sums), L.ptr(coef), _s()), 'soft_ce_fwd')
        ctx.save_for_backward(pred_c, gt_c)
        ctx.coef = coef
        ctx.meta = (outer, C, inner)
        return coef[0].clone()

    @staticmethod
    def backward(ctx, gout: Tensor):
        pred, gt = ctx.saved_tensors
        outer, C, inner = ctx.meta
        gout = gout.contiguous().float()
        dpred = torch.empty_like(pred)
        L.check(L.LIB.vkocr_soft_ce_bwd(L.ptr(pred), L.ptr(gt), outer, C, inner, L.ptr(ctx.coef), L.ptr(gout), L.ptr(dpred), _s()),
                'soft_ce_bwd')
        return dpred, None


class HardNegativeBceFn(_Function):
    """BCE-with-logits over all positives plus the k hardest negatives, k = min(round(ratio * #pos), #neg), selected by
    a device-side radix select (weighted_bce_with_logits.py:18-54; no host synchronisation)."""

    @staticmethod
    def forward(ctx, pred: Tensor, gt: Tensor, mask: Optional[Tensor], negative_ratio: float, eps: float) -> Tensor:
        pred_c, gt_c = _f32c(pred), _f32c(gt)
        mask_c = _f32c(mask) if mask is not None else None
        n = pred_c.numel()
        dev = pred_c.device
        state = torch.zeros(8 + 256, dtype=torch.int64, device=dev)
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        coef = torch.empty(4, dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_hard_negative_bce_fwd(L.ptr(pred_c), L.ptr(gt_c), L.ptr(mask_c), n, negative_ratio, eps, L.ptr(state),
                                                  L.ptr(sums), L.ptr(coef), _s()), 'hard_negative_bce_fwd')
        ctx.save_for_backward(pred_c, gt_c, mask_c if mask_c is not None else torch.empty(0, device=dev))
        ctx.coef, ctx.state = coef, state
        ctx.meta = (mask_c is not None, tuple(pred.shape))
        return coef[0].clone()

    @staticmethod
    def backward(ctx, gout: Tensor):
        pred, gt, mask = ctx.saved_tensors
        has_mask, shape = ctx.meta
        gout = gout.contiguous().float()
        dpred = torch.empty_like(pred)
        ties = torch.zeros(1, dtype=torch.int64, device=pred.device)
        L.check(L.LIB.vkocr_hard_negative_bce_bwd(L.ptr(pred), L.ptr(gt), L.ptr(mask if has_mask else None), pred.numel(),
                                                  L.ptr(ctx.state), L.ptr(ctx.coef), L.ptr(gout), L.ptr(ties), L.ptr(dpred), _s()),
                'hard_negative_bce_bwd')
        return dpred.reshape(shape), None, None, None, None
