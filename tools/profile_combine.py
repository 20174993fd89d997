"""GPU: the 'convolve first, resample after' head path of the precise head group at the full B=32 / 640x640 shapes, one
C-ABI call per line (for CUDA-event timing and as the target of ncu captures):
  Z GEMM        [819200, 384] x [384, 9*800]              (vkocr_gemm_nt)
  combine fwd   Z -> 4 heads' conv outputs + prediction maps   (vkocr_head_combine_fwd)
  combine bwd   d(conv) [3276800, 800] -> dZ                   (vkocr_head_combine_bwd)
  dgrad / wgrad GEMMs on (dZ, x)
python tools/profile_combine.py [reps] [what ...]   what: z fwd bwd dgrad wgrad (default all)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import ops  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import _lib as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
what = set(sys.argv[2:]) or {'z', 'fwd', 'bwd', 'dgrad', 'wgrad'}
dev = torch.device('cuda:0')
BF = torch.bfloat16
B, h, w, C = 32, 160, 160, 384
inners, outs_c, soft = (192, 193, 194, 194), (1, 2, 4, 4), (0, 0, 0, 1)
if os.environ.get('ROUGH'):
    inners, outs_c, soft = (192, 192), (1, 1), (0, 1)
nh = len(inners)
slot = (max(inners) + 7) // 8 * 8       # the product's slot width on the convolve-first path (ops.HeadGroupFn)
ntot, T = slot * nh, 9
nz = T * ntot
H, W = 2 * h, 2 * w
m_low, M = B * h * w, B * H * W
x = ops.alloc_nhwc(B, h, w, C, BF, dev)
x.normal_()
wz = (torch.randn(nz, C, device=dev) * 0.05).to(BF)
z = torch.empty(m_low, nz, dtype=BF, device=dev)
conv = ops.alloc_nhwc(B, H, W, ntot, BF, dev)
dconv = torch.randn(M, ntot, device=dev).to(BF)
dz = torch.empty(m_low, nz, dtype=BF, device=dev)
bias = torch.randn(ntot, device=dev)
par = [(torch.rand(i, device=dev) + 0.5, torch.randn(i, device=dev) * 0.1, torch.randn(o, i, device=dev) * 0.1, torch.randn(o, device=dev))
       for i, o in zip(inners, outs_c)]
outs = [torch.empty(B, o, H, W, device=dev) for o in outs_c]
ht = L.HeadTail()
ht.num_heads, ht.slot, ht.pixels_per_image = nh, slot, H * W
for i in range(nh):
    ht.gamma[i], ht.beta[i], ht.w2[i], ht.b2[i] = (t.data_ptr() for t in par[i])
    ht.out[i] = outs[i].data_ptr()
    ht.inner[i], ht.out_channels[i], ht.softplus[i] = inners[i], outs_c[i], soft[i]
k_pad = (nz + 63) // 64 * 64
wd = (torch.randn(C, k_pad, device=dev) * 0.05).to(BF)
gw = torch.zeros(nz * C, device=dev)
dx = ops.alloc_nhwc(B, h, w, C, BF, dev)
s = ops._s
algo = int(os.environ.get('ALGO', '0'))


def timed(name, fn, flops=0.0, nbytes=0.0):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{name:12s} {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TF/s  {nbytes / ms / 1e6:7.0f} GB/s', flush=True)


for r in range(reps):
    if 'z' in what:
        timed('z gemm', lambda: ops.gemm_nt(x, 1, 1, m_low, C, x.stride(3), 1, wz, C, nz, ops._epilogue(z, nz)), 2.0 * m_low * C * nz,
              2.0 * m_low * (C + nz))
    if 'fwd' in what:
        timed('combine fwd', lambda: L.check(L.LIB.vkocr_head_combine_fwd(1, L.ptr(z), nz, B, h, w, 2, 0, 3, ntot, L.ptr(bias), ctypes.byref(ht),
                                                                          L.ptr(conv), conv.stride(3), algo, s())), 0.0, 2.0 * (m_low * nz + M * ntot))
        timed('combine inf', lambda: L.check(L.LIB.vkocr_head_combine_fwd(1, L.ptr(z), nz, B, h, w, 2, 0, 3, ntot, L.ptr(bias), ctypes.byref(ht),
                                                                          None, 0, algo, s())), 0.0, 2.0 * m_low * nz)
    if 'bwd' in what:
        timed('combine bwd', lambda: L.check(L.LIB.vkocr_head_combine_bwd(1, L.ptr(dconv), ntot, B, h, w, 2, 0, 3, ntot, L.ptr(dz), nz, algo, s())),
              0.0, 2.0 * (m_low * nz + M * ntot))
    if 'dgrad' in what:
        timed('dgrad gemm', lambda: ops.gemm_nt(dz, 1, 1, m_low, nz, nz, 1, wd, k_pad, C, ops._epilogue(dx, dx.stride(3))), 2.0 * m_low * C * nz,
              2.0 * m_low * (C + nz))
    if 'wgrad' in what:
        timed('wgrad gemm', lambda: ops.gemm_tn(dz, 1, 1, m_low, nz, nz, 1, x, C, x.stride(3),
                                                ops._epilogue(gw, C, out_f32=True, accumulate=True, tn=(0, C, 1))), 2.0 * m_low * C * nz,
              2.0 * m_low * (C + nz))
