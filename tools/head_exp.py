"""Timing experiment: the fused head-group GEMM (head mode) against the plain NT GEMM of the same shape, with parts of the
head epilogue switched off (VKOCR_DEBUG_SKIP_TMA bit 6 = no LayerNorm/GELU/projection phase, bit 7 = no conv-output store).
Garbage results under the debug bits: timing only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vkit_ocr_model_adaptive_scaling_b200 import ops, _lib as L
dev = torch.device('cuda:0')
B, H, W, C = 32, 320, 320, 384


def t(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev); x.normal_()
for (nh, slot, inners, outs_c) in ((2, 192, (192, 192), (1, 1)), (4, 208, (192, 193, 194, 194), (1, 2, 4, 4))):
    ntot = nh * slot
    w = (torch.randn(ntot, 9 * C, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.randn(ntot, device=dev)
    conv = ops.alloc_nhwc(B, H, W, ntot, torch.bfloat16, dev)
    fl = 2.0 * B * H * W * 9 * C * ntot
    ht = L.HeadTail()
    ht.num_heads, ht.slot, ht.pixels_per_image = nh, slot, H * W
    keep = []
    outs = []
    for i in range(nh):
        gm, bt = torch.rand(inners[i], device=dev) + 0.5, torch.randn(inners[i], device=dev)
        w2, b2 = torch.randn(outs_c[i], inners[i], device=dev) * 0.1, torch.randn(outs_c[i], device=dev)
        o = torch.empty(B, outs_c[i], H, W, device=dev)
        keep += [gm, bt, w2, b2, o]
        ht.gamma[i], ht.beta[i], ht.w2[i], ht.b2[i], ht.out[i] = gm.data_ptr(), bt.data_ptr(), w2.data_ptr(), b2.data_ptr(), o.data_ptr()
        ht.inner[i], ht.out_channels[i], ht.softplus[i] = inners[i], outs_c[i], 0
    g = L.ConvGeom(B, H, W, 3, C, x.stride(3), C)
    ep = L.Epilogue(conv.data_ptr(), conv.stride(3), 0, 0, None, 0, bias.data_ptr(), 0, None, None, 1, None, 0, None, 0, 0, 0, 0)
    epi = L.Epilogue(None, ntot, 0, 0, None, 0, bias.data_ptr(), 0, None, None, 1, None, 0, None, 0, 0, 0, 0)

    def heads(e):
        L.check(L.LIB.vkocr_gemm_nt_heads(1, L.ptr(x), ctypes.byref(g), L.ptr(w), ntot, ctypes.byref(e), ctypes.byref(ht), L.stream_ptr()), 'heads')
    ms = t(lambda: ops.gemm_nt(x, B, H, W, C, x.stride(3), 3, w, C, ntot, ops._epilogue(conv, conv.stride(3), bias=bias)))
    print(f'N{ntot}: plain NT                      {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s', flush=True)
    for dbg, what in ((0, 'head mode (training)'), (64, 'head mode, no phase 2'), (128, 'head mode, no conv store'), (192, 'head mode, neither')):
        os.environ['VKOCR_DEBUG_SKIP_TMA'] = str(dbg)
        ms = t(lambda: heads(ep))
        print(f'N{ntot}: {what:30s} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s', flush=True)
    os.environ['VKOCR_DEBUG_SKIP_TMA'] = '0'
    ms = t(lambda: heads(epi))
    print(f'N{ntot}: head mode (inference)          {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s', flush=True)
