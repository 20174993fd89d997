"""Timing experiments on the MLP-shaped tcgen05 GEMMs: which part of the pipeline bounds a small-K / wide-N tile.
VKOCR_DEBUG_SKIP_TMA bits: 1 = producers arrive without loading, 8 = epilogue without global stores, 16 = epilogue is
barrier traffic only (garbage results: timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vkit_ocr_model_adaptive_scaling_b200 import ops
dev = torch.device('cuda:0')


def t(fn, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


cases = [(819200, 96, 384), (819200, 96, 96), (819200, 384, 96), (51200, 384, 1536), (204800, 192, 768)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (M, C, N) in cases:
    cp = (C + 63) // 64 * 64
    x = torch.randn(M, C, device=dev).bfloat16(); w1 = torch.randn(N, cp, device=dev).bfloat16()
    b1 = torch.randn(N, device=dev); g = torch.empty(M, N, device=dev, dtype=torch.bfloat16); hp = torch.empty_like(g)
    for env in (0, 17, 49):
        os.environ['VKOCR_DEBUG_SKIP_TMA'] = str(env)
        a = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, N, ops._epilogue(g, N)))
        b = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, N, ops._epilogue(g, N, bias=b1, act=1)))
        c = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, N, ops._epilogue(g, N, out_pre=hp, ld_pre=N, bias=b1, act=1)))
        print(f'M{M} K{C} N{N} dbg={env:2d}: plain {a:.3f}  gelu {b:.3f}  gelu+pre {c:.3f} ms', flush=True)
    del os.environ['VKOCR_DEBUG_SKIP_TMA']
