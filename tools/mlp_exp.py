import os, sys
sys.path.insert(0, '/root/repo')
import torch
from vkit_ocr_model_adaptive_scaling_b200 import ops
dev = torch.device('cuda:0')
def t(fn, reps=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for (M, C) in ((819200, 96), (51200, 384)):
    hid = 4 * C
    cp = (C + 63) // 64 * 64
    x = torch.randn(M, C, device=dev).bfloat16(); w1 = torch.randn(hid, cp, device=dev).bfloat16()
    b1 = torch.randn(hid, device=dev); g = torch.empty(M, hid, device=dev, dtype=torch.bfloat16); hp = torch.empty_like(g)
    for env in ({}, {'VKOCR_DEBUG_SKIP_TMA': '1'}):
        for k, v in env.items(): os.environ[k] = v
        a = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, hid, ops._epilogue(g, hid)))
        b = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, hid, ops._epilogue(g, hid, bias=b1, act=1)))
        c = t(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, hid, ops._epilogue(g, hid, out_pre=hp, ld_pre=hid, bias=b1, act=1)))
        print(f'M{M} K{C} N{hid} {env}: plain {a:.3f}  gelu {b:.3f}  gelu+pre {c:.3f} ms', flush=True)
        for k in env: del os.environ[k]
