"""GPU check of the CTA-pair (cta_group::2) GEMM mode against the one-CTA mode: same inputs, outputs must agree to fp32
summation-order noise; then timing of both.  VKOCR_CTA2=0/1 selects the mode at call time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vkit_ocr_model_adaptive_scaling_b200 import ops
dev = torch.device('cuda:0')


def run(mode, B, H, W, C, N, ks, reps=0):
    os.environ['VKOCR_CTA2'] = str(mode)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev); x.copy_(torch.randn(B, C, H, W, device=dev, generator=g))
    w = (torch.randn(N, ks * ks * C, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev, generator=g)
    out = ops.alloc_nhwc(B, H, W, N, torch.bfloat16, dev)
    fn = lambda: ops.gemm_nt(x, B, H, W, C, x.stride(3), ks, w, C, N, ops._epilogue(out, out.stride(3), bias=bias))
    fn(); torch.cuda.synchronize()
    ms = None
    if reps:
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
    return out.float().clone(), ms


for (B, H, W, C, N, ks) in ((2, 32, 32, 384, 384, 3), (2, 40, 56, 384, 832, 3), (4, 16, 16, 1536, 384, 1), (3, 20, 28, 384, 192, 3)):
    a, _ = run(0, B, H, W, C, N, ks)
    b, _ = run(1, B, H, W, C, N, ks)
    err = float((a - b).abs().max()), float(a.abs().max())
    print(f'B{B} {H}x{W} C{C} N{N} ks{ks}: max |pair - single| = {err[0]:.3e} (max |out| {err[1]:.2f})', flush=True)
for (B, H, W, C, N, ks) in ((32, 320, 320, 384, 832, 3), (32, 320, 320, 384, 384, 3)):
    fl = 2.0 * B * H * W * ks * ks * C * N
    for mode in (0, 1, 0, 1):
        _, ms = run(mode, B, H, W, C, N, ks, reps=5)
        print(f'N{N} cta2={mode}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s', flush=True)
