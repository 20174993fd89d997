"""GPU: launches the dominant kernels of the training step at their full B=32 / 640x640 shapes a few times (for ncu):
  NT  implicit-GEMM 3x3 conv of the precise head group  M=3 276 800, K=9*384, N=832  (forward; dgrad is the mirror)
  TN  weight gradient of the same conv                     pixels=3 276 800, I=832, J=384, 9 taps
python tools/profile_head_conv.py [reps]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import ops  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import _lib as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda:0')
B, H, W, C, N = 32, 320, 320, 384, 832
x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
x.normal_()
w = torch.randn(N, 9 * 384, device=dev).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
conv = ops.alloc_nhwc(B, H, W, N, torch.bfloat16, dev)
gw = torch.zeros(N * C * 9, dtype=torch.float32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    ops.gemm_nt(x, B, H, W, C, x.stride(3), 3, w, 384, N, ops._epilogue(conv, conv.stride(3), bias=bias))
    ev[1].record()
    ev[2].record()
    ops.gemm_tn(conv, B, H, W, N, conv.stride(3), 3, x, C, x.stride(3),
                ops._epilogue(gw, C, out_f32=True, accumulate=True, tn=(1, C * 9, 9)))
    ev[3].record()
    torch.cuda.synchronize()
    fl = 2.0 * B * H * W * 9 * C * N
    print(f'rep {r}: NT {ev[0].elapsed_time(ev[1]):.3f} ms {fl / ev[0].elapsed_time(ev[1]) / 1e9:.1f} TF/s | '
          f'TN {ev[2].elapsed_time(ev[3]):.3f} ms {fl / ev[2].elapsed_time(ev[3]) / 1e9:.1f} TF/s', flush=True)
