"""GPU: the memory-bound ConvNeXt MLP GEMMs of stage 0 at B=32/640x640 (for ncu):
  MLP-1  [819200 x 96] x [96 x 384] + bias, GELU, keeps the pre-activation  (reads 157 MB, writes 2 x 629 MB)
  MLP-2  [819200 x 384] x [384 x 96] + bias, layer scale, residual           (reads 629 + 157 MB, writes 157 MB)
python tools/profile_mlp1.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda:0')
M, C = 819200, 96
x = torch.randn(M, C, device=dev).to(torch.bfloat16)
w1 = torch.randn(4 * C, 128, device=dev).to(torch.bfloat16)
w2 = torch.randn(C, 4 * C, device=dev).to(torch.bfloat16)
b1 = torch.randn(4 * C, device=dev)
b2 = torch.randn(C, device=dev)
gam = torch.rand(C, device=dev)
g = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
hpre = torch.empty_like(g)
y = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for r in range(reps):
    ev[0].record()
    ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, 128, 4 * C, ops._epilogue(g, 4 * C, out_pre=hpre, ld_pre=4 * C, bias=b1, act=1))
    ev[1].record()
    ops.gemm_nt(g, 1, 1, M, 4 * C, 4 * C, 1, w2, 4 * C, C, ops._epilogue(y, C, bias=b2, col_scale=gam, residual=x, ld_res=C))
    ev[2].record()
    ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, 128, 4 * C, ops._epilogue(g, 4 * C))
    ev[3].record()
    torch.cuda.synchronize()
    t1, t2, t3 = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    print(f'rep {r}: MLP-1 {t1:.3f} ms ({(M * C * 2 + 2 * M * 4 * C * 2) / t1 / 1e6:.0f} GB/s) | '
          f'MLP-2 {t2:.3f} ms ({(M * 4 * C * 2 + 2 * M * C * 2) / t2 / 1e6:.0f} GB/s) | plain {t3:.3f} ms ({(M * C * 2 + M * 4 * C * 2) / t3 / 1e6:.0f} GB/s)', flush=True)
