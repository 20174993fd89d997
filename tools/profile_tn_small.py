"""GPU: the stage-2 / stage-3 ConvNeXt weight-gradient GEMMs (TN) at B=32/640x640, a few launches each (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vkit_ocr_model_adaptive_scaling_b200 import ops
dev = torch.device('cuda:0')
for (M, I, J) in ((51200, 1536, 392), (51200, 384, 1536), (12800, 3072, 776)):
    p = torch.randn(M, I, device=dev).to(torch.bfloat16)
    q = torch.randn(M, J, device=dev).to(torch.bfloat16)
    out = torch.zeros(I * J, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for r in range(3):
        ev[0].record()
        ops.gemm_tn(p, 1, 1, M, I, I, 1, q, J, J, ops._epilogue(out, J, out_f32=True, accumulate=True, tn=(0, J, 1)))
        ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    print(f'TN M{M} I{I} J{J}: {ms:.3f} ms {2.0 * M * I * J / ms / 1e9:.0f} TF/s', flush=True)
