"""GPU: one forward and one weight-gradient launch of the depthwise 7x7 kernels at the stage-0 shape (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import ops  # noqa: E402
dev = torch.device('cuda:0')
B, H, W, C = 32, 160, 160, 96
x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev); x.normal_()
y = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
dy = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev); dy.normal_()
wt = torch.randn(49, C, device=dev); bias = torch.randn(C, device=dev); dw = torch.zeros(C, 1, 7, 7, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.dwconv7(x, y, wt, bias, None)
    ops.dwconv7_wgrad(dy, x, dw)
torch.cuda.synchronize()
print('ok')
