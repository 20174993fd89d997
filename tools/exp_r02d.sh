#!/bin/bash
O=gpurun_out
mkdir -p $O
for v in 0 1; do
  echo "== pmajor $v"; VKOCR_HCB_PMAJOR=$v python tools/profile_combine.py 3 bwd 2>&1 | tail -2; ROUGH=1 VKOCR_HCB_PMAJOR=$v python tools/profile_combine.py 3 bwd 2>&1 | tail -2
done > $O/r02d_hcb_ab.log 2>&1
cat $O/r02d_hcb_ab.log
VKOCR_HCB_PMAJOR=1 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "combine or convolve_first" 2>&1 | tail -2
