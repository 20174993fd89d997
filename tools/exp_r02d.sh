#!/bin/bash
# Round-2 late experiments on one B200 (run through gpurun from the repo root).
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_ops.py -q -m gpu -k "combine or convolve_first or head" > $O/r02d_tests.log 2>&1; tail -3 $O/r02d_tests.log
for lib in _bisect/lib_base.so ""; do
    echo "== precise group, lib '$lib'"; VKOCR_B200_LIB=$lib python tools/profile_combine.py 3 fwd 2>&1 | tail -4
done > $O/r02d_combine_ab.log 2>&1
cat $O/r02d_combine_ab.log
