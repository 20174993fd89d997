mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "head" > gpurun_out/x8_tests.log 2>&1; tail -5 gpurun_out/x8_tests.log
ROUGH=1 python tools/profile_combine.py 3 fwd > gpurun_out/x8_rough_half.log 2>&1
VKOCR_HC_NOHALF=1 ROUGH=1 python tools/profile_combine.py 3 fwd > gpurun_out/x8_rough_nohalf.log 2>&1
python tools/profile_combine.py 3 fwd > gpurun_out/x8_precise_half.log 2>&1
echo half; tail -2 gpurun_out/x8_rough_half.log; echo nohalf; tail -2 gpurun_out/x8_rough_nohalf.log; echo precise; tail -2 gpurun_out/x8_precise_half.log
