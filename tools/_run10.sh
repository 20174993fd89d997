mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/x10_tests.log 2>&1; tail -3 gpurun_out/x10_tests.log
python tools/kbench.py mlp > gpurun_out/x10_mlp_prefetch.log 2>&1
VKOCR_NO_ACC_PREFETCH=1 python tools/kbench.py mlp > gpurun_out/x10_mlp_noprefetch.log 2>&1
paste -d'|' <(cut -c1-72 gpurun_out/x10_mlp_prefetch.log) <(cut -c60-72 gpurun_out/x10_mlp_noprefetch.log)
python tools/profile_combine.py 2 z dgrad wgrad | tail -3
VKOCR_NO_ACC_PREFETCH=1 python tools/profile_combine.py 2 z dgrad wgrad | tail -3
