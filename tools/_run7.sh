mkdir -p gpurun_out
python tools/kbench.py ln > gpurun_out/x7_ln_occ.log 2>&1; cat gpurun_out/x7_ln_occ.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "layer or neck" 2>&1 | tail -2
