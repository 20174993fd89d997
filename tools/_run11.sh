mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/x11_tests.log 2>&1; tail -3 gpurun_out/x11_tests.log
python tools/kbench.py mlp ln 2>&1 | grep -E "eval|act1" > gpurun_out/x11_kbench.log; cat gpurun_out/x11_kbench.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-eager-baseline --profile > gpurun_out/x11_bench.json 2> gpurun_out/x11_bench.err; head -c 330 gpurun_out/x11_bench.json; echo
cp gpurun_out/kernel_table_train_upernext.json gpurun_out/x11_kernel_table_train_upernext.json
python bench.py --workload infer --steps 10 --warmup 3 --profile > gpurun_out/x11_bench_infer.json 2> gpurun_out/x11_bench_infer.err; head -c 330 gpurun_out/x11_bench_infer.json; echo
cp gpurun_out/kernel_table_infer_upernext.json gpurun_out/x11_kernel_table_infer_upernext.json
