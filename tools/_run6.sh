mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/x6_tests.log 2>&1; tail -3 gpurun_out/x6_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-eager-baseline --profile > gpurun_out/x6_bench.json 2> gpurun_out/x6_bench.err; head -c 600 gpurun_out/x6_bench.json; echo
cp gpurun_out/kernel_table_train_upernext.json gpurun_out/x6_kernel_table_train_upernext.json
