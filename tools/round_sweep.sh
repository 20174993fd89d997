#!/bin/bash
# Round-end measurement sweep on one B200 (run through gpurun from the repo root): parity tests, every bench line, the
# ncu launch list of the bench command and `ncu --set full` captures of the dominant kernels.  Outputs: gpurun_out/<tag>_*.
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu > $O/${TAG}_tests.log 2>&1; tail -2 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>&1
python bench.py --neck fpn --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_fpn.json 2>/dev/null
python bench.py --workload backbone --steps 5 --warmup 3 > $O/${TAG}_bench_backbone.json 2>/dev/null
python bench.py --workload infer --steps 5 --warmup 3 > $O/${TAG}_bench_infer.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile > $O/${TAG}_bench_profile.log 2>&1
cp $O/kernel_table_train_upernext.json $O/${TAG}_kernel_table_train_upernext.json
# launch list of the bench command (only after it exited 0 without ncu above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
# full captures of the dominant kernels (each tool exits 0 without ncu first)
python tools/profile_head_conv.py 3 > $O/${TAG}_plain_head.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vkocr_gemm_tc_kernel -s 4 -c 2 -f -o $O/prof_head_${TAG} \
    python tools/profile_head_conv.py 3 > $O/${TAG}_ncu_head.log 2>&1
python tools/profile_dw.py 2 > $O/${TAG}_plain_dw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dwconv7 -s 2 -c 2 -f -o $O/prof_dw_${TAG} \
    python tools/profile_dw.py 2 > $O/${TAG}_ncu_dw.log 2>&1
python tools/kbench.py head > $O/${TAG}_kbench_head.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_tail_bwd_kernel -s 3 -c 2 -f -o $O/prof_htb_${TAG} \
    python tools/kbench.py head > $O/${TAG}_ncu_htb.log 2>&1
python tools/kbench.py mlp dw ln colsum up > $O/${TAG}_kbench.log 2>&1
ls -la $O | grep ${TAG} | head -40
