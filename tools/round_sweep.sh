#!/bin/bash
# Round measurement sweep on one B200 (run through gpurun from the repo root).
#   tools/round_sweep.sh <tag> bench     parity tests, smoke, every bench line                      (~2.5 min)
#   tools/round_sweep.sh <tag> ncu       launch list of the bench command + ncu --set full captures  (~6 min)
# Outputs: gpurun_out/<tag>_*.  gpurun copies back at most 64 MiB IN TOTAL, so the ncu reports are exported to CSV on the box
# (raw metrics of every captured launch) and stay there.
TAG=${1:-r02}
WHAT=${2:-bench}
O=gpurun_out
mkdir -p $O

if [ "$WHAT" = bench ]; then
    python -m pytest tests -q -m gpu > $O/${TAG}_tests.log 2>&1; tail -2 $O/${TAG}_tests.log
    python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
    # the headline line: value, e2e, roofline, cpu baseline, stock-eager GPU baseline, extras (FPN / optimizer / config #2 / #5)
    python bench.py --steps 10 --warmup 3 --profile > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; head -c 700 $O/${TAG}_bench.json; echo
    cp $O/kernel_table_train_upernext.json $O/${TAG}_kernel_table_train_upernext.json
    python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>&1
    python bench.py --neck fpn --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-eager-baseline --profile > $O/${TAG}_bench_fpn.json 2> $O/${TAG}_bench_fpn.err
    cp $O/kernel_table_train_fpn.json $O/${TAG}_kernel_table_train_fpn.json
    python bench.py --workload backbone --steps 10 --warmup 3 --profile > $O/${TAG}_bench_backbone.json 2> $O/${TAG}_bench_backbone.err
    python bench.py --workload infer --steps 10 --warmup 3 --profile > $O/${TAG}_bench_infer.json 2> $O/${TAG}_bench_infer.err
    cp $O/kernel_table_infer_upernext.json $O/${TAG}_kernel_table_infer_upernext.json
    python tools/kbench.py mlp dw ln colsum up head > $O/${TAG}_kbench.log 2>&1
    python tools/gap_probe.py > $O/${TAG}_gap_probe.log 2>&1
fi

if [ "$WHAT" = ncu ]; then
    # launch list of the bench command (the same command exited 0 without ncu in the `bench` pass)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3690 -c 1200 --csv --log-file $O/launches_${TAG}.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-eager-baseline > $O/${TAG}_ncu_bench.log 2>&1
    capture() {   # name, kernel regex, launches to skip, launches to capture, command...
        local name=$1 regex=$2 skip=$3 count=$4; shift 4
        "$@" > $O/${TAG}_plain_${name}.log 2>&1 || return
        timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f \
            -o /tmp/prof_${name}_${TAG} "$@" > $O/${TAG}_ncu_${name}.log 2>&1
        ncu -i /tmp/prof_${name}_${TAG}.ncu-rep --page raw --csv > $O/prof_${name}_${TAG}_raw.csv 2>/dev/null
        # the reports themselves (5-12 MB each) stay on the box: gpurun copies back at most 64 MiB in total
    }
    capture hcfwd 'hc_fwd' 0 4 python tools/profile_combine.py 1 fwd
    capture hcfwd_rough 'hc_fwd' 0 2 env ROUGH=1 python tools/profile_combine.py 1 fwd
    capture zgemm 'vkocr_gemm_tc' 1 1 python tools/profile_combine.py 2 z
    capture htb 'head_tail_bwd_h192' 3 1 python tools/kbench.py head
    capture dw 'dwconv7' 2 2 python tools/profile_dw.py 2
fi
du -sh $O; ls -la $O | grep ${TAG} | head -60
