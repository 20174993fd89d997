#!/bin/bash
# Final round-2 ncu evidence (one B200): launch list of the bench command + a --set full capture of the bench line's roofline kernel.
O=gpurun_out
TAG=r02e
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3690 -c 1200 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-eager-baseline > $O/${TAG}_ncu_bench.log 2>&1
python tools/profile_combine.py 1 fwd > $O/${TAG}_plain_hcfwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hc_fwd -s 0 -c 4 -f -o /tmp/prof_hcfwd python tools/profile_combine.py 1 fwd > $O/${TAG}_ncu_hcfwd.log 2>&1
ncu -i /tmp/prof_hcfwd.ncu-rep --page raw --csv > $O/${TAG}_ncu_hcfwd_raw.csv 2>/dev/null
ls -la $O | grep ${TAG} | tail -8; wc -l $O/launches_${TAG}.csv
