#!/bin/bash
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -m gpu -k "label_point or sparse or points" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-eager-baseline --profile > $O/r02d_bench2.json 2> $O/r02d_bench2.err
head -c 300 $O/r02d_bench2.json; echo; grep -i "scatter_up\|unbracketed" $O/r02d_bench2.err | head -3
python -c "
import json; d=json.loads(open('$O/r02d_bench2.json').read().strip().splitlines()[-1]); print(d['unbracketed'], d['e2e']['value'], d['e2e']['ms_per_step'])"
