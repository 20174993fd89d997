"""GPU (N ranks under torchrun): where the gradient all-reduce sits on the timeline of the data-parallel training step.
Rank 0 profiles two steps with CUPTI timestamps (torch.profiler sees the C-ABI library's launches and NCCL's kernels) and
reports, per NCCL kernel: start / duration relative to the step, the share of it that ran concurrently with compute
kernels, and the part of the step's tail during which ONLY NCCL was running (= the exposed collective).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/overlap_probe.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import vkit_ocr_model_adaptive_scaling_b200 as vk  # noqa: E402

rank = int(os.environ.get('RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
vk.set_compute_dtype(torch.bfloat16)
w = bench.build_workload(vk, 'train', 'upernext', 32, 640, dev, rank, world)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(4):
    w.step(w.rb, w.pb)
barrier()
STEPS = 2
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        w.step(w.rb, w.pb)
    torch.cuda.synchronize()
barrier()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
    t0, t1 = iv[0][0], max(e for _, e, _ in iv)
    is_nccl = lambda n: 'nccl' in n.lower()
    comp = [(s, e) for s, e, n in iv if not is_nccl(n)]
    # merged busy intervals of the compute kernels
    merged = []
    for s, e in comp:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])

    def overlap(s, e):
        return sum(max(0.0, min(e, b) - max(s, a)) for a, b in merged)

    rows = []
    for s, e, n in iv:
        if is_nccl(n):
            rows.append({'kernel': n[:60], 'start_ms': (s - t0) / 1e3, 'ms': (e - s) / 1e3, 'overlapped_with_compute': overlap(s, e) / max(e - s, 1e-9)})
    nccl_total = sum(r['ms'] for r in rows)
    exposed = sum(r['ms'] * (1.0 - r['overlapped_with_compute']) for r in rows)
    out = {'world_size': world, 'steps_profiled': STEPS, 'span_ms': (t1 - t0) / 1e3, 'step_ms': (t1 - t0) / 1e3 / STEPS,
           'compute_busy_ms_per_step': sum(b - a for a, b in merged) / 1e3 / STEPS,
           'nccl_kernels_per_step': len(rows) / STEPS, 'nccl_ms_per_step': nccl_total / STEPS,
           'nccl_ms_not_overlapped_per_step': exposed / STEPS, 'nccl_kernels': rows}
    print(json.dumps({k: v for k, v in out.items() if k != 'nccl_kernels'}))
    for r in rows:
        print(f"  {r['start_ms']:9.3f} ms  +{r['ms']:7.3f} ms  overlapped {100 * r['overlapped_with_compute']:5.1f}%  {r['kernel']}")
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            json.dump(out, f, indent=1)
if world > 1:
    dist.destroy_process_group()
