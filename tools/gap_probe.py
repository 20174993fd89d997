"""GPU: kernel-to-kernel idle gaps of one training step, from CUPTI timestamps (torch.profiler sees every launch of the
process, including the C-ABI library's).  Answers whether whole-step CUDA graphs / PDL could buy anything."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vkit_ocr_model_adaptive_scaling_b200 as vk
from vkit_ocr_model_adaptive_scaling_b200.parallel import DataParallel
from vkit_ocr_model_adaptive_scaling_b200.training import batch_to_device, train_step
from oracle import synth
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
vk.set_compute_dtype(torch.bfloat16)
M, LF = vk.model, vk.loss_function
model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType('upernext')))
model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=133), strict=True)
model.to(dev).train()
rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
rbh, pbh = bench.make_batches(32, 640, 133)
rb, pb = batch_to_device(rbh, dev), batch_to_device(pbh, dev)
dp = DataParallel(model)
step = lambda: train_step(model, rough_fn, precise_fn, rb, pb, dp)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
busy = 0.0
gaps = []
cur_end = iv[0][0]
for s, e, n in iv:
    if s > cur_end:
        gaps.append((s - cur_end, n))
    if e > cur_end:
        busy += e - max(s, cur_end)
        cur_end = e
span = iv[-1][1] - iv[0][0]
print(f'{len(iv)} device activities over {span / 1e3:.2f} ms (2 steps): busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms = {100 * (span - busy) / span:.1f} %')
g = sorted(x[0] for x in gaps)
import statistics
print(f'{len(g)} gaps: median {statistics.median(g):.1f} us, mean {sum(g) / len(g):.1f} us, p90 {g[int(.9 * len(g))]:.1f} us, max {g[-1]:.1f} us; gaps > 20 us: {sum(1 for x in g if x > 20)} totalling {sum(x for x in g if x > 20) / 1e3:.2f} ms')
big = sorted(gaps, reverse=True)[:12]
for d, n in big:
    print(f'   {d:8.1f} us before {n[:90]}')
