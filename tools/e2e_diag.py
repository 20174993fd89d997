"""Diagnostic (GPU): where does the end-to-end step (host batches in, losses out) lose time against the device-resident
step?  Times: pinned H2D copy alone, CPU-side issue time of one step, synchronised steps with / without staging."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vkit_ocr_model_adaptive_scaling_b200 as vk
from vkit_ocr_model_adaptive_scaling_b200.parallel import DataParallel
from vkit_ocr_model_adaptive_scaling_b200.training import batch_to_device, train_step
from oracle import synth
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
vk.set_compute_dtype(torch.bfloat16)
M, LF = vk.model, vk.loss_function
neck = sys.argv[1] if len(sys.argv) > 1 else 'upernext'
model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck)))
model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=133), strict=True)
model.to(dev).train()
rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
rbh, pbh = bench.make_batches(32, 640, 133)
pin = lambda d: {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
rbh, pbh = pin(rbh), pin(pbh)
rb, pb = batch_to_device(rbh, dev), batch_to_device(pbh, dev)
dp = DataParallel(model)
step = lambda r, p: train_step(model, rough_fn, precise_fn, r, p, dp)
for _ in range(3):
    step(rb, pb)
torch.cuda.synchronize()
nbytes = bench.tensor_bytes(rbh) + bench.tensor_bytes(pbh)
for _ in range(3):
    t0 = time.perf_counter(); a = batch_to_device(rbh, dev); b = batch_to_device(pbh, dev); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f'H2D alone: {nbytes / 1e6:.0f} MB in {(t1 - t0) * 1e3:.1f} ms = {nbytes / (t1 - t0) / 1e9:.1f} GB/s', flush=True)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(rb, pb); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'one synchronised step: CPU issue {(t1 - t0) * 1e3:.1f} ms, until the GPU is done {(t2 - t0) * 1e3:.1f} ms', flush=True)
t0 = time.perf_counter()
for _ in range(5):
    step(rb, pb)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f'5 steps back to back: {(t1 - t0) / 5 * 1e3:.1f} ms per step', flush=True)
cs = torch.cuda.Stream(); ms = torch.cuda.current_stream()
def stage():
    with torch.cuda.stream(cs):
        r, p_ = batch_to_device(rbh, dev), batch_to_device(pbh, dev)
        ev = torch.cuda.Event(); ev.record(cs)
    for d in (r, p_):
        for v in d.values():
            if isinstance(v, torch.Tensor):
                v.record_stream(ms)
    return r, p_, ev
for label, do_stage, do_sync in (('sync each step, no staging', False, True), ('staging, no per-step sync', True, False), ('staging + sync each step (bench e2e)', True, True)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nxt = stage() if do_stage else (rb, pb, None)
    for i in range(5):
        r, p_, ev = nxt
        if ev is not None: ms.wait_event(ev)
        if do_stage and i + 1 < 5: nxt = stage()
        a, b = step(r, p_)
        if do_sync: ms.synchronize()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f'{label}: {(t1 - t0) / 5 * 1e3:.1f} ms per step', flush=True)
