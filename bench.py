#!/usr/bin/env python
"""Headline benchmark: training images/sec of the adaptive-scaling model at 640x640 (fwd + bwd + loss [+ gradient
all-reduce]) on N B200s, beside the reference algorithm's CPU path on the box's own host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--neck upernext|fpn]
                    [--workload train|backbone|infer] [--batch B] [--size S]

One "step" = the reference's two-pass training step (experiment/adaptive_scaling/train.py:397-478 minus data loading
and optimizer): forward_rough -> rough loss/2 -> backward, forward_precise -> precise loss/2 -> backward, on B rough +
B precise synthetic images per GPU; "images/sec" = B * N / step time (image pairs).  For N > 1 launch
with torchrun (one rank per GPU, NCCL); the gradient all-reduce is bucketed and overlapped (parallel.py).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed, max over ranks.  `e2e`: the same step
through the public API from pinned HOST buffers (per step one H2D copy of both batches, issued on a copy stream one step
ahead like a training input pipeline, and a D2H read of both losses; all inside the timed region).
`roofline`: the C-ABI call with the largest share of the step -- measured in a repetition of the timed region's K steps in
which every call is bracketed with CUDA events (the brackets cost 3-4 % of a step, so the timed region itself runs without
them) --, its ALGORITHMIC FLOPs or bytes over the measured duration against MEASURED_PEAKS.json (tensor peak for the tcgen05
GEMMs, HBM bandwidth for the streaming kernels).  `cpu_baseline`: the oracle (port of the reference algorithm, plain fp32 PyTorch) on
the host cores over a bounded sample.  `gpu_eager_baseline`: the same oracle under stock PyTorch eager on the SAME GPU
(fp32 TF32-off and autocast bf16).  `extra`: the other BASELINE configurations (FPN neck, ConvNeXt backbone only, 2048^2
rough inference) and the step with the fused clip + AdamW tail, each timed like `value` with fewer steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'train images/sec @640x640 (fwd+bwd+loss)'
UNIT = 'images/s'
POINTS = 200          # label points per image (train.py:58)
INSET = 10            # core-box inset of the reference's integration test (tests/test_adaptive_scaling.py:126-169)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--neck', default='upernext', choices=['upernext', 'fpn'])
    ap.add_argument('--workload', default='train', choices=['train', 'backbone', 'infer'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--size', type=int, default=None)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--profile', action='store_true', help='after timing, run one step with every C-ABI call bracketed '
                    'by CUDA events and write the per-kernel table to gpurun_out/kernel_table_<workload>.json')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cuda-graph', action='store_true', help='time the step replayed from one CUDA graph (training.GraphedTrainStep)')
    ap.add_argument('--no-extras', action='store_true', help='skip the FPN / optimizer / config #2 / config #5 side measurements')
    ap.add_argument('--no-eager-baseline', action='store_true', help='skip the stock-PyTorch-eager arm on the same GPU')
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': float(p['hbm_gbs']), 'tflops_burst': float(p['bf16_tflops']),
                'tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int) -> None:
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
        for _, line in rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ synthetic data
def make_batches(batch: int, size: int, seed: int):
    """Synthetic batches of the collate schema (dataset/adaptive_scaling.py:282-368), CPU tensors."""
    from oracle import synth
    rb = synth.synth_rough_batch(batch, size, size, seed=seed, inset=INSET)
    pb = synth.synth_precise_batch(batch, size, size, points=POINTS, seed=seed, inset=INSET)
    return rb, pb


def tensor_bytes(d) -> int:
    import torch
    return sum(v.numel() * v.element_size() for v in d.values() if isinstance(v, torch.Tensor))


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_step_time(neck: str, batch: int, size: int, steps: int, warmup: int, threads: int):
    """Oracle (restatement of the reference's algorithm in stock fp32 PyTorch) training step on the host cores."""
    import torch
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict('tiny', neck, seed=133)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rb, pb = make_batches(batch, size, 133)
    rk = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
    pk = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
          'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
          'char_corner_distances')

    def step():
        for p in params.values():
            p.grad = None
        rl = ol.rough_loss(*om.forward_rough(params, rb['image']), *(rb[k] for k in rk))
        (rl / 2).backward()
        pl = ol.precise_loss(None, *om.forward_precise(params, pb['image']), *(pb[k] for k in pk))
        (pl / 2).backward()
        return float(rl), float(pl)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(steps, 1)


def run_reference(args) -> None:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    size = args.size or 640
    sample_batch = 2
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    sec = cpu_step_time(args.neck, sample_batch, size, steps, warmup, threads)
    value = sample_batch / sec
    sample = (f'{sample_batch} image pairs of {size}x{size} per step (the workload\'s per-GPU batch is {args.batch or 32}); '
              f'{steps} timed step(s), {warmup} warm-up')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f'adaptive-scaling TINY/{args.neck.upper()} two-pass training step, {size}x{size}, fp32, host CPU',
                   'global_batch': sample_batch, 'image_size': size, 'label_points': POINTS},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ stock-eager GPU arm
def gpu_eager_baseline(neck: str, batch: int, size: int, dev, steps: int = 2):
    """The reference algorithm (the oracle's stock torch ops: cuDNN / cuBLAS / ATen eager) on the SAME B200 and the same
    step: fp32 with TF32 off (the reference's own arithmetic) and under torch.autocast(bfloat16).  A
    reported baseline beside the CPU arm; falls back to a smaller batch when the fp32 autograd tape does not fit."""
    import torch
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.training import PRECISE_KEYS, ROUGH_KEYS
    out = {'kind': 'port', 'what': 'oracle (reference algorithm, stock PyTorch eager ops) on the same GPU, same step'}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = synth.synth_state_dict('tiny', neck, seed=133)
        params = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}
        for b in (batch, batch // 2, batch // 4):
            try:
                rb, pb = make_batches(b, size, 133)
                to = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
                rb, pb = to(rb), to(pb)

                def step(autocast: bool):
                    for p in params.values():
                        p.grad = None
                    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                        m, h = om.forward_rough(params, rb['image'])
                    (ol.rough_loss(m.float(), h.float(), *(rb[k] for k in ROUGH_KEYS)) / 2).backward()
                    del m, h
                    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                        outs = om.forward_precise(params, pb['image'])
                    (ol.precise_loss(None, *(o.float() for o in outs), *(pb[k] for k in PRECISE_KEYS)) / 2).backward()

                for name, autocast in (('fp32', False), ('autocast_bf16', True)):
                    step(autocast)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        step(autocast)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    out[name] = {'value': b / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'batch': b}
                break
            except torch.cuda.OutOfMemoryError:
                for p in params.values():
                    p.grad = None
                torch.cuda.empty_cache()
                out['note'] = f'batch {b} did not fit beside the fp32 autograd tape; halved'
    except Exception as exc:   # a baseline must never take the headline number down with it
        out['error'] = f'{type(exc).__name__}: {exc}'[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


# ------------------------------------------------------------------------------------------------ our arm
class Workload:
    """One benchmark workload: model, device / pinned-host inputs and the step closure."""


def build_workload(vk, workload: str, neck: str, batch: int, size: int, dev, rank: int, world: int, with_optimizer: bool = False,
                   label_point_forward: bool = False, cuda_graph: bool = False):
    import torch
    from vkit_ocr_model_adaptive_scaling_b200.parallel import DataParallel
    from vkit_ocr_model_adaptive_scaling_b200.training import FusedAdamW, GraphedTrainStep, batch_to_device, train_step
    from oracle import synth  # synthetic weights / batches only (test infrastructure generating inputs, never on the timed path)
    M, LF = vk.model, vk.loss_function
    w = Workload()
    cfg = M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck))
    torch.manual_seed(133)
    model = M.AdaptiveScaling(cfg)
    model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=133), strict=True)
    model.to(dev)
    rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
    w.model, w.dp = model, None
    w.rb_host, w.pb_host, w.rb, w.pb, w.h2d = {}, {}, {}, {}, 0
    if workload != 'infer':
        rb_host, pb_host = make_batches(batch, size, 133 + rank)
        pin = lambda d: {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        w.rb_host, w.pb_host = pin(rb_host), pin(pb_host)
        w.rb, w.pb = batch_to_device(w.rb_host, dev), batch_to_device(w.pb_host, dev)
        w.h2d = tensor_bytes(w.rb_host) + tensor_bytes(w.pb_host)
    if workload == 'train':
        model.train()   # stochastic depth active, as in the reference loop (train.py:396)
        # flat gradient buckets also on one GPU: begin_step() zeroes 5 buffers instead of ~300 tensors, and for N > 1 the
        # bucketed all-reduce overlaps the backward passes
        dp = DataParallel(model, flatten_params=with_optimizer)
        w.dp = dp
        opt = FusedAdamW(dp.buckets, lr=8e-4, weight_decay=0.01, max_grad_norm=2.5) if with_optimizer else None

        def step(rbatch, pbatch):
            losses = train_step(model, rough_fn, precise_fn, rbatch, pbatch, dp, label_point_forward=label_point_forward)
            if opt is not None:     # clip_grad_norm_(2.5) + AdamW (train.py:468-478): outside BASELINE's metric, reported in `extra`
                opt.step()
            return losses
        if cuda_graph:              # the same step captured once and replayed (training.GraphedTrainStep): reported in `extra`
            assert opt is None
            step = GraphedTrainStep(model, rough_fn, precise_fn, w.rb, w.pb, dp, label_point_forward=label_point_forward)
        w.images_per_step = batch
        w.text = (f'adaptive-scaling TINY/{neck.upper()} two-pass training step (fwd+bwd+loss'
                  f'{"+bucketed NCCL grad all-reduce" if world > 1 else ""}{"+clip+AdamW" if with_optimizer else ""}'
                  f'{"; offset/angle/distance heads evaluated at the label points only" if label_point_forward else ""}'
                  f'{"; replayed from one CUDA graph" if cuda_graph else ""}), '
                  f'batch {batch}/GPU, {size}x{size}, {POINTS} label points')
        w.metric, w.unit = METRIC, UNIT
    elif workload == 'backbone':
        model.train()
        backbone = model.backbone

        def step(rbatch, pbatch):
            for p in backbone.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            feats = backbone(rbatch['image'])
            loss = sum(f.float().square().mean() for f in feats)
            loss.backward()
            return loss.detach(), loss.detach()
        w.images_per_step = batch
        w.text = f'ConvNeXt-T backbone forward/backward, batch {batch}/GPU, {size}x{size}'
        w.metric, w.unit = 'ConvNeXt backbone fwd+bwd images/sec @640x640', UNIT
    else:
        # config #5: uint8 page -> pad/ingest -> forward_rough -> thresholded uint8 mask + cleaned fp32 height map
        # (inferencing/adaptive_scaling.py:92-188, tensor side), independent replicas
        from vkit_ocr_model_adaptive_scaling_b200.inferencing import rough_infer_tensors
        model.eval()
        g = torch.Generator().manual_seed(133 + rank)
        pages_host = torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8).pin_memory()
        w.rb_host, w.pb_host = {'image_u8': pages_host}, {}
        w.rb, w.pb = {'image_u8': pages_host.to(dev)}, {}
        w.h2d = tensor_bytes(w.rb_host)

        def step(rbatch, pbatch):
            mask, hmap, _ = rough_infer_tensors(model, rbatch['image_u8'])
            return mask, hmap
        w.images_per_step = batch * size * size / 1e6          # the metric is MPix/s
        w.text = f'rough inference (uint8 page -> text mask + char-height map), batch {batch}/GPU, {size}x{size}, replicas'
        w.metric, w.unit = 'infer MPix/s', 'MPix/s'
    w.step = step
    w.workload, w.neck, w.batch, w.size = workload, neck, batch, size
    return w


def release(w) -> None:
    import gc
    import torch
    if hasattr(getattr(w, 'step', None), 'close'):
        w.step.close()          # a GraphedTrainStep: release the captured graph before anything else goes away
    if w.dp is not None:
        w.dp.close()
    for k in list(vars(w)):
        delattr(w, k)
    gc.collect()
    torch.cuda.empty_cache()


def time_device_resident(w, steps: int, warmup: int, dev, rank: int, world: int, local_rank: int, sample_clocks: bool = True):
    """W >= 3 warm-up steps, then exactly `steps` steps on device-resident inputs between barriers, CUDA events on the
    launching stream, max over ranks: the timed region.  The same `steps` steps are then repeated with every C-ABI call
    bracketed by CUDA events (the per-kernel table behind `roofline`): the brackets -- two cudaEventRecord per call, ~1500
    per step -- cost 3-4 % of a step, so they measure the kernels but stay out of the number they would distort."""
    import torch
    import torch.distributed as dist
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None   # before the warm-up: nvidia-smi needs ~0.5 s
    for _ in range(max(warmup, 3)):
        w.step(w.rb, w.pb)
    barrier()
    launches0 = L.LIB.vkocr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        losses = w.step(w.rb, w.pb)
    e1.record()
    barrier()
    t1 = time.time()
    launches = (L.LIB.vkocr_launch_count() - launches0) // max(steps, 1)
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    # bracketed repetition of the timed region
    L.LIB.start_profile()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    for _ in range(steps):
        w.step(w.rb, w.pb)
    e3.record()
    barrier()
    prof = L.LIB.stop_profile()
    ms_bracketed = e2.elapsed_time(e3) / steps
    if world > 1:
        t = torch.tensor([ms, ms_bracketed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_bracketed = float(t[0].item()), float(t[1].item())
    w.ms_bracketed = ms_bracketed
    return ms, losses, int(launches), prof.summary(), clocks


def time_end_to_end(w, steps: int, dev, rank: int, world: int, local_rank: int):
    """The same step through the public API from pinned HOST buffers: per step one H2D copy of the inputs (issued on a
    copy stream one step ahead, like a training input pipeline) and a D2H read of the result, all inside the timed region."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    # Two device-side staging sets, allocated once (an input pipeline's double buffer): no allocator traffic inside the
    # timed loop -- fresh side-stream allocations every step made the caching allocator fall back to cudaMalloc now and then.
    like = lambda d: {k: (torch.empty_like(v, device=dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    bufs = [(like(w.rb_host), like(w.pb_host)) for _ in range(2)]
    consumed = [None, None]                 # main-stream event: the step that read set k has finished
    turn = [0]

    def stage():
        k = turn[0]
        turn[0] ^= 1
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            for d_dev, d_host in zip(bufs[k], (w.rb_host, w.pb_host)):
                for name, v in d_host.items():
                    if isinstance(v, torch.Tensor):
                        d_dev[name].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs[k][0], bufs[k][1], ev, k

    def mark_consumed(k):
        consumed[k] = torch.cuda.Event()
        consumed[k].record(main_stream)

    for _ in range(2):
        r, p_, ev, k = stage()
        main_stream.wait_event(ev)
        w.step(r, p_)
        mark_consumed(k)
    result_host, d2h = None, 8
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        time.sleep(0.6)                                 # let nvidia-smi come up (before the barrier: every rank starts together)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    nxt = stage()                                       # step 0's copy is exposed; every later copy overlaps a step
    for i in range(steps):
        rdev, pdev, ev, k = nxt
        main_stream.wait_event(ev)
        if i + 1 < steps:
            nxt = stage()
        a, b = w.step(rdev, pdev)
        mark_consumed(k)
        if w.workload == 'infer':                       # the caller reads the uint8 mask and the height map
            if result_host is None:
                result_host = (torch.empty(a.shape, dtype=a.dtype).pin_memory(), torch.empty(b.shape, dtype=b.dtype).pin_memory())
            result_host[0].copy_(a, non_blocking=True)
            result_host[1].copy_(b, non_blocking=True)
            d2h = a.numel() * a.element_size() + b.numel() * b.element_size()
        else:
            loss_host.copy_(torch.stack([a.float().reshape(()), b.float().reshape(())]), non_blocking=True)
        main_stream.synchronize()                       # the caller reads the step's losses (train.py:415,453)
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {'value': w.images_per_step * world / (ms / 1e3), 'unit': w.unit, 'h2d_bytes_per_step': w.h2d, 'd2h_bytes_per_step': d2h,
            'ms_per_step': ms, 'clocks': clocks}


def roofline_of(table, steps: int, step_ms: float, peaks):
    """The C-ABI call (entry point + shape) with the largest share of the timed region, against the roofline that bounds it:
    tensor FLOP/s for the tcgen05 GEMMs, HBM GB/s for everything else.  FLOPs / bytes are ALGORITHMIC (true channel counts,
    each operand read once and each result written once), the duration is measured here with CUDA events."""
    rows = {k: r for k, r in table.items() if r['flops'] > 0 or r['bytes'] > 0}
    if not rows:
        return None
    label, row = max(rows.items(), key=lambda kv: kv[1]['ms'])
    per_launch_ms = row['ms'] / row['calls']
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            ent = json.load(f).get(label)      # DRAM bytes of this very call from the committed ncu --set full capture
            traffic = ent.get('dram_bytes_per_call') if isinstance(ent, dict) else ent
    except OSError:
        pass
    gemm = {k: r for k, r in table.items() if r['flops'] > 0 and r['entry'].startswith('vkocr_gemm')}
    common = {'kernel': f'{row["entry"]} [{label}]', 'launch_ms': per_launch_ms, 'share_of_step': row['ms'] / steps / step_ms,
              'traffic': traffic,
              'all_gemm_share_of_step': sum(r['ms'] for r in gemm.values()) / steps / step_ms if gemm else 0.0,
              'all_gemm_tflops': (sum(r['flops'] for r in gemm.values()) / sum(r['ms'] for r in gemm.values()) / 1e9) if gemm else 0.0}
    if row['flops'] > 0 and row['entry'].startswith('vkocr_gemm'):
        achieved = row['flops'] / row['calls'] / (per_launch_ms * 1e-3) / 1e12
        return dict(bound='tensor', achieved=achieved, peak=peaks['tflops_sustained'], unit='TFLOP/s', frac=achieved / peaks['tflops_sustained'],
                    algorithmic_flops=row['flops'] / row['calls'],
                    peak_source=f'{peaks["source"]} sustained bf16 (kernel timed inside a long step)', **common)
    achieved = row['bytes'] / row['calls'] / (per_launch_ms * 1e-3) / 1e9
    return dict(bound='hbm', achieved=achieved, peak=peaks['hbm_gbs'], unit='GB/s', frac=achieved / peaks['hbm_gbs'],
                algorithmic_bytes=row['bytes'] / row['calls'], peak_source=f'{peaks["source"]} HBM copy bandwidth', **common)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert L.LIB.vkocr_device_check(local_rank) == 0, L.LIB.vkocr_last_error()
    peaks = load_peaks()
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    vk.set_compute_dtype(dtype)

    size = args.size or (2048 if args.workload == 'infer' else 640)
    batch = args.batch or (8 if args.workload == 'infer' else 32)
    w = build_workload(vk, args.workload, args.neck, batch, size, dev, rank, world, cuda_graph=args.cuda_graph and args.workload == 'train')
    ms, losses, launches, table, clocks = time_device_resident(w, args.steps, args.warmup, dev, rank, world, local_rank)
    value = w.images_per_step * world / (ms / 1e3)
    ms_bracketed = w.ms_bracketed
    e2e = None if args.no_e2e else time_end_to_end(w, args.steps, dev, rank, world, local_rank)
    loss_values = [float(x.float().sum()) for x in losses]

    # ---- optional: full per-kernel table of the timed region
    if args.profile and rank == 0:
        rows = sorted(table.items(), key=lambda kv: -kv[1]['ms'])
        total = sum(r['ms'] for _, r in rows) / args.steps
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'kernel_table_{args.workload}_{args.neck}.json'), 'w') as f:
            json.dump({'step_ms': ms, 'step_ms_bracketed': ms_bracketed, 'step_ms_sum_of_calls': total, 'steps': args.steps,
                       'rows': [dict(label=k, calls=v['calls'] / args.steps, ms=v['ms'] / args.steps, flops=v['flops'] / args.steps,
                                     bytes=v['bytes'] / args.steps, entry=v['entry']) for k, v in rows]}, f, indent=1)
        print(f'# per-call device time per step: {total:.2f} ms of {ms_bracketed:.2f} ms (bracketed repetition; timed region {ms:.2f} ms)',
              file=sys.stderr)
        for k, r in rows[:70]:
            tf = r['flops'] / (r['ms'] * 1e-3) / 1e12 if r['flops'] else 0.0
            gb = r['bytes'] / (r['ms'] * 1e-3) / 1e9 if r['bytes'] else 0.0
            print(f'#  {r["ms"] / args.steps:9.3f} ms {100 * r["ms"] / args.steps / total:5.1f}%  x{r["calls"] // args.steps:<4d} {tf:7.1f} TF/s {gb:7.0f} GB/s  {k}',
                  file=sys.stderr)

    workload_text, metric, unit, h2d = w.text, w.metric, w.unit, w.h2d
    release(w)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # shares are taken against the bracketed repetition the per-call times come from
    roofline = roofline_of(table, args.steps, ms_bracketed, peaks)
    if roofline is not None:
        roofline['bracketed_ms_per_step'] = ms_bracketed

    # ---- the other BASELINE configurations and the optimizer tail, observed by the same run (N = 1 only)
    extra = None
    if world == 1 and args.workload == 'train' and not args.no_extras:
        extra = {}
        for key, (wl, neck, b, sz, opt, lpf, graph) in {
                'fpn_train': ('train', 'fpn' if args.neck == 'upernext' else 'upernext', batch, size, False, False, False),
                'with_optimizer': ('train', args.neck, batch, size, True, False, False),
                'label_point_forward': ('train', args.neck, batch, size, False, True, False),
                'cuda_graph': ('train', args.neck, batch, size, False, False, True),
                'backbone_config2': ('backbone', args.neck, 32, 640, False, False, False),
                'infer_config5': ('infer', 'upernext', 8, 2048, False, False, False)}.items():
            try:
                wx = build_workload(vk, wl, neck, b, sz, dev, 0, 1, with_optimizer=opt, label_point_forward=lpf, cuda_graph=graph)
                msx, _, lx, _, _ = time_device_resident(wx, 5, 3, dev, 0, 1, local_rank, sample_clocks=False)
                extra[key] = {'workload': wx.text, 'value': wx.images_per_step / (msx / 1e3), 'unit': wx.unit, 'ms_per_step': msx,
                              'gpu_launches': lx}
                release(wx)
            except Exception as exc:
                extra[key] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
                torch.cuda.empty_cache()

    gpu_eager = None
    if world == 1 and args.workload == 'train' and not args.no_eager_baseline:
        gpu_eager = gpu_eager_baseline(args.neck, batch, size, dev)
        torch.cuda.empty_cache()

    cpu_baseline = None
    if not args.no_cpu_baseline and args.workload == 'train':
        threads = os.cpu_count() or 1
        sec = cpu_step_time(args.neck, 2, size, 1, 0, threads)
        cpu_baseline = {'value': 2 / sec, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                        'sample': f'2 image pairs of {size}x{size}, one fp32 training step of the oracle (reference algorithm) on the host cores'}

    line = {
        'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if dtype == torch.bfloat16 else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text, 'global_batch': batch * world, 'image_size': size, 'label_points': POINTS,
                   'parallelism': f'dp{world}', 'l2': f'inputs per step ({h2d / 1e6:.0f} MB) and every activation exceed the 126 MB L2',
                   'unit_note': 'images/s counts image PAIRS (a step consumes B rough + B precise images, SURVEY 8d); '
                                f'image-forwards/s = 2 x value = {2 * value:.1f}' if args.workload == 'train' else None},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'gpu_eager_baseline': gpu_eager, 'extra': extra, 'losses': loss_values,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
