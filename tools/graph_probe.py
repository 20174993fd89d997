"""GPU: the two-pass training step replayed from a CUDA graph (training.GraphedTrainStep) against the eager step.
1. eval mode (no stochastic depth): the gradients of a replay equal the eager step's (up to the order of the atomics);
2. train mode: ms per step, eager vs replay, at BASELINE config #3 (B=32, 640x640).
python tools/graph_probe.py [steps] [neck]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vkit_ocr_model_adaptive_scaling_b200 as vk
from vkit_ocr_model_adaptive_scaling_b200.training import GraphedTrainStep, train_step

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
neck = sys.argv[2] if len(sys.argv) > 2 else 'upernext'
batch = int(os.environ.get('BATCH', '32'))
size = int(os.environ.get('SIZE', '640'))
rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:        # torchrun: the bucketed NCCL all-reduces are captured into the graph too
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
vk.set_compute_dtype(torch.bfloat16)
w = bench.build_workload(vk, 'train', neck, batch, size, dev, rank, world)
model, dp = w.model, w.dp
LF = vk.loss_function
rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


out = {}
model.eval()
for _ in range(2):
    eager_losses = train_step(model, rough_fn, precise_fn, w.rb, w.pb, dp)
eager_grads = [f.clone() for f in dp.buckets.flat]
g_eval = GraphedTrainStep(model, rough_fn, precise_fn, w.rb, w.pb, dp, warmup=1)
for f in dp.buckets.flat:
    f.fill_(float('nan'))          # the replay must zero and refill the buckets itself
graph_losses = g_eval(w.rb, w.pb)
torch.cuda.synchronize()
num = sum(float((a.double() - b.double()).square().sum()) for a, b in zip(dp.buckets.flat, eager_grads))
den = sum(float(b.double().square().sum()) for b in eager_grads)
out['eval_grad_rel_l2_graph_vs_eager'] = (num / den) ** 0.5
out['eval_losses_eager'] = [float(x) for x in eager_losses]
out['eval_losses_graph'] = [float(x) for x in graph_losses]
# other inputs through the same graph: swap the two halves of the batch, eager vs replay
perm = torch.arange(batch, device=dev).flip(0)
rb2 = {k: (v[perm] if isinstance(v, torch.Tensor) else v) for k, v in w.rb.items()}
pb2 = {k: (v[perm] if isinstance(v, torch.Tensor) else v) for k, v in w.pb.items()}
l_e = [float(x) for x in train_step(model, rough_fn, precise_fn, rb2, pb2, dp)]
l_g = [float(x) for x in g_eval(rb2, pb2)]
out['eval_losses_other_batch_eager_graph'] = [l_e, l_g]
g_eval.close()
del g_eval
torch.cuda.empty_cache()

model.train()
for _ in range(3):
    train_step(model, rough_fn, precise_fn, w.rb, w.pb, dp)
out['eager_ms'], _ = timed(lambda: train_step(model, rough_fn, precise_fn, w.rb, w.pb, dp), steps)
g_train = GraphedTrainStep(model, rough_fn, precise_fn, w.rb, w.pb, dp, warmup=3)
for _ in range(2):
    g_train(w.rb, w.pb)
out['graph_ms'], losses = timed(lambda: g_train(w.rb, w.pb), steps)
out['graph_losses_train_mode'] = [float(x) for x in losses]
out['config'] = {'neck': neck, 'batch': batch, 'size': size, 'steps': steps, 'world': world, 'rank': rank}
print(json.dumps(out), flush=True)
g_train.close()        # a live graph that captured NCCL kernels makes destroy_process_group() wait forever
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
