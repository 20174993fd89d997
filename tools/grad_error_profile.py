"""Diagnostic (GPU): per-parameter gradient error of a bf16 training step against the fp64 oracle, in execution order,
to see where along the backward chain the error grows.  python tools/grad_error_profile.py [upernext|fpn] [rough|precise|both]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkit_ocr_model_adaptive_scaling_b200 as vk  # noqa: E402
from oracle import loss as ol, model as om, synth  # noqa: E402

neck = sys.argv[1] if len(sys.argv) > 1 else 'upernext'
which = sys.argv[2] if len(sys.argv) > 2 else 'both'
dt = torch.bfloat16 if (len(sys.argv) <= 3 or sys.argv[3] == 'bf16') else torch.float32
dev = torch.device('cuda')
M, LF = vk.model, vk.loss_function
model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck)))
model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=7), strict=True)
model.to(dev).eval()
B, H, W, P = 2, 160, 224, 16
to = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
f64 = lambda d: {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in d.items()}
rb, pb = to(synth.synth_rough_batch(B, H, W, seed=3, inset=6)), to(synth.synth_precise_batch(B, H, W, points=P, seed=3, inset=6))
rk = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
pk = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
      'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles', 'char_corner_distances')
rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
params = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
with vk.precision(dt):
    if which in ('rough', 'both'):
        m, h = model.forward_rough(rb['image'])
        (rough_fn(rough_char_mask_feature=m, rough_char_height_feature=h, **{k: rb[k] for k in rk}) / 2).backward()
    if which in ('precise', 'both'):
        a, b, c, d = model.forward_precise(pb['image'])
        (precise_fn(None, a, b, c, d, **{k: pb[k] for k in pk}) / 2).backward()
rb, pb = f64(rb), f64(pb)
if which in ('rough', 'both'):
    mo, ho = om.forward_rough(params, rb['image'])
    print('fwd rough mask/height rel err', float((m.double() - mo).norm() / mo.norm()), float((h.double() - ho).norm() / ho.norm()))
    (ol.rough_loss(mo, ho, *(rb[k] for k in rk)) / 2).backward()
if which in ('precise', 'both'):
    outs = om.forward_precise(params, pb['image'])
    print('fwd precise rel err', [float((x.double() - y).norm() / y.norm()) for x, y in zip((a, b, c, d), outs)])
    (ol.precise_loss(None, *outs, *(pb[k] for k in pk)) / 2).backward()
num = den = 0.0
for name, p in model.named_parameters():
    r = params[name].grad
    if r is None or p.grad is None:
        continue
    e = float((p.grad.double() - r).norm())
    n = float(r.norm())
    num += e * e
    den += n * n
    if p.dim() > 1 or 'block_scale' in name or name.endswith('ln.1.weight') or 'stem' in name:
        print(f'{e / max(n, 1e-300):10.3e}  norm {n:10.3e}  {name}')
print('GLOBAL', (num / den) ** 0.5)
