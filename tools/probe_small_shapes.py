"""GPU: bf16 gradient error (global rel L2 against the fp64 oracle) of the two-pass step for the SMALL / BASE / LARGE and FPN
configurations at three image sizes, two runs each (atomics make the sums order-dependent).  Evidence for DESIGN.md section 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import vkit_ocr_model_adaptive_scaling_b200 as vk
from oracle import loss as ol, model as om, synth
from vkit_ocr_model_adaptive_scaling_b200.training import train_step, ROUGH_KEYS, PRECISE_KEYS
from _util import oracle_params, rel_err
import test_gpu_model as T
dev = torch.device('cuda')
for size, neck in (('small', 'upernext'), ('tiny', 'fpn'), ('base', 'fpn'), ('large', 'upernext')):
    for (B, H, W, P) in ((1, 64, 96, 8), (2, 96, 128, 8), (2, 128, 192, 16)):
        errs = []
        for rep in range(2):
            model = T._build(vk, neck, size)
            model.load_state_dict(synth.synth_state_dict(size, neck, seed=11), strict=True)
            model.to(dev).eval()
            params = oracle_params(model)
            rb = T._to(synth.synth_rough_batch(B, H, W, seed=5, inset=2), dev)
            pb = T._to(synth.synth_precise_batch(B, H, W, points=P, seed=5, inset=2), dev)
            lf = vk.loss_function
            rough_fn = lf.AdaptiveScalingRoughLossFunction(lf.AdaptiveScalingRoughLossFunctionConifg())
            precise_fn = lf.AdaptiveScalingPreciseLossFunction(lf.AdaptiveScalingPreciseLossFunctionConifg())
            with vk.precision(torch.bfloat16):
                rl, pl = train_step(model, rough_fn, precise_fn, rb, pb)
            f64 = lambda d: {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in d.items()}
            rb, pb = f64(rb), f64(pb)
            rl_ref = ol.rough_loss(*om.forward_rough(params, rb['image']), *(rb[k] for k in ROUGH_KEYS))
            (rl_ref / 2).backward()
            pl_ref = ol.precise_loss(None, *om.forward_precise(params, pb['image']), *(pb[k] for k in PRECISE_KEYS))
            (pl_ref / 2).backward()
            g = torch.cat([p.grad.double().flatten() for _, p in model.named_parameters()])
            r = torch.cat([params[n].grad.double().flatten() for n, _ in model.named_parameters()])
            errs.append(float((g - r).norm() / r.norm()))
        print(f'{size}/{neck} B{B} {H}x{W}: gradient rel L2 ' + ' '.join(f'{e:.3e}' for e in errs), flush=True)
