"""Diagnostic (GPU): where the bf16 forward error comes from.  Every module (backbone, neck, heads) is run twice: on
its own upstream result ("chained") and on the fp64 oracle's upstream result rounded once to fp32 ("isolated"), and
compared with the fp64 oracle.   python tools/fwd_error_profile.py [upernext|fpn]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkit_ocr_model_adaptive_scaling_b200 as vk  # noqa: E402
from oracle import model as om, synth  # noqa: E402

neck = sys.argv[1] if len(sys.argv) > 1 else 'fpn'
dev = torch.device('cuda')
M = vk.model
model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck)))
model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=7), strict=True)
model.to(dev).eval()
B, H, W = 2, 160, 224
rb = synth.synth_rough_batch(B, H, W, seed=3, inset=6)
img = rb['image'].to(dev)
params = {k: v.detach().double() for k, v in model.state_dict().items()}


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


with torch.no_grad():
    feats_o = om.convnext_forward(params, img.double())
    neck_o = om.neck_forward(params, 'rough_neck.', feats_o, neck)
    mask_o = om.head_forward(params, 'rough_char_mask_head.', neck_o, neck, 2)
    hgt_o = om.head_forward(params, 'rough_char_height_head.0.', neck_o, neck, 2, softplus=True)
    for dt in (torch.float32, torch.bfloat16):
        with vk.precision(dt):
            feats = model.backbone(img)
            print(f'[{dt}] backbone features rel err', [f'{rel(a, b):.2e}' for a, b in zip(feats, feats_o)])
            nk = model.rough_neck(feats)
            print(f'[{dt}] neck chained  {rel(nk, neck_o):.2e}')
            nk_iso = model.rough_neck([f.float() for f in feats_o])
            print(f'[{dt}] neck isolated {rel(nk_iso, neck_o):.2e}')
            mk = model.rough_char_mask_head(nk)
            hg = model.rough_char_height_head(nk)
            print(f'[{dt}] heads chained  mask {rel(mk, mask_o):.2e} height {rel(hg, hgt_o):.2e}')
            mk = model.rough_char_mask_head(neck_o.float())
            hg = model.rough_char_height_head(neck_o.float())
            print(f'[{dt}] heads isolated mask {rel(mk, mask_o):.2e} height {rel(hg, hgt_o):.2e}')
    print('mask logits: rms', float(mask_o.square().mean().sqrt()), 'mean', float(mask_o.mean()))
