"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded synthetic inputs.

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Weights come from
``oracle.synth.synth_state_dict`` (loaded into the reference modules with ``strict=True``, which also checks that
the synthetic key/shape layout equals the reference's ``state_dict()``).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, '_stub'))

from vkit.element import Box  # noqa: E402  (stub)
from vkit_open_model.model import (  # noqa: E402
    AdaptiveScaling, AdaptiveScalingConfig, AdaptiveScalingSize, AdaptiveScalingNeckHeadType, ConvNext,
    UperNextNeck, UperNextHead, FpnNeck, FpnHead,
)
from vkit_open_model import loss_function as ref_loss  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def _ref_box(b):
    return Box(up=b.up, down=b.down, left=b.left, right=b.right)


def full_model(neck: str, height: int = 96, width: int = 128, batch: int = 2, points: int = 20):
    cfg = AdaptiveScalingConfig(
        size=AdaptiveScalingSize.TINY,
        neck_head_type=AdaptiveScalingNeckHeadType.UPERNEXT if neck == 'upernext' else AdaptiveScalingNeckHeadType.FPN,
    )
    torch.manual_seed(0)
    model = AdaptiveScaling(cfg)
    sd = synth.synth_state_dict('tiny', neck, seed=133)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), 'synthetic state_dict key order differs from the reference'
    model.load_state_dict(sd, strict=True)
    model.eval()  # no stochastic depth: deterministic parity (convnext.py:41-53)
    out = {}
    rb = synth.synth_rough_batch(batch, height, width, seed=133, inset=4)
    pb = synth.synth_precise_batch(batch, height, width, points=points, seed=133, inset=4)

    mask, hgt = model.forward_rough(rb['image'])
    rl = ref_loss.AdaptiveScalingRoughLossFunction(ref_loss.AdaptiveScalingRoughLossFunctionConifg())(
        rough_char_mask_feature=mask, rough_char_height_feature=hgt,
        downsampled_mask=rb['downsampled_mask'].clone(), downsampled_score_map=rb['downsampled_score_map'].clone(),
        downsampled_shape=rb['downsampled_shape'], downsampled_core_box=_ref_box(rb['downsampled_core_box']))
    (rl / 2).backward()
    out['rough_mask'] = _np(mask)
    out['rough_height'] = _np(hgt)
    out['rough_loss'] = _np(rl)
    names = [n for n, _ in model.named_parameters()]
    out['rough_grad_norm'] = np.array(
        [float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()], dtype=np.float64)
    out['rough_grad_sum'] = np.array(
        [float(p.grad.double().sum()) if p.grad is not None else 0.0 for _, p in model.named_parameters()], dtype=np.float64)
    model.zero_grad(set_to_none=True)

    prob, off, ang, dist = model.forward_precise(pb['image'])
    pl = ref_loss.AdaptiveScalingPreciseLossFunction(ref_loss.AdaptiveScalingPreciseLossFunctionConifg())(
        precise_char_mask_feature=None, precise_char_prob_feature=prob,
        precise_char_up_left_corner_offset_feature=off, precise_char_corner_angle_feature=ang,
        precise_char_corner_distance_feature=dist,
        downsampled_char_prob_score_map=pb['downsampled_char_prob_score_map'].clone(),
        downsampled_char_mask=pb['downsampled_char_mask'].clone(),
        downsampled_shape=pb['downsampled_shape'], downsampled_core_box=_ref_box(pb['downsampled_core_box']),
        downsampled_label_point_y=pb['downsampled_label_point_y'], downsampled_label_point_x=pb['downsampled_label_point_x'],
        char_up_left_offsets=pb['char_up_left_offsets'], char_corner_angles=pb['char_corner_angles'],
        char_corner_distances=pb['char_corner_distances'])
    (pl / 2).backward()
    out['precise_prob'] = _np(prob)
    out['precise_offset'] = _np(off)
    out['precise_angle'] = _np(ang)
    out['precise_distance'] = _np(dist)
    out['precise_loss'] = _np(pl)
    out['precise_grad_norm'] = np.array(
        [float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()], dtype=np.float64)
    out['precise_grad_sum'] = np.array(
        [float(p.grad.double().sum()) if p.grad is not None else 0.0 for _, p in model.named_parameters()], dtype=np.float64)
    # a few complete gradients (small tensors) to pin backward element-wise
    for n, p in model.named_parameters():
        if n in ('backbone.stem.0.bias', 'backbone.blocks.0.layers.0.block_scale', 'backbone.blocks.3.ln.1.weight',
                 'precise_char_corner_angle_head.step2_conv1x1.1.weight', 'precise_char_corner_angle_head.step2_conv.1.weight',
                 'backbone.blocks.1.layers.2.block.0.weight'):
            out['precise_grad::' + n] = _np(p.grad)
    out['param_names'] = np.array(names)
    out['meta'] = np.array([batch, height, width, points, 4], dtype=np.int64)  # inset = 4
    np.savez_compressed(os.path.join(OUT, f'adaptive_scaling_tiny_{neck}.npz'), **out)
    print(neck, 'rough loss', float(rl), 'precise loss', float(pl))


def backbone_features():
    torch.manual_seed(0)
    model = ConvNext.create_tiny()
    g = synth._Gen(7)
    synth.backbone_state_dict(g, *synth.SIZES['tiny'], prefix='')
    model.load_state_dict(g.sd, strict=True)
    model.eval()
    x = synth.synth_image(1, 64, 96, seed=5)
    feats = model(x)
    # reference shape facts (tests/test_convnext.py:46-50 scaled to 64x96)
    assert [tuple(f.shape) for f in feats] == [(1, 96, 16, 24), (1, 192, 8, 12), (1, 384, 4, 6), (1, 768, 2, 3)]
    np.savez_compressed(os.path.join(OUT, 'convnext_tiny_features.npz'), **{f'f{i}': _np(f) for i, f in enumerate(feats)})


def primitive_losses():
    g = torch.Generator().manual_seed(11)
    out = {}
    pred = torch.randn(4, 25, generator=g) * 2
    gt01 = (torch.rand(4, 25, generator=g) > 0.5).float()
    gtf = torch.rand(4, 25, generator=g)
    mask = (torch.rand(4, 25, generator=g) > 0.3).float()
    out.update(pred=_np(pred), gt01=_np(gt01), gtf=_np(gtf), mask=_np(mask))
    out['focal'] = _np(ref_loss.FocalWithLogitsLossFunction()(pred.clone(), gt01.clone()))
    out['focal_masked'] = _np(ref_loss.FocalWithLogitsLossFunction()(pred.clone(), gt01.clone(), mask.clone()))
    out['dice'] = _np(ref_loss.DiceLossFunction()(torch.sigmoid(pred), gt01.clone()))
    out['dice_masked'] = _np(ref_loss.DiceLossFunction()(torch.sigmoid(pred), gt01.clone(), mask.clone()))
    out['l1'] = _np(ref_loss.L1LossFunction()(pred.clone(), gtf.clone()))
    out['l1_masked'] = _np(ref_loss.L1LossFunction()(pred.clone(), gtf.clone(), mask.clone()))
    out['smooth_l1'] = _np(ref_loss.L1LossFunction(smooth=True, smooth_beta=2.5)(pred.clone(), gtf.clone()))
    out['smooth_l1_masked'] = _np(ref_loss.L1LossFunction(smooth=True)(pred.clone(), gtf.clone(), mask.clone()))
    out['l2'] = _np(ref_loss.L2LossFunction()(pred.clone(), gtf.clone()))
    out['l2_masked'] = _np(ref_loss.L2LossFunction()(pred.clone(), gtf.clone(), mask.clone()))
    out['wahr'] = _np(ref_loss.WeightAdaptiveHeatmapRegressionLossFunction()(torch.sigmoid(pred), gtf.clone()))
    out['bce'] = _np(ref_loss.WeightedBceWithLogitsLossFunction()(pred.clone(), gt01.clone()))
    out['bce_masked'] = _np(ref_loss.WeightedBceWithLogitsLossFunction()(pred.clone(), gt01.clone(), mask.clone()))
    ce_pred = torch.randn(3, 4, 7, generator=g)
    ce_gt = torch.softmax(torch.randn(3, 4, 7, generator=g), dim=1)
    out.update(ce_pred=_np(ce_pred), ce_gt=_np(ce_gt))
    out['ce'] = _np(ref_loss.CrossEntropyWithLogitsLossFunction()(ce_pred, ce_gt))
    np.savez_compressed(os.path.join(OUT, 'primitive_losses.npz'), **out)


def neck_head_units():
    """Sub-module goldens on small, odd shapes (edge cases: non-divisible pooling bins, odd sizes)."""
    out = {}
    chans = (8, 16, 24, 32)
    g = synth._Gen(21)
    synth.neck_state_dict(g, '', 'upernext', chans, 16)
    neck = UperNextNeck(chans, 16)
    neck.load_state_dict(g.sd, strict=True)
    gen = torch.Generator().manual_seed(3)
    feats = [torch.randn(2, c, 40 >> i, 56 >> i, generator=gen) for i, c in enumerate(chans)]   # level 3 is 5x7
    for i, f in enumerate(feats):
        out[f'feat{i}'] = _np(f)
    out['upernext_neck'] = _np(neck(feats))
    g = synth._Gen(22)
    synth.neck_state_dict(g, '', 'fpn', chans, 16)
    neck = FpnNeck(chans, 16)
    neck.load_state_dict(g.sd, strict=True)
    out['fpn_neck'] = _np(neck(feats))
    x = torch.randn(2, 16, 9, 11, generator=gen)
    out['head_in'] = _np(x)
    for name, cls, kind in (('upernext_head', UperNextHead, 'upernext'), ('fpn_head', FpnHead, 'fpn')):
        for factor in (1, 2):
            g = synth._Gen(23 + factor)
            synth.head_state_dict(g, '', kind, 16, 3, out_bias=0.5)
            head = cls(16, 3, upsampling_factor=factor, init_output_bias=0.0)
            head.load_state_dict(g.sd, strict=True)
            out[f'{name}_x{factor}'] = _np(head(x))
    np.savez_compressed(os.path.join(OUT, 'neck_head_units.npz'), **out)


def inference_tensor_ops():
    """The tensor side of the reference's inference passes, run UNMODIFIED: ``pad_length_to_make_divisible`` /
    ``pad_mat_to_make_divisible`` (inferencing/opt.py:16-41), ``AdaptiveScalingInferencing.rough_infer`` (inferencing/
    adaptive_scaling.py:92-188) and ``.precise_infer`` (:295-396) with a stand-in ``model_jit`` that returns seeded network
    outputs (the network itself is pinned by the other fixtures), on images that need bottom / right / no padding."""
    from vkit.element import Image
    from vkit_open_model.inferencing import adaptive_scaling as ref_inf
    from vkit_open_model.inferencing.opt import pad_length_to_make_divisible, pad_mat_to_make_divisible

    out = {}
    lengths = [(6, 3), (7, 3), (1, 32), (32, 32), (33, 32), (720, 32), (721, 32), (2048, 32)]
    out['pad_length_cases'] = np.array(lengths, dtype=np.int64)
    out['pad_length_results'] = np.array([pad_length_to_make_divisible(a, b) for a, b in lengths], dtype=np.int64)

    class FakeJit:
        """forward_rough / forward_precise of a scripted model: seeded outputs at (H / 2, W / 2) of the padded input."""

        def __init__(self, seed):
            self.g = torch.Generator().manual_seed(seed)
            self.last = None

        def forward_rough(self, x):
            _, _, H, W = x.shape
            mask = torch.randn(1, 1, H // 2, W // 2, generator=self.g) * 3
            height = torch.rand(1, 1, H // 2, W // 2, generator=self.g) * 8
            self.last = (x.clone(), mask.clone(), height.clone())
            return mask, height

        def forward_precise(self, x):
            _, _, H, W = x.shape
            shape = (H // 2, W // 2)
            feats = (torch.randn(1, 1, *shape, generator=self.g) * 3, torch.randn(1, 2, *shape, generator=self.g) * 10,
                     torch.randn(1, 4, *shape, generator=self.g) * 2, torch.rand(1, 4, *shape, generator=self.g) * 30)
            self.last = (x.clone(),) + tuple(f.clone() for f in feats)
            return feats

    cfg = ref_inf.AdaptiveScalingInferencingConfig(model_jit='unused')
    rng = np.random.default_rng(133)
    sizes = [(64, 96), (70, 101), (33, 64), (96, 33), (1, 1)]
    out['sizes'] = np.array(sizes, dtype=np.int64)
    for idx, (H, W) in enumerate(sizes):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        out[f'image{idx}'] = img
        out[f'padded{idx}'] = pad_mat_to_make_divisible(img, 32)
        inf = object.__new__(ref_inf.AdaptiveScalingInferencing)       # __init__ only loads the TorchScript file
        inf.config = cfg
        inf.model_jit = FakeJit(1000 + idx)
        with torch.no_grad():
            r = inf.rough_infer(Image(mat=img))
        x, mask_f, height_f = inf.model_jit.last
        out[f'rough{idx}_input'] = _np(x)
        out[f'rough{idx}_mask_feature'] = _np(mask_f)
        out[f'rough{idx}_height_feature'] = _np(height_f)
        out[f'rough{idx}_mask'] = r.rough_char_mask.mat
        out[f'rough{idx}_height_map'] = r.rough_char_height_score_map.mat
        out[f'rough{idx}_resized_shape'] = np.array(r.resized_shape, dtype=np.int64)
        inf.model_jit = FakeJit(2000 + idx)
        with torch.no_grad():
            p = inf.precise_infer(Image(mat=img))
        feats = inf.model_jit.last
        for name, t in zip(('input', 'prob_feature', 'offset_feature', 'angle_feature', 'distance_feature'), feats):
            out[f'precise{idx}_{name}'] = _np(t)
        out[f'precise{idx}_prob_map'] = p.precise_char_prob_score_map.mat
        out[f'precise{idx}_offset'] = p.precise_np_char_up_left_corner_offset
        out[f'precise{idx}_angle'] = p.precise_np_char_corner_angle_distribution
        out[f'precise{idx}_distance'] = p.precise_np_char_corner_distance
    np.savez_compressed(os.path.join(OUT, 'inference_tensor_ops.npz'), **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == 'inference':
        inference_tensor_ops()
        print('inference fixture written to', OUT)
        sys.exit(0)
    inference_tensor_ops()
    primitive_losses()
    backbone_features()
    neck_head_units()
    full_model('upernext')
    full_model('fpn')
    print('golden fixtures written to', OUT)
