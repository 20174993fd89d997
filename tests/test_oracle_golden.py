"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as ol
from oracle import model as om
from oracle import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _close(a, b, rtol=2e-5, atol=2e-5):
    a = (a.detach() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))).to(torch.float64)
    b = (b.detach() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b))).to(torch.float64)
    assert a.shape == b.shape
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f'max abs err {err}'


def test_primitive_losses(golden_dir):
    g = _load(golden_dir, 'primitive_losses.npz')
    pred, gt01, gtf, mask = (torch.from_numpy(g[k]) for k in ('pred', 'gt01', 'gtf', 'mask'))
    _close(ol.focal_with_logits(pred, gt01), g['focal'])
    _close(ol.focal_with_logits(pred, gt01, mask), g['focal_masked'])
    _close(ol.dice(torch.sigmoid(pred), gt01), g['dice'])
    _close(ol.dice(torch.sigmoid(pred), gt01, mask), g['dice_masked'])
    _close(ol.l1(pred, gtf), g['l1'])
    _close(ol.l1(pred, gtf, mask), g['l1_masked'])
    _close(ol.l1(pred, gtf, None, True, 2.5), g['smooth_l1'])
    _close(ol.l1(pred, gtf, mask, True, 1.0), g['smooth_l1_masked'])
    _close(ol.l2(pred, gtf), g['l2'])
    _close(ol.l2(pred, gtf, mask), g['l2_masked'])
    _close(ol.weight_adaptive_heatmap_regression(torch.sigmoid(pred), gtf), g['wahr'])
    _close(ol.weighted_bce_with_logits(pred, gt01), g['bce'])
    _close(ol.weighted_bce_with_logits(pred, gt01, mask), g['bce_masked'])
    _close(ol.cross_entropy_with_logits(torch.from_numpy(g['ce_pred']), torch.from_numpy(g['ce_gt'])), g['ce'])


def test_convnext_features(golden_dir):
    g = _load(golden_dir, 'convnext_tiny_features.npz')
    gen = synth._Gen(7)
    synth.backbone_state_dict(gen, *synth.SIZES['tiny'], prefix='')
    feats = om.convnext_forward(gen.sd, synth.synth_image(1, 64, 96, seed=5), prefix='')
    for i, f in enumerate(feats):
        _close(f, g[f'f{i}'], rtol=1e-4, atol=1e-4)


def test_neck_head_units(golden_dir):
    g = _load(golden_dir, 'neck_head_units.npz')
    chans = (8, 16, 24, 32)
    feats = [torch.from_numpy(g[f'feat{i}']) for i in range(4)]
    gen = synth._Gen(21)
    synth.neck_state_dict(gen, '', 'upernext', chans, 16)
    _close(om.neck_forward(gen.sd, '', feats, 'upernext'), g['upernext_neck'], rtol=1e-4, atol=1e-4)
    gen = synth._Gen(22)
    synth.neck_state_dict(gen, '', 'fpn', chans, 16)
    _close(om.neck_forward(gen.sd, '', feats, 'fpn'), g['fpn_neck'], rtol=1e-4, atol=1e-4)
    x = torch.from_numpy(g['head_in'])
    for kind in ('upernext', 'fpn'):
        for factor in (1, 2):
            gen = synth._Gen(23 + factor)
            synth.head_state_dict(gen, '', kind, 16, 3, out_bias=0.5)
            _close(om.head_forward(gen.sd, '', x, kind, factor), g[f'{kind}_head_x{factor}'], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_full_model_forward_loss_backward(golden_dir, neck):
    g = _load(golden_dir, f'adaptive_scaling_tiny_{neck}.npz')
    batch, height, width, points, inset = (int(v) for v in g['meta'])
    sd = synth.synth_state_dict('tiny', neck, seed=133)
    names = [str(n) for n in g['param_names']]
    assert names == list(sd.keys())
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rb = synth.synth_rough_batch(batch, height, width, seed=133, inset=inset)
    pb = synth.synth_precise_batch(batch, height, width, points=points, seed=133, inset=inset)

    mask, hgt = om.forward_rough(params, rb['image'])
    _close(mask, g['rough_mask'], rtol=2e-4, atol=2e-4)
    _close(hgt, g['rough_height'], rtol=2e-4, atol=2e-4)
    rl = ol.rough_loss(mask, hgt, rb['downsampled_mask'], rb['downsampled_score_map'], rb['downsampled_shape'],
                       rb['downsampled_core_box'])
    _close(rl, g['rough_loss'], rtol=1e-4, atol=1e-5)
    grads = torch.autograd.grad(rl / 2, list(params.values()), allow_unused=True)
    norms = np.array([float(x.norm()) if x is not None else -1.0 for x in grads])
    sums = np.array([float(x.double().sum()) if x is not None else 0.0 for x in grads])
    np.testing.assert_allclose(norms, g['rough_grad_norm'], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(sums, g['rough_grad_sum'], rtol=5e-3, atol=2e-4)

    prob, off, ang, dist = om.forward_precise(params, pb['image'])
    _close(prob, g['precise_prob'], rtol=2e-4, atol=2e-4)
    _close(off, g['precise_offset'], rtol=2e-4, atol=2e-4)
    _close(ang, g['precise_angle'], rtol=2e-4, atol=2e-4)
    _close(dist, g['precise_distance'], rtol=2e-4, atol=2e-4)
    pl = ol.precise_loss(None, prob, off, ang, dist, pb['downsampled_char_prob_score_map'], pb['downsampled_char_mask'],
                         pb['downsampled_shape'], pb['downsampled_core_box'], pb['downsampled_label_point_y'],
                         pb['downsampled_label_point_x'], pb['char_up_left_offsets'], pb['char_corner_angles'],
                         pb['char_corner_distances'])
    _close(pl, g['precise_loss'], rtol=1e-4, atol=1e-5)
    grads = torch.autograd.grad(pl / 2, list(params.values()), allow_unused=True)
    norms = np.array([float(x.norm()) if x is not None else -1.0 for x in grads])
    np.testing.assert_allclose(norms, g['precise_grad_norm'], rtol=2e-3, atol=1e-6)
    by_name = dict(zip(names, grads))
    for key in g.files:
        if key.startswith('precise_grad::'):
            _close(by_name[key.split('::', 1)[1]], g[key], rtol=2e-3, atol=1e-5)


def test_inference_tensor_ops_against_reference_fixture(golden_dir):
    """oracle/infer.py against the outputs of the UNMODIFIED reference inference code (inferencing/opt.py:16-41,
    inferencing/adaptive_scaling.py:92-188,295-396; fixture written by oracle/make_golden.py) and the reference's own test
    vectors (tests/test_evaluation.py:15-22): bit-exact."""
    from oracle import infer as oi
    g = _load(golden_dir, 'inference_tensor_ops.npz')
    assert oi.pad_length_to_make_divisible(6, 3) == (6, 0)          # tests/test_evaluation.py:16-18
    assert oi.pad_length_to_make_divisible(7, 3) == (9, 2)          # tests/test_evaluation.py:20-22
    for (length, factor), want in zip(g['pad_length_cases'], g['pad_length_results']):
        assert oi.pad_length_to_make_divisible(int(length), int(factor)) == (int(want[0]), int(want[1]))
    for idx, (H, W) in enumerate(g['sizes']):
        H, W = int(H), int(W)
        img = g[f'image{idx}']
        padded = oi.pad_mat_to_make_divisible(img, 32)
        assert padded.dtype == np.uint8 and np.array_equal(padded, g[f'padded{idx}'])
        x = oi.network_input(img, 32)
        assert x.dtype == torch.float32 and np.array_equal(x.numpy(), g[f'rough{idx}_input'])
        assert np.array_equal(x.numpy(), g[f'precise{idx}_input'])
        Hp, Wp = padded.shape[:2]
        mask, hmap, shape = oi.rough_postprocess(torch.from_numpy(g[f'rough{idx}_mask_feature'])[0, 0],
                                                 torch.from_numpy(g[f'rough{idx}_height_feature'])[0, 0], H, W, Hp, Wp)
        assert mask.dtype == np.uint8 and np.array_equal(mask, g[f'rough{idx}_mask'])
        assert hmap.dtype == np.float32 and np.array_equal(hmap, g[f'rough{idx}_height_map'])
        assert tuple(shape) == tuple(int(v) for v in g[f'rough{idx}_resized_shape'])
        feats = [torch.from_numpy(g[f'precise{idx}_{n}_feature']) for n in ('prob', 'offset', 'angle', 'distance')]
        prob, off, ang, dist = oi.precise_postprocess(*feats, H, W, Hp, Wp)
        for got, name in ((prob, 'prob_map'), (off, 'offset'), (ang, 'angle'), (dist, 'distance')):
            want = g[f'precise{idx}_{name}']
            assert got.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want), (idx, name)
