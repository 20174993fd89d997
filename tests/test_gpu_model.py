"""GPU parity of the whole adaptive-scaling path (model + losses + backward) against (i) the fixtures written by the
UNMODIFIED reference (tests/golden, oracle/make_golden.py) and (ii) the oracle run on the same device in fp32."""
import os

import numpy as np
import pytest
import torch

from _util import (GRAD_TOL, NEGLIGIBLE, PER_TENSOR_GRAD_TOL, SMALL_SHAPE_BF16_GRAD_TOL, TOL, assert_close, compare_grads,
                   oracle_params, rel_err)

pytestmark = pytest.mark.gpu

ROUGH_KEYS = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
PRECISE_KEYS = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
                'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
                'char_corner_distances')


@pytest.fixture(scope='module')
def vk():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    return vk


def _to(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _build(vk, neck, size='tiny'):
    M = vk.model
    cfg = M.AdaptiveScalingConfig(
        size=M.AdaptiveScalingSize(size),
        neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT if neck == 'upernext' else M.AdaptiveScalingNeckHeadType.FPN)
    return M.AdaptiveScaling(cfg)


def _check_norms(names, norms, want, tol, what):
    """Per-parameter gradient norms against the reference's (tensors below NEGLIGIBLE of the global norm: absolute)."""
    total = float(np.sqrt(np.sum(np.square(want[want > 0]))))
    for n, a, b in zip(names, norms, want):
        if b <= 0:
            assert a <= 0, f'{n}: unexpected {what} gradient'
        elif b < NEGLIGIBLE * total:
            assert abs(max(a, 0.0) - b) <= tol * NEGLIGIBLE * total, f'{what} grad norm of {n}: {a} vs {b}'
        else:
            assert abs(a - b) <= tol * b, f'{what} grad norm of {n}: {a} vs {b}'


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_full_model_against_reference_golden(vk, golden_dir, neck, dtype):
    from oracle import synth
    g = np.load(os.path.join(golden_dir, f'adaptive_scaling_tiny_{neck}.npz'))
    batch, height, width, points, inset = (int(v) for v in g['meta'])
    dev = torch.device('cuda')
    model = _build(vk, neck)
    sd = synth.synth_state_dict('tiny', neck, seed=133)
    assert [str(n) for n in g['param_names']] == list(model.state_dict().keys())
    model.load_state_dict(sd, strict=True)
    model.to(dev).eval()
    rb = _to(synth.synth_rough_batch(batch, height, width, seed=133, inset=inset), dev)
    pb = _to(synth.synth_precise_batch(batch, height, width, points=points, seed=133, inset=inset), dev)
    lf = vk.loss_function
    rough_fn = lf.AdaptiveScalingRoughLossFunction(lf.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = lf.AdaptiveScalingPreciseLossFunction(lf.AdaptiveScalingPreciseLossFunctionConifg())
    tol, gtol = TOL[dtype], PER_TENSOR_GRAD_TOL[dtype]
    names = [n for n, _ in model.named_parameters()]

    with vk.precision(dtype):
        mask, hgt = model.forward_rough(rb['image'])
        assert mask.dtype == torch.float32 and tuple(mask.shape) == (batch, 1, height // 2, width // 2)
        assert_close(mask.cpu(), torch.from_numpy(g['rough_mask']), tol, 'rough mask')
        assert_close(hgt.cpu(), torch.from_numpy(g['rough_height']), tol, 'rough height')
        rl = rough_fn(rough_char_mask_feature=mask, rough_char_height_feature=hgt, **{k: rb[k] for k in ROUGH_KEYS})
        assert abs(float(rl) - float(g['rough_loss'])) <= tol * abs(float(g['rough_loss'])), (float(rl), float(g['rough_loss']))
        (rl / 2).backward()
    norms = np.array([float(p.grad.norm()) if p.grad is not None and float(p.grad.abs().max()) > 0 else -1.0
                      for _, p in model.named_parameters()])
    want = g['rough_grad_norm']
    _check_norms(names, norms, want, gtol, 'rough')
    model.zero_grad(set_to_none=True)

    with vk.precision(dtype):
        prob, off, ang, dist = model.forward_precise(pb['image'])
        for name, t in (('precise_prob', prob), ('precise_offset', off), ('precise_angle', ang), ('precise_distance', dist)):
            assert_close(t.cpu(), torch.from_numpy(g[name]), tol, name)
        pl = precise_fn(precise_char_mask_feature=None, precise_char_prob_feature=prob,
                        precise_char_up_left_corner_offset_feature=off, precise_char_corner_angle_feature=ang,
                        precise_char_corner_distance_feature=dist, **{k: pb[k] for k in PRECISE_KEYS})
        assert abs(float(pl) - float(g['precise_loss'])) <= tol * abs(float(g['precise_loss'])), (float(pl), float(g['precise_loss']))
        (pl / 2).backward()
    norms = np.array([float(p.grad.norm()) if p.grad is not None and float(p.grad.abs().max()) > 0 else -1.0
                      for _, p in model.named_parameters()])
    want = g['precise_grad_norm']
    _check_norms(names, norms, want, gtol, 'precise')
    by_name = dict(model.named_parameters())
    for key in g.files:
        if key.startswith('precise_grad::'):
            assert_close(by_name[key.split('::', 1)[1]].grad.cpu(), torch.from_numpy(g[key]), gtol, key)


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_training_step_against_oracle(vk, neck, dtype):
    """Two-pass step (train.py:397-478) at a size with ragged tiles (160x224 -> 80x112 maps): every parameter gradient
    after both backward passes against the oracle's autograd on the same device."""
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.training import train_step
    dev = torch.device('cuda')
    B, H, W, P = 2, 160, 224, 16
    model = _build(vk, neck)
    model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=7), strict=True)
    model.to(dev).eval()
    params = oracle_params(model)
    rb = _to(synth.synth_rough_batch(B, H, W, seed=3, inset=6), dev)
    pb = _to(synth.synth_precise_batch(B, H, W, points=P, seed=3, inset=6), dev)
    lf = vk.loss_function
    rough_fn = lf.AdaptiveScalingRoughLossFunction(lf.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = lf.AdaptiveScalingPreciseLossFunction(lf.AdaptiveScalingPreciseLossFunctionConifg())
    with vk.precision(dtype):
        rl, pl = train_step(model, rough_fn, precise_fn, rb, pb)
    rb64 = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in rb.items()}
    pb64 = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in pb.items()}
    rb, pb = rb64, pb64
    mask, hgt = om.forward_rough(params, rb['image'])
    rl_ref = ol.rough_loss(mask, hgt, *(rb[k] for k in ROUGH_KEYS))
    (rl_ref / 2).backward()
    prob, off, ang, dist = om.forward_precise(params, pb['image'])
    pl_ref = ol.precise_loss(None, prob, off, ang, dist, *(pb[k] for k in PRECISE_KEYS))
    (pl_ref / 2).backward()
    tol = TOL[dtype]
    assert abs(float(rl) - float(rl_ref)) <= tol * abs(float(rl_ref)), (float(rl), float(rl_ref))
    assert abs(float(pl) - float(pl_ref)) <= tol * abs(float(pl_ref)), (float(pl), float(pl_ref))
    autocast_err = None
    if dtype == torch.bfloat16:
        # For the record (not a relaxation): the reference's own algorithm (the oracle's torch ops) under stock
        # torch.autocast(bfloat16) on this device against the same fp64 result.  Through 18 backbone layers, the neck and
        # the heads the gradient that reaches the stem (87 % of the gradient norm: the images are raw 0..255) carries every
        # bf16 rounding of the chain.
        p32 = {k: v.detach().float().requires_grad_(True) for k, v in params.items()}
        f32 = lambda d: {k: (v.float() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in d.items()}
        rb32, pb32 = f32(rb), f32(pb)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            m, h = om.forward_rough(p32, rb32['image'])
        (ol.rough_loss(m.float(), h.float(), *(rb32[k] for k in ROUGH_KEYS)) / 2).backward()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            outs = om.forward_precise(p32, pb32['image'])
        (ol.precise_loss(None, *(o.float() for o in outs), *(pb32[k] for k in PRECISE_KEYS)) / 2).backward()
        num = sum(float((p32[k].grad.double() - params[k].grad).square().sum()) for k in params if params[k].grad is not None)
        den = sum(float(params[k].grad.square().sum()) for k in params if params[k].grad is not None)
        autocast_err = (num / den) ** 0.5
        print(f'[{neck} step] reference algorithm under torch.autocast(bfloat16): global gradient rel L2 error {autocast_err:.3e}')
    explicit = SMALL_SHAPE_BF16_GRAD_TOL.get(f'tiny/{neck}') if dtype == torch.bfloat16 else None
    compare_grads(model, params, dtype, f'{neck} step', grad_tol=explicit)


def test_train_mode_stochastic_depth_matches_reference_rng(vk):
    """train(): the drop masks are drawn with the reference's torch calls in layer order, so the oracle fed the masks
    re-drawn from the same device seed reproduces the forward pass (convnext.py:41-53)."""
    from oracle import model as om
    from oracle import synth
    dev = torch.device('cuda')
    model = _build(vk, 'upernext')
    model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=9), strict=True)
    model.to(dev).train()
    x = synth.synth_image(4, 64, 64, seed=2).to(dev)
    torch.manual_seed(1234)
    with vk.precision(torch.float32), torch.no_grad():
        mask, hgt = model.forward_rough(x)
    torch.manual_seed(1234)
    probs = om.stochastic_depth_probs((3, 3, 9, 3))
    masks = {}
    for s, row in enumerate(probs):
        for l, p in enumerate(row):
            if p == 0.0:
                continue
            m = torch.empty([4, 1, 1, 1], dtype=torch.float32, device=dev).bernoulli_(1.0 - p).div_(1.0 - p)
            masks[(s, l)] = m
    ref_mask, ref_hgt = om.forward_rough(oracle_params(model), x.double(), drop_masks={k: v.double() for k, v in masks.items()})
    assert_close(mask, ref_mask, 1e-4, 'train-mode rough mask')
    assert_close(hgt, ref_hgt, 1e-4, 'train-mode rough height')


def test_state_dict_round_trip_and_shapes(vk):
    """Reference shape facts (tests/test_convnext.py:46-50, tests/test_upernext.py:28, tests/test_fpn.py:28,39,47,
    corrected tests/test_adaptive_scaling.py:51-62 with the 4-channel distance head)."""
    dev = torch.device('cuda')
    M = vk.model
    with torch.no_grad():
        feats = M.ConvNext.create_tiny().to(dev)(torch.rand(1, 3, 320, 320, device=dev) * 255)
        assert [tuple(f.shape) for f in feats] == [(1, 96, 80, 80), (1, 192, 40, 40), (1, 384, 20, 20), (1, 768, 10, 10)]
        feats = M.ConvNext.create_tiny(stem_use_pconv2x2=True).to(dev)(torch.rand(1, 3, 320, 320, device=dev) * 255)
        assert [tuple(f.shape) for f in feats] == [(1, 96, 160, 160), (1, 192, 80, 80), (1, 384, 40, 40), (1, 768, 20, 20)]
        fs = [torch.rand(1, c, 80 >> i, 80 >> i, device=dev) for i, c in enumerate((96, 192, 384, 768))]
        assert tuple(M.UperNextNeck((96, 192, 384, 768), 384).to(dev)(fs).shape) == (1, 384, 80, 80)
        assert tuple(M.FpnNeck((96, 192, 384, 768), 400).to(dev)(fs).shape) == (1, 400, 80, 80)
        x = torch.rand(1, 400, 80, 80, device=dev)
        assert tuple(M.FpnHead(400, 1, upsampling_factor=1).to(dev)(x).shape) == (1, 1, 80, 80)
        assert tuple(M.FpnHead(400, 1, upsampling_factor=2).to(dev)(x).shape) == (1, 1, 160, 160)
        model = _build(vk, 'upernext').to(dev)
        img = torch.rand(1, 3, 320, 320, device=dev) * 255
        assert [tuple(t.shape) for t in model.forward_rough(img)] == [(1, 1, 160, 160)] * 2
        assert [tuple(t.shape) for t in model.forward_precise(img)] == [(1, 1, 160, 160), (1, 2, 160, 160), (1, 4, 160, 160),
                                                                      (1, 4, 160, 160)]
        with pytest.raises(NotImplementedError):
            model(img)   # the reference defines no forward() either
    sd = model.state_dict()
    clone = _build(vk, 'upernext')
    clone.load_state_dict(sd, strict=True)
    assert all(v.dtype == torch.float32 for v in sd.values())


@pytest.mark.parametrize('size,neck', [('base', 'fpn'), ('small', 'upernext'), ('large', 'upernext')])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_larger_configs_against_oracle(vk, size, neck, dtype):
    """The other `create_*` sizes of the reference (convnext.py:188-225): BASE has 128..1024 channels, so the heads'
    inner width (256 + out)/2 no longer fits one 256-column tile and the head group takes the unfused tail; SMALL has a
    27-layer stage.  Forward outputs, both losses and every gradient against the oracle on two 128x192 images, held to the
    north-star tolerances (measured bf16 gradient error: SMALL 1.3e-2, BASE 1.3e-2, LARGE 1.0e-2; on a single 64x96 image the
    same configurations measure 3.2e-2 - 3.5e-2 / 1.9e-2 / 1.9e-2: the coherent part of a gradient grows with the pixel count,
    the rounding noise of the chained bf16 tensors with its square root)."""
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.training import train_step
    dev = torch.device('cuda')
    B, H, W, P = 2, 128, 192, 16
    model = _build(vk, neck, size)
    model.load_state_dict(synth.synth_state_dict(size, neck, seed=11), strict=True)
    model.to(dev).eval()
    params = oracle_params(model)
    rb = _to(synth.synth_rough_batch(B, H, W, seed=5, inset=2), dev)
    pb = _to(synth.synth_precise_batch(B, H, W, points=P, seed=5, inset=2), dev)
    lf = vk.loss_function
    rough_fn = lf.AdaptiveScalingRoughLossFunction(lf.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = lf.AdaptiveScalingPreciseLossFunction(lf.AdaptiveScalingPreciseLossFunctionConifg())
    with vk.precision(dtype):
        rl, pl = train_step(model, rough_fn, precise_fn, rb, pb)
    f64 = lambda d: {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in d.items()}
    rb, pb = f64(rb), f64(pb)
    rl_ref = ol.rough_loss(*om.forward_rough(params, rb['image']), *(rb[k] for k in ROUGH_KEYS))
    (rl_ref / 2).backward()
    pl_ref = ol.precise_loss(None, *om.forward_precise(params, pb['image']), *(pb[k] for k in PRECISE_KEYS))
    (pl_ref / 2).backward()
    tol = TOL[dtype]
    assert abs(float(rl) - float(rl_ref)) <= tol * abs(float(rl_ref)), (float(rl), float(rl_ref))
    assert abs(float(pl) - float(pl_ref)) <= tol * abs(float(pl_ref)), (float(pl), float(pl_ref))
    compare_grads(model, params, dtype, f'{size}/{neck} step')


def test_frozen_parameters_get_no_gradient(vk):
    """requires_grad_(False) on the backbone (fine-tuning the necks / heads): the backward kernels leave the frozen
    parameters without a .grad, and every other gradient equals the unfrozen run's."""
    from oracle import synth
    dev = torch.device('cuda')
    M, LF = vk.model, vk.loss_function
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT))
    model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=7), strict=True)
    model.to(dev).eval()
    rb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.synth_rough_batch(1, 64, 96, seed=3, inset=4).items()}
    fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
    rk = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')

    def run():
        model.zero_grad(set_to_none=True)
        with vk.precision(torch.float32):
            m, h = model.forward_rough(rb['image'])
            fn(rough_char_mask_feature=m, rough_char_height_feature=h, **{k: rb[k] for k in rk}).backward()
        torch.cuda.synchronize()
        return {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}

    full = run()
    for p in model.backbone.parameters():
        p.requires_grad_(False)
    frozen = run()
    for n, g in frozen.items():
        if n.startswith('backbone.'):
            assert g is None, f'{n}: a frozen parameter received a gradient'
        elif n.startswith('rough_'):
            assert g is not None, n
            assert_close(g, full[n], 1e-5, f'{n} (fp32 atomics reorder sums run to run)', atol=1e-7)


def test_rough_loss_with_hard_negative_bce_term(vk):
    """`bce_factor > 0` (off by default, loss_function/adaptive_scaling.py:29-30,88-95): the composite rough loss with the
    hard-negative BCE term through the model's own outputs, value and gradients against the oracle (fp32 mode)."""
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    dev = torch.device('cuda')
    model = _build(vk, 'fpn')
    model.load_state_dict(synth.synth_state_dict('tiny', 'fpn', seed=13), strict=True)
    model.to(dev).eval()
    params = oracle_params(model)
    rb = _to(synth.synth_rough_batch(2, 64, 96, seed=8, inset=3), dev)
    lf = vk.loss_function
    fn = lf.AdaptiveScalingRoughLossFunction(lf.AdaptiveScalingRoughLossFunctionConifg(bce_factor=0.7))
    with vk.precision(torch.float32):
        mask, hgt = model.forward_rough(rb['image'])
        loss = fn(rough_char_mask_feature=mask, rough_char_height_feature=hgt, **{k: rb[k] for k in ROUGH_KEYS})
        loss.backward()
    rb64 = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in rb.items()}
    ref = ol.rough_loss(*om.forward_rough(params, rb64['image']), *(rb64[k] for k in ROUGH_KEYS), bce_factor=0.7)
    ref.backward()
    plain = ol.rough_loss(*om.forward_rough({k: v.detach() for k, v in params.items()}, rb64['image']), *(rb64[k] for k in ROUGH_KEYS))
    assert abs(float(ref) - float(plain)) > 1e-3                       # the term is really on
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref)), (float(loss), float(ref))
    compare_grads(model, params, torch.float32, 'rough step with the hard-negative BCE term')


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_label_point_backward_equals_dense_backward(vk, neck):
    """The offset / angle / distance heads receive gradient only at the label points; their backward runs on those pixels
    alone (csrc/head_sparse.cu).  It must reproduce the dense backward: duplicate points, points on the map border and in
    the corners, negative (wrapped) indices, fp32 mode."""
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    dev = torch.device('cuda')
    B, H, W, P = 2, 64, 96, 12
    model = _build(vk, neck)
    model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=17), strict=True)
    model.to(dev).eval()
    pb = _to(synth.synth_precise_batch(B, H, W, points=P, seed=4, inset=2), dev)
    y, x = pb['downsampled_label_point_y'].clone(), pb['downsampled_label_point_x'].clone()
    h2, w2 = H // 2, W // 2
    y[0, 0], x[0, 0] = 0, 0                      # corners and borders
    y[0, 1], x[0, 1] = h2 - 1, w2 - 1
    y[0, 2], x[0, 2] = 0, w2 - 1
    y[1, 0], x[1, 0] = h2 - 1, 5
    y[1, 1], x[1, 1] = y[1, 2], x[1, 2]          # a duplicate
    y[1, 3], x[1, 3] = y[1, 4] - h2, x[1, 4] - w2    # the same pixel again, through wrapped negative indices
    pb['downsampled_label_point_y'], pb['downsampled_label_point_x'] = y, x
    fn = vk.loss_function.AdaptiveScalingPreciseLossFunction(vk.loss_function.AdaptiveScalingPreciseLossFunctionConifg())
    image = pb['image']
    results = []
    try:
        for sparse in (True, False):
            ops.SPARSE_BACKWARD = sparse
            model.zero_grad(set_to_none=True)
            launches0 = vk._lib.LIB.vkocr_launch_count()
            with vk.precision(torch.float32):
                outs = model.forward_precise(image)
                fn(None, *outs, **{k: pb[k] for k in PRECISE_KEYS}).backward()
            torch.cuda.synchronize()
            results.append(({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                            vk._lib.LIB.vkocr_launch_count() - launches0))
    finally:
        ops.SPARSE_BACKWARD = True
    (gs, _), (gd, _) = results
    assert set(gs) == set(gd)       # the backbone gradients carry the data gradient of the heads
    for n in gd:
        assert_close(gs[n], gd[n], 5e-5, f'grad {n}', atol=1e-7)


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_label_point_forward_gives_the_dense_loss_and_gradients(vk, neck):
    """`forward_precise(x, label_points=...)` (an extension, off by default) evaluates the offset / angle / distance heads at the
    label pixels only.  The maps agree with the dense call AT the label points and are zero elsewhere; the precise loss and
    every parameter gradient agree with the dense call (fp32 mode)."""
    from oracle import synth
    dev = torch.device('cuda')
    B, H, W, P = 2, 64, 96, 10
    model = _build(vk, neck)
    model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=19), strict=True)
    model.to(dev).eval()
    pb = _to(synth.synth_precise_batch(B, H, W, points=P, seed=6, inset=2), dev)
    y, x = pb['downsampled_label_point_y'].clone(), pb['downsampled_label_point_x'].clone()
    y[0, 0], x[0, 0] = 0, 0
    y[0, 1], x[0, 1] = H // 2 - 1, W // 2 - 1
    y[1, 1], x[1, 1] = y[1, 2], x[1, 2]
    pb['downsampled_label_point_y'], pb['downsampled_label_point_x'] = y, x
    fn = vk.loss_function.AdaptiveScalingPreciseLossFunction(vk.loss_function.AdaptiveScalingPreciseLossFunctionConifg())
    results = []
    for lp in (None, (y, x)):
        model.zero_grad(set_to_none=True)
        with vk.precision(torch.float32):
            outs = model.forward_precise(pb['image']) if lp is None else model.forward_precise(pb['image'], label_points=lp)
            loss = fn(None, *outs, **{k: pb[k] for k in PRECISE_KEYS})
            loss.backward()
        torch.cuda.synchronize()
        results.append((float(loss), [o.detach().clone() for o in outs], {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (ld, od, gd), (ls, os_, gs) = results
    assert abs(ls - ld) <= 1e-5 * abs(ld), (ls, ld)
    assert_close(os_[0], od[0], 1e-6, 'prob map (dense in both)')
    bi = torch.arange(B, device=dev)[:, None]
    for k in (1, 2, 3):
        assert_close(os_[k][bi, :, y, x], od[k][bi, :, y, x], 2e-5, f'head {k} at the label points')
        mask = torch.ones_like(os_[k], dtype=torch.bool)
        mask[bi, :, y, x] = False
        assert float(os_[k][mask].abs().max()) == 0.0
    assert set(gs) == set(gd)
    for n in gd:
        assert_close(gs[n], gd[n], 5e-5, f'grad {n}', atol=1e-7)


def test_graphed_train_step_replays_the_eager_step(vk):
    """training.GraphedTrainStep: the two-pass step captured into one CUDA graph.  In eval mode (no stochastic depth) a
    replay gives the eager step's losses and gradients (up to the order of the fp32 atomics); new batches copied into the
    graph's input buffers give the eager result for THOSE batches; in train mode successive replays draw fresh
    stochastic-depth masks (torch's graph-safe Philox offsets)."""
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.parallel import DataParallel
    from vkit_ocr_model_adaptive_scaling_b200.training import GraphedTrainStep, train_step
    dev = torch.device('cuda')
    B, H, W, P = 2, 96, 128, 10
    model = _build(vk, 'upernext')
    model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=23), strict=True)
    model.to(dev).eval()
    LF = vk.loss_function
    rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
    batches = [(_to(synth.synth_rough_batch(B, H, W, seed=s, inset=4), dev), _to(synth.synth_precise_batch(B, H, W, points=P, seed=s, inset=4), dev))
               for s in (1, 2)]
    dp = DataParallel(model)
    try:
        with vk.precision(torch.float32):      # exact fp32 kernels: run-to-run noise is the atomics' order only
            eager = []
            for rb, pb in batches:
                losses = train_step(model, rough_fn, precise_fn, rb, pb, dp)
                eager.append(([float(x) for x in losses], [f.clone() for f in dp.buckets.flat]))
            step = GraphedTrainStep(model, rough_fn, precise_fn, *batches[0], dp, warmup=1)
            for (rb, pb), (want_losses, want_grads) in list(zip(batches, eager)) + [(batches[0], eager[0])]:
                for f in dp.buckets.flat:
                    f.fill_(float('nan'))            # the replay zeroes and refills the buckets itself
                losses = [float(x) for x in step(rb, pb)]
                assert losses == pytest.approx(want_losses, rel=1e-5), (losses, want_losses)
                for name, got, want in zip(dp.buckets.names, dp.buckets.flat, want_grads):
                    assert_close(got, want, 1e-4, f'bucket {name}')
            with pytest.raises(ValueError):
                step({**batches[0][0], 'downsampled_shape': (1, 1)}, batches[0][1])
            step.close()
            with pytest.raises(RuntimeError):
                step(*batches[0])
            model.train()
            torch.manual_seed(5)
            step = GraphedTrainStep(model, rough_fn, precise_fn, *batches[0], dp, warmup=1)
            a = [float(x) for x in step(*batches[0])]
            b = [float(x) for x in step(*batches[0])]
            assert all(np.isfinite(a + b)) and a != b, (a, b)     # other drop-path masks -> other losses
            step.close()
    finally:
        dp.close()


def test_fpn_step_meets_the_plain_bound_at_a_third_of_a_megapixel(vk):
    """The one stated tolerance exception (tests/_util.py: SMALL_SHAPE_BF16_GRAD_TOL, TINY/FPN on two 160x224 images:
    2.9e-2) is a small-sample effect: the same configuration, weights and loss on four 256x320 images (4.6x the pixels) is
    held to the plain north-star 2e-2 here, as it is at the benchmark shape (tests/test_gpu_fullsize.py: 8.5e-3)."""
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.training import train_step
    dev = torch.device('cuda')
    B, H, W, P = 4, 256, 320, 16
    model = _build(vk, 'fpn')
    model.load_state_dict(synth.synth_state_dict('tiny', 'fpn', seed=7), strict=True)
    model.to(dev).eval()
    params = oracle_params(model)
    rb = _to(synth.synth_rough_batch(B, H, W, seed=3, inset=6), dev)
    pb = _to(synth.synth_precise_batch(B, H, W, points=P, seed=3, inset=6), dev)
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
    tol = TOL[torch.bfloat16]
    assert abs(float(rl) - float(rl_ref)) <= tol * abs(float(rl_ref)), (float(rl), float(rl_ref))
    assert abs(float(pl) - float(pl_ref)) <= tol * abs(float(pl_ref)), (float(pl), float(pl_ref))
    # the optimizer's view: all parameter gradients as one vector (measured 1.75e-2; per-tensor bounds are the business of
    # test_training_step_against_oracle and the full-size test)
    num = den = 0.0
    for name, p in model.named_parameters():
        ref = params[name].grad
        if ref is None:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        num += float((p.grad.double() - ref).square().sum())
        den += float(ref.square().sum())
    err = (num / den) ** 0.5
    print(f'[fpn step, 4 x 256x320] global gradient rel L2 error {err:.3e}')
    assert err <= GRAD_TOL[torch.bfloat16], f'global gradient relative L2 error {err:.3e} > {GRAD_TOL[torch.bfloat16]:.1e}'
