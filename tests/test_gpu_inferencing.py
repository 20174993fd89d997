"""Tensor-side rough inference ops on the device against the oracle restatement of the reference lines
(inferencing/opt.py:16-41, inferencing/adaptive_scaling.py:109-180)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vk():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    return vk


def test_inference_tensor_ops_against_reference_fixture(vk, golden_dir):
    """The device kernels against the outputs of the UNMODIFIED reference inference code (tests/golden/
    inference_tensor_ops.npz, written by oracle/make_golden.py from inferencing/opt.py:16-41 and inferencing/
    adaptive_scaling.py:92-188,295-396), fed the same network outputs: uint8 ingest and the rough mask / height map
    bit-exact, the precise post-ops to fp32 rounding of sigmoid / softmax."""
    import math
    import os
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L
    g = np.load(os.path.join(golden_dir, 'inference_tensor_ops.npz'))
    for (length, factor), want in zip(g['pad_length_cases'], g['pad_length_results']):
        assert vk.inferencing.pad_length_to_make_divisible(int(length), int(factor)) == (int(want[0]), int(want[1]))
    dev = torch.device('cuda')
    for idx, (H, W) in enumerate(g['sizes']):
        H, W = int(H), int(W)
        x = vk.inferencing.ingest_images(torch.from_numpy(g[f'image{idx}']).to(dev), 32)
        assert torch.equal(x.cpu(), torch.from_numpy(g[f'rough{idx}_input']))
        logit, height = (torch.from_numpy(g[f'rough{idx}_{n}_feature']).to(dev) for n in ('mask', 'height'))
        _, _, h, w = logit.shape
        mask = torch.empty((1, h, w), dtype=torch.uint8, device=dev)
        hmap = torch.empty((1, h, w), dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_rough_postprocess(L.ptr(logit), L.ptr(height), 1, h, w, math.ceil(H / 2), math.ceil(W / 2), 0.5, 3.0,
                                              L.ptr(mask), L.ptr(hmap), L.stream_ptr()), 'rough_postprocess')
        want_mask = g[f'rough{idx}_mask']
        sure = np.abs(g[f'rough{idx}_mask_feature'][0, 0]) > 1e-6          # sigmoid(x) >= 0.5 <=> x >= 0, up to rounding at x ~ 0
        assert np.array_equal(mask[0].cpu().numpy()[sure], want_mask[sure])
        assert np.array_equal(hmap[0].cpu().numpy(), g[f'rough{idx}_height_map'])
        feats = [torch.from_numpy(g[f'precise{idx}_{n}_feature']).to(dev) for n in ('prob', 'offset', 'angle', 'distance')]
        prob = torch.empty((1, h, w), dtype=torch.float32, device=dev)
        offs = torch.empty((1, h, w, 2), dtype=torch.float32, device=dev)
        angs = torch.empty((1, h, w, 4), dtype=torch.float32, device=dev)
        dists = torch.empty((1, h, w, 4), dtype=torch.float32, device=dev)
        L.check(L.LIB.vkocr_precise_postprocess(*(L.ptr(f) for f in feats), 1, h, w, 4, math.ceil(H / 2), math.ceil(W / 2), L.ptr(prob),
                                                L.ptr(offs), L.ptr(angs), L.ptr(dists), L.stream_ptr()), 'precise_postprocess')
        assert np.array_equal(offs[0].cpu().numpy(), g[f'precise{idx}_offset'])
        assert np.array_equal(dists[0].cpu().numpy(), g[f'precise{idx}_distance'])
        assert np.allclose(prob[0].cpu().numpy(), g[f'precise{idx}_prob_map'], rtol=1e-6, atol=1e-7)
        assert np.allclose(angs[0].cpu().numpy(), g[f'precise{idx}_angle'], rtol=1e-5, atol=1e-7)
        zero = g[f'precise{idx}_prob_map'] == 0.0
        assert np.array_equal(prob[0].cpu().numpy() == 0.0, zero)                # the padding is forced to exactly 0


@pytest.mark.parametrize('hw', [(64, 96), (70, 101), (33, 32), (1, 1)])
def test_ingest_matches_reference_padding(vk, hw):
    from oracle import infer as oi
    H, W = hw
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    want = oi.network_input(img, 32)
    got = vk.inferencing.ingest_images(torch.from_numpy(img).cuda(), 32)
    assert tuple(got.shape) == tuple(want.shape)
    assert torch.equal(got.cpu(), want)                       # bit-exact: integers 0..255 and zeros


def test_rough_infer_tensors_against_oracle(vk):
    from oracle import infer as oi
    from oracle import model as om
    from oracle import synth
    dev = torch.device('cuda')
    M = vk.model
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY,
                                                      neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT))
    sd = synth.synth_state_dict('tiny', 'upernext', seed=21, rough_height_bias=3.0)   # heights straddle the 3.0 cut
    model.load_state_dict(sd, strict=True)
    model.to(dev).eval()
    H, W = 90, 141                                            # pads to 96 x 160
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    with vk.precision(torch.float32):
        mask, hmap, shape = vk.inferencing.rough_infer_tensors(model, torch.from_numpy(img).cuda())
    x = oi.network_input(img, 32).double().to(dev)
    params = {k: v.detach().double().to(dev) for k, v in sd.items()}
    with torch.no_grad():
        logit, height = om.forward_rough(params, x)
    want_mask, want_h, want_shape = oi.rough_postprocess(logit[0, 0].float().cpu(), height[0, 0].float().cpu(), H, W,
                                                         x.shape[2], x.shape[3])
    assert shape == want_shape
    assert tuple(mask.shape) == (1,) + want_mask.shape and mask.dtype == torch.uint8
    got_mask, got_h = mask[0].cpu().numpy(), hmap[0].cpu().numpy()
    # outside a 1e-4 band around the two thresholds (fp32 kernels vs the fp64 oracle) the maps must agree exactly
    lo = logit[0, 0].float().cpu().numpy()
    hh = height[0, 0].float().cpu().numpy()
    sure_m = np.abs(lo) > 1e-4
    sure_h = np.abs(hh - 3.0) > 1e-3
    assert np.array_equal(got_mask[sure_m], want_mask[sure_m])
    zero_w, zero_g = want_h == 0.0, got_h == 0.0
    assert np.array_equal(zero_w[sure_h], zero_g[sure_h])
    assert (zero_w.sum() > 0) and (~zero_w).sum() > 0
    keep = sure_h & ~zero_w
    assert np.allclose(got_h[keep], want_h[keep], rtol=1e-4, atol=1e-4)
    assert got_mask[shape[0]:].sum() == 0 and got_mask[:, shape[1]:].sum() == 0     # padding forced negative


def test_rough_postprocess_kernel_on_synthetic_maps(vk):
    """The post-op kernel alone on random logits / heights (both mask classes, heights on both sides of the cut)."""
    import ctypes
    from oracle import infer as oi
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L
    g = torch.Generator().manual_seed(4)
    h, w, H, W = 48, 80, 90, 141
    logit = torch.randn(1, 1, h, w, generator=g) * 3
    height = torch.rand(1, 1, h, w, generator=g) * 8
    mask = torch.empty((1, h, w), dtype=torch.uint8, device='cuda')
    hmap = torch.empty((1, h, w), dtype=torch.float32, device='cuda')
    ld, hd = logit.cuda(), height.cuda()
    L.check(L.LIB.vkocr_rough_postprocess(L.ptr(ld), L.ptr(hd), 1, h, w, 45, 71, 0.5, 3.0, L.ptr(mask), L.ptr(hmap), L.stream_ptr()),
            'rough_postprocess')
    want_mask, want_h, shape = oi.rough_postprocess(logit[0, 0], height[0, 0], H, W, 96, 160)
    assert shape == (45, 71)
    assert np.array_equal(mask[0].cpu().numpy(), want_mask) and 0.2 < want_mask.mean() < 0.8
    assert np.array_equal(hmap[0].cpu().numpy(), want_h) and 0.1 < (want_h == 0).mean() < 0.9


def test_precise_infer_tensors_against_oracle(vk):
    """uint8 image -> forward_precise -> device post-ops, against the oracle restatement of inferencing/adaptive_scaling.py:
    322-386 applied to the same network outputs (the network itself is covered by tests/test_gpu_model.py)."""
    from oracle import infer as oi
    from oracle import synth
    M = vk.model
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT))
    model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=5), strict=True)
    model.cuda().eval()
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(75, 118, 3), dtype=np.uint8)       # pads to 96 x 128
    with vk.precision(torch.float32):
        prob, offs, angs, dists = vk.inferencing.precise_infer_tensors(model, torch.from_numpy(img).cuda())
        x = vk.inferencing.ingest_images(torch.from_numpy(img).cuda(), 32)
        with torch.no_grad():
            feats = [f.cpu() for f in model.forward_precise(x)]
    want = oi.precise_postprocess(*feats, 75, 118, 96, 128)
    for got, ref, what in zip((prob, offs, angs, dists), want, ('prob', 'offset', 'angle', 'distance')):
        got = got[0].cpu().numpy()
        assert got.shape == ref.shape, (what, got.shape, ref.shape)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6), what
    assert np.abs(angs[0].cpu().numpy().sum(-1) - 1).max() < 1e-5
    assert prob[0, 38:].abs().sum() == 0 and prob[0, :, 59:].abs().sum() == 0          # ceil(75 / 2), ceil(118 / 2)


@pytest.mark.parametrize('size', [5, 3, 4])
def test_peak_mask_against_scipy_maximum_filter(vk, size):
    """Peak picking (inferencing/adaptive_scaling.py:477-491) bit-exactly against scipy's maximum_filter, with plateaus (ties),
    a char mask, and maps smaller than the window."""
    from oracle import infer as oi
    rng = np.random.default_rng(size)
    for (h, w) in ((37, 53), (4, 3), (1, 9)):
        score = rng.random((2, h, w)).astype(np.float32)
        score[0] = np.round(score[0] * 8) / 8                           # plateaus: ties must count as peaks on both sides
        cmask = (rng.random((2, h, w)) > 0.3).astype(np.uint8)
        cfg = vk.inferencing.PreciseInferConfig(precise_build_polygons_maximum_filter_size=size)
        for use_mask in (False, True):
            got = vk.inferencing.find_peaks(torch.from_numpy(score).cuda(), torch.from_numpy(cmask).cuda() if use_mask else None, cfg)
            for b in range(2):
                want = oi.peak_mask(score[b], cmask[b] if use_mask else None, size=size, positive_thr=0.7)
                assert np.array_equal(got[b].cpu().numpy(), want), (h, w, b, use_mask)
        assert got.sum() > 0
