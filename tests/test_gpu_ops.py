"""GPU parity of every operator of the hot path against the oracle (unit level), in fp32 mode (rel 1e-4) and bf16 mode
(rel 2e-2): forward values, input gradients and every parameter gradient.  All calls go through the C ABI."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import GRAD_TOL, PER_TENSOR_GRAD_TOL, TOL, assert_close, compare_grads, oracle_params, randomize, rel_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


@pytest.fixture(scope='module')
def vk():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    assert torch.cuda.is_available()
    assert vk._lib.LIB.vkocr_device_check(torch.cuda.current_device()) == 0
    return vk


def _probe(shape, seed, device):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(device)


def _run_pair(vk, dtype, module, module_fn, oracle_fn, x_cpu, seed=0, input_grad=True):
    """Runs module_fn(module, x) on the GPU product path and oracle_fn(params, x) in fp32 torch on the GPU; compares."""
    dev = torch.device('cuda')
    module.to(dev)
    randomize(module, seed + 100)
    params = oracle_params(module)
    x = x_cpu.to(dev).requires_grad_(input_grad)
    xo = x_cpu.to(dev).double().requires_grad_(input_grad)
    with vk.precision(dtype):
        out = module_fn(module, x)
    ref = oracle_fn(params, xo)
    outs = out if isinstance(out, (tuple, list)) else [out]
    refs = ref if isinstance(ref, (tuple, list)) else [ref]
    total, total_ref = 0.0, 0.0
    for i, (o, r) in enumerate(zip(outs, refs)):
        assert_close(o.float(), r, TOL[dtype], f'output {i}')
        pr = _probe(tuple(r.shape), seed + i, dev)
        total = total + (o.float() * pr).sum()
        total_ref = total_ref + (r * pr.double()).sum()
    with vk.precision(dtype):
        total.backward()
    total_ref.backward()
    torch.cuda.synchronize()
    if input_grad:
        assert_close(x.grad.float(), xo.grad, PER_TENSOR_GRAD_TOL[dtype], 'input gradient')
    compare_grads(module, params, dtype, 'parameters')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('shape', [(2, 96, 20, 28), (1, 192, 9, 7), (3, 32, 5, 33)])
def test_convnext_layer(vk, dtype, shape):
    from oracle import model as om
    B, C, H, W = shape
    layer = vk.model.ConvNextBlockLayer(C)
    layer.eval()
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    _run_pair(vk, dtype, layer, lambda m, t: m(t), lambda p, t: om.convnext_layer(p, '', t), x)


@pytest.mark.parametrize('dtype', DTYPES)
def test_convnext_layer_stochastic_depth_mask(vk, dtype):
    """train mode: the drop mask is drawn with the reference's torch calls; with the same device seed the oracle, fed
    the identical mask, must agree (convnext.py:41-53)."""
    from oracle import model as om
    dev = torch.device('cuda')
    layer = vk.model.ConvNextBlockLayer(64, prob_bypass=0.5).to(dev)
    randomize(layer, 5)
    layer.train()
    x = torch.randn(6, 64, 8, 8, generator=torch.Generator().manual_seed(2)).to(dev)
    torch.manual_seed(77)
    xin = x.clone().requires_grad_(True)
    with vk.precision(dtype):
        y = layer(xin)
    torch.manual_seed(77)
    mask = torch.empty([6, 1, 1, 1], dtype=torch.float32, device=dev).bernoulli_(0.5).div_(0.5)
    assert 0 < int((mask == 0).sum()) < 6, 'seed should drop some but not all samples'
    params = oracle_params(layer)
    xo = x.double().requires_grad_(True)
    ref = om.convnext_layer(params, '', xo, mask.double())
    assert_close(y.float(), ref, TOL[dtype], 'stochastic depth output')
    dropped = (mask.reshape(-1) == 0).nonzero().reshape(-1)
    assert_close(y.float()[dropped], x[dropped], TOL[dtype], 'dropped samples are the identity')
    # backward through the masked branch: the mask rides in the GELU-derivative side channel and in the mask column of the
    # activation buffer (ops.ConvNextLayerFn), the oracle applies it the reference's way
    probe = _probe(tuple(ref.shape), 9, dev)
    with vk.precision(dtype):
        (y.float() * probe).sum().backward()
    (ref * probe.double()).sum().backward()
    torch.cuda.synchronize()
    assert_close(xin.grad.float(), xo.grad, PER_TENSOR_GRAD_TOL[dtype], 'input gradient under the mask')
    compare_grads(layer, params, dtype, 'parameters under the mask')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('hw', [(64, 96), (36, 44)])
def test_convnext_backbone(vk, dtype, hw):
    from oracle import model as om
    net = vk.model.ConvNext(3, ((32, 2), (64, 1), (96, 2), (128, 1)), False)
    net.eval()
    x = torch.randint(0, 256, (2, 3, *hw), generator=torch.Generator().manual_seed(3)).float()

    def scale_stem(module):
        with torch.no_grad():
            module.stem[0].weight.mul_(1.0 / 128)   # raw 0..255 pixels come in

    dev = torch.device('cuda')
    net.to(dev)
    randomize(net, 11)
    scale_stem(net)
    params = oracle_params(net)
    with vk.precision(dtype):
        feats = net(x.to(dev))
    refs = om.convnext_forward(params, x.to(dev).double(), prefix='')
    assert [tuple(f.shape) for f in feats] == [tuple(r.shape) for r in refs]
    total, total_ref = 0.0, 0.0
    for i, (f, r) in enumerate(zip(feats, refs)):
        assert_close(f.float(), r, TOL[dtype], f'feature {i}')
        pr = _probe(tuple(r.shape), 20 + i, dev)
        total = total + (f.float() * pr).sum()
        total_ref = total_ref + (r * pr.double()).sum()
    with vk.precision(dtype):
        total.backward()
    total_ref.backward()
    compare_grads(net, params, dtype, 'backbone')


def test_convnext_golden_features(vk, golden_dir):
    """fp32 mode against the fixture produced by the UNMODIFIED reference (oracle/make_golden.py: backbone_features)."""
    import os
    from oracle import synth
    g = np.load(os.path.join(golden_dir, 'convnext_tiny_features.npz'))
    gen = synth._Gen(7)
    synth.backbone_state_dict(gen, *synth.SIZES['tiny'], prefix='')
    net = vk.model.ConvNext.create_tiny()
    net.load_state_dict(gen.sd, strict=True)
    net.cuda().eval()
    x = synth.synth_image(1, 64, 96, seed=5).cuda()
    for dtype in DTYPES:
        with vk.precision(dtype), torch.no_grad():
            feats = net(x)
        for i, f in enumerate(feats):
            assert_close(f.float().cpu(), torch.from_numpy(g[f'f{i}']), TOL[dtype], f'{dtype} golden feature {i}')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('kind', ['upernext', 'fpn'])
def test_neck(vk, dtype, kind):
    from oracle import model as om
    chans = (32, 64, 96, 128)
    neck = vk.model.UperNextNeck(chans, 64) if kind == 'upernext' else vk.model.FpnNeck(chans, 64)
    gen = torch.Generator().manual_seed(4)
    feats_cpu = [torch.randn(2, c, 40 >> i, 56 >> i, generator=gen) for i, c in enumerate(chans)]   # level 3 is 5x7
    dev = torch.device('cuda')
    neck.to(dev)
    randomize(neck, 12)
    params = oracle_params(neck)
    feats = [f.to(dev).requires_grad_(True) for f in feats_cpu]
    feats_o = [f.to(dev).double().requires_grad_(True) for f in feats_cpu]
    with vk.precision(dtype):
        out = neck(feats)
    ref = om.neck_forward(params, '', feats_o, kind)
    assert_close(out.float(), ref, TOL[dtype], 'neck output')
    pr = _probe(tuple(ref.shape), 9, dev)
    with vk.precision(dtype):
        (out.float() * pr).sum().backward()
    (ref * pr.double()).sum().backward()
    for i in range(4):
        assert_close(feats[i].grad.float(), feats_o[i].grad, PER_TENSOR_GRAD_TOL[dtype], f'feature {i} gradient')
    compare_grads(neck, params, dtype, 'neck')


def test_neck_head_golden_units(vk, golden_dir):
    """fp32 mode against reference-generated unit fixtures on small odd shapes (non-divisible pooling bins, odd sizes)."""
    import os
    from oracle import synth
    g = np.load(os.path.join(golden_dir, 'neck_head_units.npz'))
    chans = (8, 16, 24, 32)
    feats = [torch.from_numpy(g[f'feat{i}']).cuda() for i in range(4)]
    with vk.precision(torch.float32), torch.no_grad():
        gen = synth._Gen(21)
        synth.neck_state_dict(gen, '', 'upernext', chans, 16)
        neck = vk.model.UperNextNeck(chans, 16)
        neck.load_state_dict(gen.sd, strict=True)
        assert_close(neck.cuda()(feats).float().cpu(), torch.from_numpy(g['upernext_neck']), 1e-4, 'upernext neck')
        gen = synth._Gen(22)
        synth.neck_state_dict(gen, '', 'fpn', chans, 16)
        neck = vk.model.FpnNeck(chans, 16)
        neck.load_state_dict(gen.sd, strict=True)
        assert_close(neck.cuda()(feats).float().cpu(), torch.from_numpy(g['fpn_neck']), 1e-4, 'fpn neck')
        x = torch.from_numpy(g['head_in']).cuda()
        for kind, cls in (('upernext', vk.model.UperNextHead), ('fpn', vk.model.FpnHead)):
            for factor in (1, 2):
                gen = synth._Gen(23 + factor)
                synth.head_state_dict(gen, '', kind, 16, 3, out_bias=0.5)
                head = cls(16, 3, upsampling_factor=factor)
                head.load_state_dict(gen.sd, strict=True)
                assert_close(head.cuda()(x).float().cpu(), torch.from_numpy(g[f'{kind}_head_x{factor}']), 1e-4,
                             f'{kind} head x{factor}')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('kind,factor,out_ch', [('upernext', 2, 1), ('upernext', 1, 4), ('fpn', 2, 2), ('fpn', 4, 3)])
def test_head(vk, dtype, kind, factor, out_ch):
    from oracle import model as om
    cls = vk.model.UperNextHead if kind == 'upernext' else vk.model.FpnHead
    head = cls(64, out_ch, upsampling_factor=factor, init_output_bias=0.3)
    x = torch.randn(2, 64, 9, 13, generator=torch.Generator().manual_seed(6))
    _run_pair(vk, dtype, head, lambda m, t: m(t), lambda p, t: om.head_forward(p, '', t, kind, factor), x)


@pytest.mark.parametrize('dtype', DTYPES)
def test_softplus_head_and_odd_inner(vk, dtype):
    """inner = (384 + out)//2 is 193 / 194 for the 2- and 4-channel heads: not a multiple of the vector width."""
    from oracle import model as om
    from vkit_ocr_model_adaptive_scaling_b200.model.adaptive_scaling import SoftplusHead
    head = SoftplusHead(vk.model.UperNextHead(384, 4, upsampling_factor=2))
    x = torch.randn(1, 384, 6, 10, generator=torch.Generator().manual_seed(8))
    _run_pair(vk, dtype, head, lambda m, t: m(t), lambda p, t: om.head_forward(p, '0.', t, 'upernext', 2, softplus=True), x)


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('kind,shape', [('upernext', (2, 64, 9, 13)), ('fpn', (2, 64, 9, 13)), ('upernext', (1, 384, 1, 1)),
                                        ('fpn', (1, 64, 1, 7)), ('upernext', (3, 64, 7, 1)), ('upernext', (1, 64, 5, 20))])
def test_head_x2_convolve_first_against_oracle(vk, dtype, kind, shape):
    """x2 heads run 'convolve first, resample after' (csrc/head_combine.cu): odd sizes, 1-pixel-wide maps (every clamp and
    padding case of the tiled kernels), tiles that straddle the image edge."""
    from oracle import model as om
    cls = vk.model.UperNextHead if kind == 'upernext' else vk.model.FpnHead
    head = cls(shape[1], 3, upsampling_factor=2, init_output_bias=-0.2)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(16))
    _run_pair(vk, dtype, head, lambda m, t: m(t), lambda p, t: om.head_forward(p, '', t, kind, 2), x)


@pytest.mark.parametrize('kind', ['upernext', 'fpn'])
def test_head_convolve_first_kernels_agree(vk, kind):
    """The shared-memory-tiled combine / adjoint kernels, the generic gather kernels and the legacy path (up-sample, then
    implicit-GEMM conv) are three statements of one linear map: in fp32 mode they must agree to rounding."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    dev = torch.device('cuda')
    cls = vk.model.UperNextHead if kind == 'upernext' else vk.model.FpnHead
    head = cls(48, 4, upsampling_factor=2, init_output_bias=0.1).to(dev)
    randomize(head, 31)
    x0 = torch.randn(2, 48, 11, 19, generator=torch.Generator().manual_seed(17)).to(dev)
    pr = _probe((2, 4, 22, 38), 5, dev)
    results = []
    try:
        for tapsplit, algo in ((True, 0), (True, 1), (False, 0)):
            ops.TAPSPLIT, ops.COMBINE_ALGO = tapsplit, algo
            head.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            with vk.precision(torch.float32):
                y = head(x)
                (y * pr).sum().backward()
            torch.cuda.synchronize()
            results.append((y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in head.named_parameters()}))
    finally:
        ops.TAPSPLIT, ops.COMBINE_ALGO = True, 0
    for other, what in ((results[1], 'generic kernels'), (results[2], 'legacy path')):
        assert_close(results[0][0], other[0], 2e-5, f'{what}: output')
        assert_close(results[0][1], other[1], 2e-5, f'{what}: input gradient')
        for n in results[0][2]:
            assert_close(results[0][2][n], other[2][n], 5e-5, f'{what}: grad {n}', atol=1e-6)


@pytest.mark.parametrize('shape,train', [((2, 9, 13), True), ((1, 16, 8), False), ((3, 5, 7), True), ((1, 1, 1), True)])
def test_head_combine_half_warp_tails(vk, shape, train, monkeypatch):
    """Heads with inner = 192 and one map take the half-warp tail of the combine kernel (bf16): same conv rows bit for bit and
    the same prediction maps up to the order of the warp reductions as the lane-per-vector tail."""
    import ctypes
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L, ops
    dev = torch.device('cuda')
    B, h, w = shape
    H, W, slot, nh = 2 * h, 2 * w, 200, 2
    ntot, nz = nh * slot, 9 * nh * slot
    g = torch.Generator().manual_seed(h * 100 + w)
    z = torch.randn(B * h * w, nz, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(ntot, generator=g).to(dev)
    par = [((torch.rand(192, generator=g) + 0.5).to(dev), (torch.randn(192, generator=g) * 0.1).to(dev),
            (torch.randn(1, 192, generator=g) * 0.1).to(dev), torch.randn(1, generator=g).to(dev)) for _ in range(nh)]
    # the pad channels of a slot: zero taps / zero bias, as the packer writes them
    zv = z.view(B * h * w, 9, nh, slot)
    zv[..., 192:] = 0
    bias.view(nh, slot)[:, 192:] = 0

    def run():
        outs = [torch.full((B, 1, H, W), -3.0, device=dev) for _ in range(nh)]
        conv = torch.full((B * H * W, ntot), 5.0, device=dev, dtype=torch.bfloat16)
        ht = L.HeadTail()
        ht.num_heads, ht.slot, ht.pixels_per_image = nh, slot, H * W
        for i in range(nh):
            ht.gamma[i], ht.beta[i], ht.w2[i], ht.b2[i] = (t.data_ptr() for t in par[i])
            ht.out[i] = outs[i].data_ptr()
            ht.inner[i], ht.out_channels[i], ht.softplus[i] = 192, 1, i
        L.check(L.LIB.vkocr_head_combine_fwd(1, L.ptr(z), nz, B, h, w, 2, 0, 3, ntot, L.ptr(bias), ctypes.byref(ht),
                                             L.ptr(conv) if train else None, ntot if train else 0, 0, L.stream_ptr()), 'head_combine_fwd')
        torch.cuda.synchronize()
        return outs, conv
    outs_h, conv_h = run()
    monkeypatch.setenv('VKOCR_HC_NOHALF', '1')
    outs_v, conv_v = run()
    if train:
        assert torch.equal(conv_h, conv_v), 'conv rows differ between the two tails'
    for a, b in zip(outs_h, outs_v):
        assert_close(a, b, 1e-5, 'prediction map: half-warp vs lane-per-vector tail', atol=1e-5)


@pytest.mark.parametrize('rows_shape,slot,softplus', [((3, 37, 41), 192, 0), ((2, 24, 31), 200, 1), ((1, 1, 1), 192, 1), ((1, 64, 64), 200, 0)])
def test_head_tail_backward_half_warp_kernel(vk, rows_shape, slot, softplus):
    """The inner = 192 / one-map backward (a half-warp per pixel row: every dense call of the training step) against the
    generic warp-per-row kernel (reached through the label-point entry with an identity pixel list) and against a torch fp32
    restatement of LN -> GELU -> 1x1 (-> Softplus) on the same bf16 conv rows.  Odd row counts, a padded slot, a slice inside
    a wider buffer."""
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L
    dev = torch.device('cuda')
    B, H, W = rows_shape
    M, inner, ld = B * H * W, 192, 2 * slot + 8
    g = torch.Generator().manual_seed(M + slot)
    buf = (torch.randn(M, ld, generator=g) * 1.5).to(dev).to(torch.bfloat16)
    conv = buf[:, slot:2 * slot]
    gamma = (torch.rand(inner, generator=g) + 0.5).to(dev)
    beta = (torch.randn(inner, generator=g) * 0.2).to(dev)
    w2 = (torch.randn(1, inner, generator=g) * 0.2).to(dev)
    b2 = torch.randn(1, generator=g).to(dev)
    dout = torch.randn(B, 1, H, W, generator=g).to(dev)
    # torch reference (fp32 math on the bf16-rounded rows)
    xr = conv[:, :inner].float().requires_grad_(True)
    gr, br, wr, b2r = (t.clone().requires_grad_(True) for t in (gamma, beta, w2, b2))
    pre = F.linear(F.gelu(F.layer_norm(xr, (inner,), gr, br, 1e-6)), wr, b2r)
    outr = F.softplus(pre) if softplus else pre
    (outr.view(B, H, W, 1).permute(0, 3, 1, 2) * dout).sum().backward()
    out = outr.detach().view(B, H, W, 1).permute(0, 3, 1, 2).contiguous()

    def run(points):
        dx = torch.full((M, ld), 7.0, device=dev, dtype=torch.bfloat16)
        gs = [torch.zeros(inner, device=dev) for _ in range(3)] + [torch.zeros(1, inner, device=dev), torch.zeros(1, device=dev)]
        dsl = dx[:, slot:2 * slot]
        if points:
            idx = torch.arange(M, device=dev, dtype=torch.int32)
            L.check(L.LIB.vkocr_head_tail_bwd_points(1, L.ptr(conv), ld, inner, slot, L.ptr(gamma), L.ptr(beta), L.ptr(w2), 1, softplus,
                                                     L.ptr(out), L.ptr(dout), H * W, L.ptr(idx), M, 0, L.ptr(dsl), ld, L.ptr(gs[0]),
                                                     L.ptr(gs[1]), L.ptr(gs[3]), L.ptr(gs[4]), L.ptr(gs[2]), L.stream_ptr()), 'bwd_points')
        else:
            L.check(L.LIB.vkocr_head_tail_bwd(1, L.ptr(conv), ld, inner, slot, L.ptr(gamma), L.ptr(beta), L.ptr(w2), 1, softplus, L.ptr(out),
                                              L.ptr(dout), H * W, M, L.ptr(dsl), ld, L.ptr(gs[0]), L.ptr(gs[1]), L.ptr(gs[3]), L.ptr(gs[4]),
                                              L.ptr(gs[2]), L.stream_ptr()), 'bwd')
        torch.cuda.synchronize()
        return dx, gs
    dx_h, gs_h = run(False)
    dx_g, gs_g = run(True)
    # the two kernels do the same fp32 arithmetic per row: identical bf16 rows up to the order of the warp reductions
    assert_close(dx_h[:, slot:slot + inner].float(), dx_g[:, slot:slot + inner].float(), 1e-2, 'dx: half-warp vs generic kernel')
    assert_close(dx_h[:, slot:slot + inner].float(), xr.grad, 1e-2, 'dx vs torch')
    assert torch.all(dx_h[:, slot + inner:2 * slot] == 0), 'pad columns of the slice must be written as zeros'
    assert torch.all(dx_h[:, :slot] == 7.0) and torch.all(dx_h[:, 2 * slot:] == 7.0), 'columns outside the slice must not be touched'
    for a, b, r, name in zip(gs_h, gs_g, (gr.grad, br.grad, xr.grad.sum(0), wr.grad, b2r.grad), ('dgamma', 'dbeta', 'dbias', 'dw2', 'db2')):
        assert_close(a, b, 1e-4, f'{name}: half-warp vs generic kernel', atol=1e-5)
        assert_close(a, r, 2e-3 if name != 'dbias' else 1e-2, f'{name} vs torch', atol=1e-4)


# ------------------------------------------------------------------------------------------------------- losses
def _loss_inputs(B, H, W, inset, P, seed, dev):
    from oracle import synth
    rb = synth.synth_rough_batch(B, 2 * H, 2 * W, seed=seed, inset=inset)
    pb = synth.synth_precise_batch(B, 2 * H, 2 * W, points=P, seed=seed, inset=inset)
    to = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    return to(rb), to(pb)


@pytest.mark.parametrize('B,H,W,inset', [(2, 48, 64, 4), (1, 33, 29, 0), (3, 40, 40, 10)])
def test_rough_loss(vk, B, H, W, inset):
    from oracle import loss as ol
    dev = torch.device('cuda')
    rb, _ = _loss_inputs(B, H, W, inset, 8, 50 + B, dev)
    g = torch.Generator().manual_seed(B)
    logit = (torch.randn(B, 1, H, W, generator=g) * 2).to(dev)
    height = (torch.rand(B, 1, H, W, generator=g) * 14).to(dev)   # some below the 1.1 floor
    a, b = logit.clone().requires_grad_(True), height.clone().requires_grad_(True)
    ao, bo = logit.clone().requires_grad_(True), height.clone().requires_grad_(True)
    fn = vk.loss_function.AdaptiveScalingRoughLossFunction(vk.loss_function.AdaptiveScalingRoughLossFunctionConifg())
    keys = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
    loss = fn(rough_char_mask_feature=a, rough_char_height_feature=b, **{k: rb[k] for k in keys})
    ref = ol.rough_loss(ao, bo, *(rb[k] for k in keys))
    assert loss.dim() == 0 and loss.is_cuda
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-6, (float(loss), float(ref))
    (loss / 2).backward()
    (ref / 2).backward()
    assert_close(a.grad, ao.grad, 1e-4, 'd rough / d mask logits')
    assert_close(b.grad, bo.grad, 1e-4, 'd rough / d height')


@pytest.mark.parametrize('B,H,W,inset,P', [(2, 48, 64, 4, 20), (1, 33, 29, 0, 7), (3, 40, 40, 10, 200)])
def test_precise_loss(vk, B, H, W, inset, P):
    from oracle import loss as ol
    dev = torch.device('cuda')
    _, pb = _loss_inputs(B, H, W, inset, P, 60 + B, dev)
    if P >= 2:   # duplicate label points must accumulate in the backward scatter
        pb['downsampled_label_point_y'][:, 1] = pb['downsampled_label_point_y'][:, 0]
        pb['downsampled_label_point_x'][:, 1] = pb['downsampled_label_point_x'][:, 0]
    g = torch.Generator().manual_seed(B)
    maps = [(torch.randn(B, c, H, W, generator=g) * s).to(dev) for c, s in ((1, 2.0), (2, 8.0), (4, 1.5), (4, 6.0))]
    maps[3] = maps[3].abs()
    ours = [m.clone().requires_grad_(True) for m in maps]
    refs = [m.clone().requires_grad_(True) for m in maps]
    fn = vk.loss_function.AdaptiveScalingPreciseLossFunction(vk.loss_function.AdaptiveScalingPreciseLossFunctionConifg())
    keys = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
            'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
            'char_corner_distances')
    loss = fn(precise_char_mask_feature=None, precise_char_prob_feature=ours[0],
              precise_char_up_left_corner_offset_feature=ours[1], precise_char_corner_angle_feature=ours[2],
              precise_char_corner_distance_feature=ours[3], **{k: pb[k] for k in keys})
    ref = ol.precise_loss(None, *refs, *(pb[k] for k in keys))
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-6, (float(loss), float(ref))
    (loss / 2).backward()
    (ref / 2).backward()
    for i, name in enumerate(('prob', 'offset', 'angle', 'distance')):
        assert_close(ours[i].grad, refs[i].grad, 1e-4, f'd precise / d {name}')


def test_precise_loss_guards(vk):
    """Label points outside the map (the reference's advanced indexing raises IndexError, :167-179) give a NaN loss and touch
    no memory outside the maps; negative indices wrap like torch; float offsets promote like F.smooth_l1_loss (no truncation);
    label tensors of another batch size are rejected; changed primitive hyper-parameters are refused, not ignored."""
    from oracle import loss as ol
    dev = torch.device('cuda')
    B, H, W, P = 2, 24, 32, 6
    _, pb = _loss_inputs(B, H, W, 3, P, 81, dev)
    g = torch.Generator().manual_seed(5)
    maps = [(torch.randn(B, c, H, W, generator=g)).to(dev) for c in (1, 2, 4, 4)]
    fn = vk.loss_function.AdaptiveScalingPreciseLossFunction(vk.loss_function.AdaptiveScalingPreciseLossFunctionConifg())
    keys = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
            'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
            'char_corner_distances')

    def run(batch, maps_in):
        ours = [m.clone().requires_grad_(True) for m in maps_in]
        loss = fn(None, *ours, **{k: batch[k] for k in keys})
        loss.backward()
        torch.cuda.synchronize()
        return loss, ours

    base, base_maps = run(pb, maps)
    # negative indices wrap (y - H addresses the same pixel)
    wrapped = dict(pb)
    wrapped['downsampled_label_point_y'] = pb['downsampled_label_point_y'] - H
    loss_w, maps_w = run(wrapped, maps)
    assert float(loss_w) == float(base)
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(maps_w, base_maps))
    # float offsets with fractional parts: the oracle (F.smooth_l1_loss) promotes, so must we
    frac = dict(pb)
    frac['char_up_left_offsets'] = pb['char_up_left_offsets'].float() + 0.4
    loss_f, _ = run(frac, maps)
    ref = ol.precise_loss(None, *maps, *(frac[k] for k in keys))
    assert abs(float(loss_f) - float(ref)) <= 1e-4 * abs(float(ref)), (float(loss_f), float(ref))
    assert abs(float(loss_f) - float(base)) > 1e-4
    # a point outside the map: NaN loss, finite gradients that ignore the point
    bad = dict(pb)
    y = pb['downsampled_label_point_y'].clone()
    y[1, 2] = H + 5
    bad['downsampled_label_point_y'] = y
    loss_b, maps_b = run(bad, maps)
    assert torch.isnan(loss_b)
    assert all(torch.isfinite(m.grad).all() for m in maps_b)
    # batch mismatch
    short = dict(pb)
    short['char_corner_angles'] = pb['char_corner_angles'][:1]
    with pytest.raises(RuntimeError):
        run(short, maps)
    # hyper-parameters of the primitives are compiled into the fused kernels
    fn.char_corner_distance_l1.smooth_beta = 1.0
    with pytest.raises(NotImplementedError):
        run(pb, maps)


def test_precise_loss_optional_terms(vk):
    """Off-by-default terms: masked focal on the optional char-mask head, prob smooth-L1, WAHR (reference :272-307)."""
    from oracle import loss as ol
    dev = torch.device('cuda')
    B, H, W, P = 2, 24, 32, 5
    _, pb = _loss_inputs(B, H, W, 3, P, 71, dev)
    g = torch.Generator().manual_seed(3)
    maps = [(torch.randn(B, c, H, W, generator=g)).to(dev) for c in (1, 1, 2, 4, 4)]
    ours = [m.clone().requires_grad_(True) for m in maps]
    refs = [m.clone().requires_grad_(True) for m in maps]
    cfg = vk.loss_function.AdaptiveScalingPreciseLossFunctionConifg(
        char_mask_focal_factor=1.5, char_prob_l1_factor=0.7, char_prob_wahr_factor=0.9)
    fn = vk.loss_function.AdaptiveScalingPreciseLossFunction(cfg)
    keys = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
            'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
            'char_corner_distances')
    loss = fn(ours[0], ours[1], ours[2], ours[3], ours[4], **{k: pb[k] for k in keys})
    ref = ol.precise_loss(refs[0], refs[1], refs[2], refs[3], refs[4], *(pb[k] for k in keys),
                          char_mask_focal_factor=1.5, char_prob_l1_factor=0.7, char_prob_wahr_factor=0.9)
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-6, (float(loss), float(ref))
    loss.backward()
    ref.backward()
    for i in range(5):
        assert_close(ours[i].grad, refs[i].grad, 2e-4, f'gradient of map {i}')


def test_primitive_losses_golden(vk, golden_dir):
    """Every primitive against the reference-generated fixture values (oracle/make_golden.py: primitive_losses)."""
    import os
    g = np.load(os.path.join(golden_dir, 'primitive_losses.npz'))
    lf = vk.loss_function
    dev = torch.device('cuda')
    pred, gt01, gtf, mask = (torch.from_numpy(g[k]).to(dev) for k in ('pred', 'gt01', 'gtf', 'mask'))
    cases = [
        ('focal', lf.FocalWithLogitsLossFunction()(pred, gt01)),
        ('focal_masked', lf.FocalWithLogitsLossFunction()(pred, gt01, mask)),
        ('dice', lf.DiceLossFunction()(torch.sigmoid(pred), gt01)),
        ('dice_masked', lf.DiceLossFunction()(torch.sigmoid(pred), gt01, mask)),
        ('l1', lf.L1LossFunction()(pred, gtf)),
        ('l1_masked', lf.L1LossFunction()(pred, gtf, mask)),
        ('smooth_l1', lf.L1LossFunction(smooth=True, smooth_beta=2.5)(pred, gtf)),
        ('smooth_l1_masked', lf.L1LossFunction(smooth=True)(pred, gtf, mask)),
        ('l2', lf.L2LossFunction()(pred, gtf)),
        ('l2_masked', lf.L2LossFunction()(pred, gtf, mask)),
        ('wahr', lf.WeightAdaptiveHeatmapRegressionLossFunction()(torch.sigmoid(pred), gtf)),
        ('bce', lf.WeightedBceWithLogitsLossFunction()(pred, gt01)),
        ('bce_masked', lf.WeightedBceWithLogitsLossFunction()(pred, gt01, mask)),
        ('ce', lf.CrossEntropyWithLogitsLossFunction()(torch.from_numpy(g['ce_pred']).to(dev), torch.from_numpy(g['ce_gt']).to(dev))),
    ]
    for name, value in cases:
        want = float(g[name])
        assert abs(float(value) - want) <= 2e-5 * abs(want) + 2e-6, (name, float(value), want)


@pytest.mark.parametrize('name', ['focal', 'dice', 'l1', 'smooth_l1', 'l2', 'wahr', 'bce', 'ce'])
@pytest.mark.parametrize('masked', [False, True])
def test_primitive_loss_gradients(vk, name, masked):
    from oracle import loss as ol
    lf = vk.loss_function
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(13)
    n = (5, 1000)
    pred = (torch.randn(n, generator=g) * 2).to(dev)
    gt01 = (torch.rand(n, generator=g) > 0.7).float().to(dev)
    gtf = torch.rand(n, generator=g).to(dev)
    mask = (torch.rand(n, generator=g) > 0.3).float().to(dev) if masked else None
    a = pred.clone().requires_grad_(True)
    b = pred.clone().requires_grad_(True)
    if name == 'ce':
        if masked:
            pytest.skip('cross entropy takes no mask')
        a = torch.randn(6, 4, 50, generator=g).to(dev).requires_grad_(True)
        b = a.detach().clone().requires_grad_(True)
        t = torch.softmax(torch.randn(6, 4, 50, generator=g), dim=1).to(dev)
        ours, ref = lf.CrossEntropyWithLogitsLossFunction()(a, t), ol.cross_entropy_with_logits(b, t)
    elif name == 'focal':
        ours, ref = lf.FocalWithLogitsLossFunction()(a, gt01, mask), ol.focal_with_logits(b, gt01, mask)
    elif name == 'dice':
        ours = lf.DiceLossFunction()(torch.sigmoid(a), gt01.clone(), mask)
        ref = ol.dice(torch.sigmoid(b), gt01, mask)
    elif name == 'l1':
        ours, ref = lf.L1LossFunction()(a, gtf, mask), ol.l1(b, gtf, mask)
    elif name == 'smooth_l1':
        ours, ref = lf.L1LossFunction(smooth=True, smooth_beta=0.8)(a, gtf, mask), ol.l1(b, gtf, mask, True, 0.8)
    elif name == 'l2':
        ours, ref = lf.L2LossFunction()(a, gtf, mask), ol.l2(b, gtf, mask)
    elif name == 'wahr':
        if masked:
            pytest.skip('WAHR takes no mask')
        ours = lf.WeightAdaptiveHeatmapRegressionLossFunction()(torch.sigmoid(a), gtf)
        ref = ol.weight_adaptive_heatmap_regression(torch.sigmoid(b), gtf)
    else:
        ours = lf.WeightedBceWithLogitsLossFunction()(a, gt01.clone(), mask)
        ref = ol.weighted_bce_with_logits(b, gt01.clone(), mask)
    assert abs(float(ours) - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-6, (float(ours), float(ref))
    ours.backward()
    ref.backward()
    assert_close(a.grad, b.grad, 1e-4, f'{name} gradient')


def test_errors_are_loud(vk):
    """No CPU fallback: CPU tensors are rejected; unsupported shapes come back as errors from the C ABI."""
    layer = vk.model.ConvNextBlockLayer(32)
    with pytest.raises(Exception):
        layer(torch.randn(1, 32, 8, 8))
    from vkit_ocr_model_adaptive_scaling_b200 import _lib as L
    rc = L.LIB.vkocr_dwconv7_fwd(1, None, 8, None, 8, 1, 4, 4, 8, None, None, None, 0, None)
    assert rc != 0 and b'null' in L.LIB.vkocr_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize('strip2', ['0', '1', '2'])
@pytest.mark.parametrize('shape', [(2, 32, 16, 40), (1, 64, 19, 80), (2, 32, 20, 20), (1, 96, 13, 35), (3, 32, 8, 120),
                                   (2, 64, 32, 48), (1, 32, 16, 16), (2, 32, 48, 32), (1, 32, 24, 80)])
def test_dwconv7_tiled_kernels_against_conv2d(vk, shape, strip2, monkeypatch):
    """The bf16 tiled depthwise kernels (helper.dconv7x7, model/helper.py:61-73) at every tile width (40 / 20 / 32 / 16), with
    partial tiles and the fused residual operand, 8 x 1 and 8 x 2 strips: forward and weight gradient against torch's own
    convolution of the same bf16-rounded operands in fp64 (only the final bf16 rounding / fp32 summation order may differ)."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    monkeypatch.setenv('VKOCR_DW_STRIP2', strip2)
    B, C, H, W = shape
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
    x.copy_(torch.randn(B, C, H, W, generator=g))
    add = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
    add.copy_(torch.randn(B, C, H, W, generator=g))
    y = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
    w = torch.randn(C, 1, 7, 7, generator=g).to(dev)
    bias = torch.randn(C, generator=g).to(dev)
    wt = w.reshape(C, 49).t().contiguous()                      # [49][C] tap table
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=3, groups=C)
    ops.dwconv7(x, y, wt, bias, None)
    assert_close(y.float(), ref, 4e-3, 'dwconv7 forward')
    ops.dwconv7(x, y, wt, None, add)
    ref2 = torch.nn.functional.conv2d(x.double(), w.double(), None, padding=3, groups=C) + add.double()
    assert_close(y.float(), ref2, 4e-3, 'dwconv7 + add')
    dw = torch.zeros(C, 1, 7, 7, device=dev)
    ops.dwconv7_wgrad(add, x, dw)
    xd = x.double().requires_grad_(False)
    wd = w.double().requires_grad_(True)
    (torch.nn.functional.conv2d(xd, wd, None, padding=3, groups=C) * add.double()).sum().backward()
    assert_close(dw, wd.grad, 1e-5, 'dwconv7 weight gradient')


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 32, 32, 384, 384, 3), (3, 20, 28, 384, 208, 3), (4, 16, 16, 1536, 384, 1), (1, 24, 40, 128, 832, 3)])
def test_gemm_cta_pair_mode_is_bit_identical(vk, shape, monkeypatch):
    """The CTA-pair schedule of the tcgen05 GEMM (cluster of 2, cta_group::2, each CTA stages half of the weight tile)
    accumulates every output element over the same K order as the one-CTA schedule: results must be bit-identical."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    B, H, W, C, N, ks = shape
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(N + H)
    x = ops.alloc_nhwc(B, H, W, C, torch.bfloat16, dev)
    x.copy_(torch.randn(B, C, H, W, generator=g))
    w = (torch.randn(N, ks * ks * C, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    outs = []
    for mode in ('0', '1'):
        monkeypatch.setenv('VKOCR_CTA2', mode)
        out = ops.alloc_nhwc(B, H, W, N, torch.bfloat16, dev)
        ops.gemm_nt(x, B, H, W, C, x.stride(3), ks, w, C, N, ops._epilogue(out, out.stride(3), bias=bias))
        torch.cuda.synchronize()
        outs.append(out.float().clone())
    assert torch.isfinite(outs[1]).all()
    assert torch.equal(outs[0], outs[1])
    # and both agree with the convolution itself
    wt = w.float().reshape(N, ks, ks, C).permute(0, 3, 1, 2).double()
    ref = torch.nn.functional.conv2d(x.double(), wt, bias.double(), padding=ks // 2)
    assert_close(outs[1], ref, 4e-3, 'pair-mode implicit-GEMM convolution')


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(640, 96), (300, 192), (1000, 384)])
def test_mlp_epilogue_gelu_with_derivative_side_channel(vk, shape):
    """ConvNeXt MLP up-projection epilogue act 3 (convnext.py:29-37): out = GELU(xW^T + b), second output = GELU'(xW^T + b)
    (fp16 bits in the 16-bit buffer); act 4 multiplies a data gradient by that side channel.  Checked against the erf
    formulas in fp64 on the same bf16 operands."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    M, C = shape
    hid = 4 * C
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(M)
    cp = (C + 63) // 64 * 64
    x = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
    w = torch.zeros(hid, cp)
    w[:, :C] = torch.randn(hid, C, generator=g) * (C ** -0.5)
    w = w.to(torch.bfloat16).to(dev)
    b = torch.randn(hid, generator=g).to(dev)
    out = torch.empty(M, hid, dtype=torch.bfloat16, device=dev)
    side = torch.empty(M, hid, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(x, 1, 1, M, C, C, 1, w, cp, hid, ops._epilogue(out, hid, out_pre=side, ld_pre=hid, bias=b, act=3))
    h = x.double() @ w[:, :C].double().t() + b.double()
    cdf = 0.5 * (1 + torch.erf(h / 2 ** 0.5))
    pdf = torch.exp(-h * h / 2) / (2 * np.pi) ** 0.5
    assert_close(out.float(), h * cdf, 4e-3, 'GELU output')
    dgelu = side.view(torch.float16).double()
    assert float((dgelu - (cdf + h * pdf)).abs().max()) <= 1.5e-3, 'GELU derivative side channel (fp16 bits)'
    # backward use: dH = (U W2) * side
    u = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
    w2 = torch.zeros(hid, cp)
    w2[:, :C] = torch.randn(hid, C, generator=g) * (C ** -0.5)
    w2 = w2.to(torch.bfloat16).to(dev)
    dh = torch.empty(M, hid, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(u, 1, 1, M, C, C, 1, w2, cp, hid, ops._epilogue(dh, hid, act=4, aux=side, ld_aux=hid))
    ref = (u.double() @ w2[:, :C].double().t()) * dgelu
    assert_close(dh.float(), ref, 4e-3, 'data gradient times the side channel')


@pytest.mark.gpu
@pytest.mark.parametrize('case', [(2, 8, 20, 20, 160, 160, 0), (2, 8, 1, 1, 20, 20, 0), (1, 16, 6, 6, 20, 20, 0), (2, 8, 3, 3, 20, 20, 0),
                                  (1, 8, 5, 7, 12, 17, 0), (2, 8, 10, 10, 20, 20, 0), (2, 8, 20, 20, 80, 80, 1), (1, 8, 7, 5, 29, 23, 1),
                                  (2, 16, 9, 13, 18, 26, 1), (1, 8, 1, 1, 2, 2, 1)])
def test_upsample_adjoint_against_interpolate(vk, case):
    """Up-sampling (F.interpolate bilinear align_corners=False / nearest: upernext.py:59-82,174-197, fpn.py:121-144) and its
    gather-form adjoint at every lane-split width of the kernel, against torch's own forward and autograd in fp32."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    B, C, h, w, H, W, mode = case
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(h * 100 + H)
    src = ops.alloc_nhwc(B, h, w, C, torch.float32, dev)
    src.copy_(torch.randn(B, C, h, w, generator=g))
    dst = ops.alloc_nhwc(B, H, W, C, torch.float32, dev)
    ops.upsample_fwd(src, dst, C, mode, False)
    xs = src.detach().clone().contiguous().requires_grad_(True)
    ref = F.interpolate(xs, size=(H, W), mode='bilinear' if mode == 0 else 'nearest', **({'align_corners': False} if mode == 0 else {}))
    assert_close(dst, ref.detach(), 1e-6, 'up-sampling forward')
    gd = ops.alloc_nhwc(B, H, W, C, torch.float32, dev)
    gd.copy_(torch.randn(B, C, H, W, generator=g))
    gs = ops.alloc_nhwc(B, h, w, C, torch.float32, dev, zero=True)
    ops.upsample_bwd(gd, gs, C, mode, False)
    ref.backward(gd.contiguous())
    assert_close(gs, xs.grad, 1e-5, 'up-sampling adjoint')
    base = torch.randn(B, C, h, w, generator=g).to(dev)
    gs2 = ops.alloc_nhwc(B, h, w, C, torch.float32, dev)
    gs2.copy_(base)
    ops.upsample_bwd(gd, gs2, C, mode, True)
    assert_close(gs2, xs.grad + base, 1e-5, 'accumulating adjoint')
    dst2 = ops.alloc_nhwc(B, H, W, C, torch.float32, dev)
    dst2.copy_(gd)
    ops.upsample_fwd(src, dst2, C, mode, True)
    assert_close(dst2, ref.detach() + gd, 1e-6, 'accumulating forward')


@pytest.mark.gpu
@pytest.mark.parametrize('case', [(2, 16, 20, 20, 1), (2, 16, 20, 20, 3), (1, 8, 64, 64, 6), (2, 8, 17, 23, 2), (1, 8, 5, 7, 3)])
def test_adaptive_avgpool_against_torch(vk, case):
    """nn.AdaptiveAvgPool2d(S) of the PPM block (upernext.py:59-65) on NHWC, one-pass and separable kernels, with gradient."""
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    B, C, H, W, S = case
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(H * 10 + S)
    x = ops.alloc_nhwc(B, H, W, C, torch.float32, dev)
    x.copy_(torch.randn(B, C, H, W, generator=g))
    xin = x.detach().requires_grad_(True)
    y = ops.AvgPoolFn.apply(xin, S)
    xr = x.detach().clone().contiguous().requires_grad_(True)
    ref = F.adaptive_avg_pool2d(xr, S)
    assert_close(y, ref.detach(), 1e-6, 'adaptive average pool')
    pr = torch.randn(B, C, S, S, generator=g).to(dev)
    (y * pr).sum().backward()
    (ref * pr).sum().backward()
    assert_close(xin.grad, xr.grad, 1e-6, 'adaptive average pool gradient')
