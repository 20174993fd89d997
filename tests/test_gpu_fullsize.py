"""Parity at the BASELINE.json shapes: the two-pass training step at B=32 / 640x640 (configs #3/#4) and
forward_rough at B=8 / 2048x2048 (config #5), product (bf16 mode) against the oracle in fp32 on the same GPU (TF32 off, see
conftest.py).  These are the shapes bench.py times; they index buffers beyond 2^31 elements (Z of the precise head group:
819 200 x 7 488, its conv output 3 276 800 x 832), which no small-shape test reaches.

The oracle is evaluated in batch chunks so that its fp32 autograd tape fits beside the product's buffers: outputs first
(no_grad), the losses and d(loss)/d(outputs) on the full batch (the loss normalisers are global), then per chunk a forward
with grad and a backward from that chunk's output gradients -- exactly the full-batch gradient."""
import pytest
import torch

from _util import TOL, assert_close, compare_grads

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


def _chunked_oracle_pass(forward, loss_of_outputs, params, image, chunk):
    """Outputs (detached, full batch), loss value, and parameter gradients accumulated into params[k].grad."""
    with torch.no_grad():
        outs = [forward(params, image[i:i + chunk]) for i in range(0, image.shape[0], chunk)]
    full = [torch.cat([o[j] for o in outs]).requires_grad_(True) for j in range(len(outs[0]))]
    del outs
    loss = loss_of_outputs(*full)
    loss.backward()
    for i in range(0, image.shape[0], chunk):
        part = forward(params, image[i:i + chunk])
        torch.autograd.backward(list(part), [f.grad[i:i + chunk] for f in full])
        del part
    return [f.detach() for f in full], loss.detach()


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_training_step_at_the_benchmark_shape(vk, neck):
    from oracle import loss as ol
    from oracle import model as om
    from oracle import synth
    from vkit_ocr_model_adaptive_scaling_b200.training import train_step
    dev = torch.device('cuda')
    B, S, P = 32, 640, 200
    M, LF = vk.model, vk.loss_function
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck)))
    model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=133), strict=True)
    model.to(dev).eval()
    rb = _to(synth.synth_rough_batch(B, S, S, seed=133, inset=10), dev)
    pb = _to(synth.synth_precise_batch(B, S, S, points=P, seed=133, inset=10), dev)
    rough_fn = LF.AdaptiveScalingRoughLossFunction(LF.AdaptiveScalingRoughLossFunctionConifg())
    precise_fn = LF.AdaptiveScalingPreciseLossFunction(LF.AdaptiveScalingPreciseLossFunctionConifg())
    dtype = torch.bfloat16
    with vk.precision(dtype):
        rl, pl = train_step(model, rough_fn, precise_fn, rb, pb)
        with torch.no_grad():
            ours_rough = [t.float() for t in model.forward_rough(rb['image'])]
            ours_precise = [t.float() for t in model.forward_precise(pb['image'])]
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}   # fp32 oracle
    rough_ref, rl_ref = _chunked_oracle_pass(
        om.forward_rough, lambda m, h: ol.rough_loss(m, h, *(rb[k] for k in ROUGH_KEYS)) / 2, params, rb['image'], 4)
    precise_ref, pl_ref = _chunked_oracle_pass(
        om.forward_precise, lambda *o: ol.precise_loss(None, *o, *(pb[k] for k in PRECISE_KEYS)) / 2, params, pb['image'], 4)
    tol = TOL[dtype]
    for got, ref, name in zip(ours_rough + ours_precise, rough_ref + precise_ref,
                              ('rough mask', 'rough height', 'precise prob', 'precise offset', 'precise angle', 'precise distance')):
        assert_close(got, ref, tol, f'{neck} B=32/640: {name}')
    assert abs(float(rl) / 2 - float(rl_ref)) <= tol * abs(float(rl_ref)), (float(rl) / 2, float(rl_ref))
    assert abs(float(pl) / 2 - float(pl_ref)) <= tol * abs(float(pl_ref)), (float(pl) / 2, float(pl_ref))
    err = compare_grads(model, params, dtype, f'{neck} step at B=32/640x640')
    print(f'[{neck} B=32/640x640] losses {float(rl):.5f} / {float(pl):.5f} (oracle {2 * float(rl_ref):.5f} / {2 * float(pl_ref):.5f}), '
          f'global gradient rel L2 error {err:.3e}')


def test_rough_inference_at_the_benchmark_shape(vk):
    """Config #5: forward_rough at B=8 / 2048x2048 (eval, no_grad) against the fp32 oracle, one page at a time."""
    from oracle import model as om
    from oracle import synth
    dev = torch.device('cuda')
    M = vk.model
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT))
    model.load_state_dict(synth.synth_state_dict('tiny', 'upernext', seed=133), strict=True)
    model.to(dev).eval()
    image = synth.synth_image(8, 2048, 2048, seed=5).to(dev)
    with vk.precision(torch.bfloat16), torch.no_grad():
        mask, height = (t.float() for t in model.forward_rough(image))
    torch.cuda.synchronize()
    params = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for i in range(8):
            ref_mask, ref_height = om.forward_rough(params, image[i:i + 1])
            assert_close(mask[i:i + 1], ref_mask, TOL[torch.bfloat16], f'2048x2048 page {i}: mask logits')
            assert_close(height[i:i + 1], ref_height, TOL[torch.bfloat16], f'2048x2048 page {i}: char height')
