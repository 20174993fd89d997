"""Checkpoint / export interop with the reference's train script (experiment/adaptive_scaling/train.py:91-96,307-338,586-644):
host-side logic, runs on CPU (the modules are only constructed and (de)serialised, no kernel is launched)."""
import os
import sys

import pytest
import torch


@pytest.fixture(scope='module')
def vk():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    return vk


def _model(vk, neck='upernext'):
    M = vk.model
    cfg = M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck))
    return M.AdaptiveScaling(cfg), cfg


def test_restore_state_round_trip(vk, tmp_path):
    """A `.pt` in the reference's layout (cattrs.unstructure(RestoreState), train.py:597-603) loads into our modules, and
    what we save has the same four keys and a `model_jit_state_dict` with the reference's keys / shapes / order."""
    from oracle import synth
    sd = synth.synth_state_dict('tiny', 'upernext', seed=3)
    ref_opt_state = {'state': {}, 'param_groups': [{'lr': 8e-4, 'initial_lr': 8e-4, 'params': list(range(len(sd)))}]}
    path = tmp_path / 'state_dict_7.pt'
    torch.save({'epoch_idx': 7, 'model_jit_state_dict': sd, 'optimizer_state_dict': ref_opt_state,
                'optimizer_scheduler_state_dict': {'base_lrs': [8e-4], 'eta_min': 8e-6, 'last_epoch': 7}}, path)
    state = vk.checkpoint.load_restore_state(path)
    assert state['epoch_idx'] == 7 and state['optimizer_scheduler_state_dict']['eta_min'] == 8e-6
    model, _ = _model(vk)
    result = vk.checkpoint.load_model_state(model, state)
    assert not result.missing_keys and not result.unexpected_keys
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    out = tmp_path / 'state_dict_8_not_best.pt'
    vk.checkpoint.save_restore_state(out, 8, model, ref_opt_state, state['optimizer_scheduler_state_dict'])
    again = torch.load(out, map_location='cpu', weights_only=False)
    assert tuple(again.keys()) == vk.checkpoint.RESTORE_STATE_KEYS and again['epoch_idx'] == 8
    assert list(again['model_jit_state_dict'].keys()) == list(sd.keys())
    assert all(torch.equal(again['model_jit_state_dict'][k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        torch.save({'something': 1}, tmp_path / 'bad.pt')
        vk.checkpoint.load_restore_state(tmp_path / 'bad.pt')


REFERENCE = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'vkit_open_model')), reason='reference checkout not present (GPU box)')
@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_export_into_reference_modules(vk, neck, tmp_path):
    """Our weights load (strict) into the reference's eager module and into its torch.jit.script-ed form, and the dumped
    TorchScript file (train.py:635-644) reloads and evaluates like the oracle on the same weights."""
    sys.path.insert(0, REFERENCE)
    try:
        from vkit_open_model import model as ref_model
        from oracle import model as om
        from oracle import synth
        model, cfg = _model(vk, neck)
        model.load_state_dict(synth.synth_state_dict('tiny', neck, seed=9), strict=True)
        ref_cfg = ref_model.AdaptiveScalingConfig(size=ref_model.AdaptiveScalingSize.TINY,
                                                  neck_head_type=ref_model.AdaptiveScalingNeckHeadType(neck))
        ref_eager = vk.checkpoint.export_to_reference_module(model, ref_model.AdaptiveScaling(ref_cfg))
        for (ka, va), (kb, vb) in zip(ref_eager.state_dict().items(), model.state_dict().items()):
            assert ka == kb and torch.equal(va, vb)
        path = tmp_path / 'model_jit.pt'
        vk.checkpoint.build_reference_model_jit(model, cfg, output_model_jit=str(path))
        jit = torch.jit.load(str(path), map_location='cpu')
        x = torch.randint(0, 256, (1, 3, 64, 96), generator=torch.Generator().manual_seed(1)).float()
        with torch.no_grad():
            mask, height = jit.forward_rough(x)
            want = om.forward_rough({k: v for k, v in model.state_dict().items()}, x)
        assert torch.allclose(mask, want[0], rtol=1e-4, atol=1e-5) and torch.allclose(height, want[1], rtol=1e-4, atol=1e-5)
    finally:
        sys.path.remove(REFERENCE)


def test_fused_adamw_state_dict_is_torch_adamw_layout(vk):
    """FusedAdamW.state_dict() / load_state_dict() speak torch.optim.AdamW's per-parameter layout in model.parameters()
    order (what the reference checkpoints, train.py:94-96,307-322)."""
    from vkit_ocr_model_adaptive_scaling_b200.parallel import GradientBuckets
    from vkit_ocr_model_adaptive_scaling_b200.training import FusedAdamW
    torch.manual_seed(0)
    model, _ = _model(vk)
    ref_opt = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.02)
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    ref_opt.step()
    ref_state = ref_opt.state_dict()
    buckets = GradientBuckets(model, flatten_params=True)
    opt = FusedAdamW(buckets)
    opt.load_state_dict(ref_state)
    assert opt.steps == 1 and opt.lr == 3e-4 and tuple(opt.betas) == (0.8, 0.95) and opt.eps == 1e-7 and opt.weight_decay == 0.02
    mine = opt.state_dict()
    assert mine['param_groups'][0]['params'] == ref_state['param_groups'][0]['params']
    for idx, ent in ref_state['state'].items():
        assert torch.equal(mine['state'][idx]['exp_avg'], ent['exp_avg'])
        assert torch.equal(mine['state'][idx]['exp_avg_sq'], ent['exp_avg_sq'])
        assert float(mine['state'][idx]['step']) == float(ent['step'])
    fresh = torch.optim.AdamW(model.parameters())
    fresh.load_state_dict(mine)                                   # and torch accepts ours


def test_fused_adamw_rejects_frozen_parameters(vk):
    from vkit_ocr_model_adaptive_scaling_b200.parallel import GradientBuckets
    from vkit_ocr_model_adaptive_scaling_b200.training import FusedAdamW
    model, _ = _model(vk)
    model.backbone.stem[0].weight.requires_grad_(False)
    with pytest.raises(ValueError):
        FusedAdamW(GradientBuckets(model, flatten_params=True))


def test_cosine_warm_restarts_schedule_equals_torch(vk):
    """training.CosineWarmRestartsSchedule against torch's CosineAnnealingWarmRestarts stepped the way the reference loop
    steps it (fractional epochs, experiment/adaptive_scaling/train.py:293-298,474-477), across two restarts, and the
    state_dict interchange in both directions."""
    import torch
    from vkit_ocr_model_adaptive_scaling_b200.training import CosineWarmRestartsSchedule
    for t0, t_mult, batches in ((10, 10, 7), (3, 2, 5), (4, 1, 3)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=8e-4)
        ref = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=t0, T_mult=t_mult, eta_min=8e-6)
        ours = CosineWarmRestartsSchedule(8e-4, t0, t_mult, 8e-6)
        assert abs(ours.last_lr - ref.get_last_lr()[0]) <= 1e-12
        for epoch_idx in range(0, t0 * (1 + t_mult) + 3):
            for batch_idx in range(1, batches + 1):
                epoch = epoch_idx + (batch_idx - 1) / batches
                ref.step(epoch)
                assert abs(ours.step(epoch) - ref.get_last_lr()[0]) <= 1e-12 * 8e-4 + 1e-15, (t0, t_mult, epoch)
        theirs = ref.state_dict()
        mine = ours.state_dict()
        for k in ('T_0', 'T_i', 'T_mult', 'eta_min', 'base_lrs', 'last_epoch'):
            assert mine[k] == theirs[k], k
        assert abs(mine['T_cur'] - theirs['T_cur']) <= 1e-9 and abs(mine['_last_lr'][0] - theirs['_last_lr'][0]) <= 1e-15
        resumed = CosineWarmRestartsSchedule()
        resumed.load_state_dict(theirs)                       # a reference RestoreState's optimizer_scheduler_state_dict
        assert resumed.state_dict() == {**mine, 'T_cur': theirs['T_cur'], '_last_lr': theirs['_last_lr']}
        ref2 = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(torch.optim.AdamW([p], lr=8e-4), T_0=1)
        ref2.load_state_dict(mine)                            # ... and ours resumes the torch scheduler
        nxt = theirs['last_epoch'] + 1.25
        ref2.step(nxt)
        assert abs(resumed.step(nxt) - ref2.get_last_lr()[0]) <= 1e-15


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'vkit_open_model')), reason='reference checkout not present (GPU box)')
def test_debug_gradient_helpers_equal_the_reference(vk, caplog):
    """AdaptiveScaling.debug_get_rough_name_to_grad / debug_get_precise_name_to_grad / debug_inspect_name_to_grad (SURVEY
    8 a19; model/adaptive_scaling.py:179-237, called from train.py:420-462): the reference's static helpers and ours, run on
    the same module with the same hand-set gradients, give the same tables and log the same statistics."""
    import logging
    sys.path.insert(0, REFERENCE)
    try:
        from vkit_open_model import model as ref_model
        model, _ = _model(vk, 'fpn')
        g = torch.Generator().manual_seed(21)
        named = list(model.named_parameters())
        for i, (_, p) in enumerate(named):
            p.grad = None if i % 7 == 3 else torch.randn(p.shape, generator=g)       # some parameters without a gradient
        ours_r = vk.model.AdaptiveScaling.debug_get_rough_name_to_grad(model)
        ref_r = ref_model.AdaptiveScaling.debug_get_rough_name_to_grad(model)
        for i, (_, p) in enumerate(named):                                        # "after the precise backward"
            if i % 5 == 1:
                p.grad = None
            elif p.grad is not None:
                p.grad = p.grad + torch.randn(p.shape, generator=g)
            else:
                p.grad = torch.randn(p.shape, generator=g)
        ours_p = vk.model.AdaptiveScaling.debug_get_precise_name_to_grad(model, ours_r)
        ref_p = ref_model.AdaptiveScaling.debug_get_precise_name_to_grad(model, ref_r)
        for ours, ref in ((ours_r, ref_r), (ours_p, ref_p)):
            assert list(ours) == list(ref) and len(ours) > 100
            assert all(torch.equal(ours[k], ref[k]) for k in ref)
        numbers = []
        for cls in (vk.model.AdaptiveScaling, ref_model.AdaptiveScaling):
            caplog.clear()
            with caplog.at_level(logging.INFO):
                cls.debug_inspect_name_to_grad(ours_r, ours_p)
            text = ' '.join(r.getMessage() for r in caplog.records)
            numbers.append([float(t) for t in __import__('re').findall(r'= ([0-9.eE+-]+)', text)])
        assert len(numbers[0]) == 5 and numbers[0] == pytest.approx(numbers[1], rel=1e-6)
    finally:
        sys.path.remove(REFERENCE)
