"""Fused optimizer tail against torch: clip_grad_norm_ + torch.optim.AdamW on the same parameters and
gradients (experiment/adaptive_scaling/train.py:73-80,468-478)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vk():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    return vk


@pytest.mark.parametrize('max_norm', [None, 2.5, 1e6])
def test_fused_adamw_matches_torch(vk, max_norm):
    from vkit_ocr_model_adaptive_scaling_b200.parallel import GradientBuckets
    from vkit_ocr_model_adaptive_scaling_b200.training import FusedAdamW
    dev = torch.device('cuda')
    M = vk.model
    torch.manual_seed(3)
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY,
                                                      neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT)).to(dev)
    ref = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    keys_before = list(model.state_dict().keys())
    buckets = GradientBuckets(model, flatten_params=True)
    assert list(model.state_dict().keys()) == keys_before
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p.detach(), r.detach())          # flattening moved the storage, not the values
    opt = FusedAdamW(buckets, lr=8e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm)
    topt = torch.optim.AdamW(ref, lr=8e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    g = torch.Generator(device='cuda').manual_seed(11)
    for step in range(4):
        opt.zero_grad()
        for p, r in zip(model.parameters(), ref):
            grad = torch.randn(p.shape, device=dev, generator=g) * (0.01 if step % 2 else 1.0)
            p.grad.copy_(grad)                               # bucket view, as the backward kernels fill it
            r.grad = grad.clone()
        lr = 8e-4 * (1.0 - 0.1 * step)
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        for grp in topt.param_groups:
            grp['lr'] = lr
        topt.step()
        opt.step(lr=lr)
        if max_norm is not None:
            assert abs(float(opt.grad_norm()) - float(total)) <= 1e-5 * float(total)
    num = sum(float((p.detach().double() - r.detach().double()).square().sum()) for p, r in zip(model.parameters(), ref))
    den = sum(float(r.detach().double().square().sum()) for r in ref)
    assert (num / den) ** 0.5 <= 1e-6, (num / den) ** 0.5
    worst = max(float((p.detach() - r.detach()).abs().max()) for p, r in zip(model.parameters(), ref))
    assert worst <= 2e-6, worst


def test_fused_adamw_invalidates_packed_weights(vk):
    """The kernels read packed bf16 copies of the weights; an optimizer step outside torch's version counters must
    refresh them (otherwise the next forward would use stale weights)."""
    from vkit_ocr_model_adaptive_scaling_b200.parallel import GradientBuckets
    from vkit_ocr_model_adaptive_scaling_b200.training import FusedAdamW
    dev = torch.device('cuda')
    M = vk.model
    torch.manual_seed(5)
    backbone = M.ConvNext.create_tiny().to(dev).eval()
    x = torch.rand(1, 3, 64, 64, device=dev) * 255
    with vk.precision(torch.bfloat16), torch.no_grad():
        before = [f.float().clone() for f in backbone(x)]
    buckets = GradientBuckets(backbone, plan=[('all', [n for n, _ in backbone.named_parameters()])], flatten_params=True)
    opt = FusedAdamW(buckets, lr=1e-2, weight_decay=0.0)
    for p in backbone.parameters():
        p.grad.fill_(1.0)
    opt.step()
    with vk.precision(torch.bfloat16), torch.no_grad():
        after = [f.float() for f in backbone(x)]
    assert any(float((a - b).abs().max()) > 1e-3 for a, b in zip(after, before))
