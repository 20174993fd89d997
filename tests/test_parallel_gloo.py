"""Host-side logic of the data-parallel path on CPU: world_size-2 `gloo` process group, flat gradient buckets, bucket
arming per backward pass, launch order, and the reduced result.  No kernels run here: the gradients
are written by hand into the bucket views exactly where the backward kernels would accumulate them, and the
grad-ready hook is fired the way `ops._ready` fires it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class _Stage(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.w = nn.Parameter(torch.zeros(n, 3))
        self.b = nn.Parameter(torch.zeros(n))


class _Backbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = _Stage(2)
        self.blocks = nn.ModuleList([_Stage(3), _Stage(4), _Stage(5), _Stage(6)])


class _Toy(nn.Module):
    """Same top-level attribute names as AdaptiveScaling (model/adaptive_scaling.py:51-141)."""

    def __init__(self):
        super().__init__()
        self.backbone = _Backbone()
        self.rough_neck = _Stage(7)
        self.rough_char_mask_head = _Stage(2)
        self.precise_neck = _Stage(8)
        self.precise_char_prob_head = _Stage(2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, results) -> None:
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from vkit_ocr_model_adaptive_scaling_b200 import ops
        from vkit_ocr_model_adaptive_scaling_b200.parallel import DataParallel, bucket_plan
        torch.manual_seed(100 + rank)      # different initial weights per rank: the constructor must broadcast rank 0's
        model = _Toy()
        with torch.no_grad():
            for p in model.parameters():
                p.normal_()
        plan = bucket_plan(model)
        assert [b for b, _ in plan] == ['rough', 'precise', 'backbone.stage3', 'backbone.stage2', 'backbone.shallow'], plan
        dp = DataParallel(model)
        ref = [torch.zeros_like(p) for p in model.parameters()]
        for r, p in zip(ref, model.parameters()):
            r.copy_(p.detach())
            dist.broadcast(r, src=0)
            assert torch.equal(r, p.detach()), 'parameters were not broadcast from rank 0'
        names = dict(model.named_parameters())

        def write(prefix: str, value: float) -> None:
            ps = [p for n, p in names.items() if n.startswith(prefix)]
            for p in ps:
                p.grad.add_(value * (rank + 1))     # what a backward kernel does: accumulate into the bucket view
            ops._ready(*ps)

        launched = None
        for step in range(2):                       # twice: begin_step must re-zero and re-arm
            dp.begin_step()
            dp.begin_pass(final=('rough',))
            write('rough_', 1.0)
            write('backbone.', 0.5)                 # backbone grads of pass 1 are NOT final: bucket must not launch
            assert dp._launched == [0], dp._launched
            dp.begin_pass(final=None)
            write('precise_', 2.0)
            for s in (3, 2, 1, 0):
                write(f'backbone.blocks.{s}.', 0.25)
            write('backbone.stem.', 0.25)
            launched = dp.finish_step()
        assert launched == ['rough', 'precise', 'backbone.stage3', 'backbone.stage2', 'backbone.shallow'], launched
        tot = sum(r + 1 for r in range(world))      # SUM over ranks of (rank + 1)
        for n, p in names.items():
            want = (1.0 if n.startswith('rough_') else 2.0 if n.startswith('precise_') else 0.75) * tot
            assert torch.allclose(p.grad, torch.full_like(p.grad, want)), (n, float(p.grad.flatten()[0]), want)
        assert dp.loss_scale == 1.0 / world
        flat_norm = float(dp.grad_norm())
        per_tensor = float(torch.sqrt(sum(p.grad.double().square().sum() for p in model.parameters())))
        assert abs(flat_norm - per_tensor) <= 1e-5 * per_tensor
        dp.close()
        results[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_bucketed_allreduce_world2_gloo():
    world = 2
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        assert dict(results) == {0: 'ok', 1: 'ok'}


def test_bucket_plan_covers_every_parameter_once():
    from vkit_ocr_model_adaptive_scaling_b200 import model as M
    from vkit_ocr_model_adaptive_scaling_b200.parallel import bucket_plan
    for neck in (M.AdaptiveScalingNeckHeadType.UPERNEXT, M.AdaptiveScalingNeckHeadType.FPN):
        m = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=neck))
        plan = bucket_plan(m)
        flat = [n for _, names in plan for n in names]
        assert sorted(flat) == sorted(n for n, _ in m.named_parameters())
        assert [b for b, _ in plan] == ['rough', 'precise', 'backbone.stage3', 'backbone.stage2', 'backbone.shallow']
        sizes = {b: sum(dict(m.named_parameters())[n].numel() for n in names) for b, names in plan}
        # only the small shallow bucket (stage 1 + stage 0 + stem) is exposed after backward ends
        assert sizes["backbone.shallow"] < 0.12 * sizes["backbone.stage3"]
