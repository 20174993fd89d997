"""Data-parallel training of the adaptive-scaling model: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference has no distributed code and ``AdaptiveScaling`` has no ``forward`` (only
``forward_rough`` / ``forward_precise``, model/adaptive_scaling.py:143-177), so ``DistributedDataParallel`` cannot
wrap it.  The one exchange step of the path is the gradient all-reduce, done here over flat fp32 buckets:

* every parameter's ``.grad`` is a view into a flat bucket (the backward kernels accumulate straight into it);
* buckets follow the *reverse execution order* of the two-pass step (train.py:397-478): ``rough`` (rough neck + heads,
  final after backward #1, so its all-reduce hides behind the whole precise pass), ``precise`` (precise neck + heads),
  then the backbone stages 3, 2 and 1+0+stem (final in that order during backward #2), so that only the last, small
  bucket is exposed;
* a bucket's all-reduce is launched on a side stream the moment the backward node of its last parameter has run
  (``ops.set_grad_ready_hook``), overlapping the rest of backward;
* the 1/world_size of the gradient mean is folded into the loss scale, so no extra pass touches the buckets.
"""
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import ops

Tensor = torch.Tensor


def bucket_plan(model: nn.Module) -> List[Tuple[str, List[str]]]:
    """Parameter names per bucket, in the order the buckets become final during one training step."""
    groups: Dict[str, List[str]] = {}
    order: List[str] = []

    def add(bucket: str, name: str) -> None:
        if bucket not in groups:
            groups[bucket] = []
            order.append(bucket)
        groups[bucket].append(name)

    names = [n for n, _ in model.named_parameters()]
    stage_ids = sorted({int(n.split('.')[2]) for n in names if n.startswith('backbone.blocks.')})
    deep = [s for s in stage_ids if s >= 2]
    for n in names:
        if n.startswith('rough_'):
            add('rough', n)
    for n in names:
        if n.startswith('precise_'):
            add('precise', n)
    for s in reversed(deep):
        for n in names:
            if n.startswith(f'backbone.blocks.{s}.'):
                add(f'backbone.stage{s}', n)
    for n in names:
        if n.startswith('backbone.') and not any(n.startswith(f'backbone.blocks.{s}.') for s in deep):
            add('backbone.shallow', n)
    seen = {n for b in order for n in groups[b]}
    for n in names:
        if n not in seen:
            add('other', n)
    return [(b, groups[b]) for b in order]


class GradientBuckets:
    """Flat fp32 gradient storage; ``param.grad`` of every parameter aliases a slice of its bucket."""

    def __init__(self, model: nn.Module, plan: Optional[Sequence[Tuple[str, Sequence[str]]]] = None,
                 flatten_params: bool = False) -> None:
        plan = bucket_plan(model) if plan is None else plan
        params = dict(model.named_parameters())
        self.ordered_params: List[nn.Parameter] = [p for _, p in model.named_parameters()]   # model.parameters() order
        self.frozen: List[str] = [n for n, p in params.items() if not p.requires_grad]
        self.names: List[str] = []
        self.flat: List[Tensor] = []
        self.flat_params: List[Tensor] = []     # only with flatten_params: param.data of every member aliases a slice
        self.members: List[List[nn.Parameter]] = []
        self.bucket_of: Dict[int, int] = {}
        for idx, (bucket, names) in enumerate(plan):
            members = [params[n] for n in names]
            total = sum(p.numel() for p in members)
            flat = torch.zeros(total, dtype=torch.float32, device=members[0].device)
            flat_p = torch.empty(total, dtype=torch.float32, device=members[0].device) if flatten_params else None
            offset = 0
            for p in members:
                p.grad = flat[offset:offset + p.numel()].view(p.shape)
                if flat_p is not None:
                    # the fused optimizer (training.FusedAdamW) updates whole buckets: parameters move into flat storage
                    # (same values, same shapes; state_dict() is unchanged)
                    view = flat_p[offset:offset + p.numel()].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                offset += p.numel()
                self.bucket_of[id(p)] = idx
            self.names.append(bucket)
            self.flat.append(flat)
            if flat_p is not None:
                self.flat_params.append(flat_p)
            self.members.append(members)

    def zero(self) -> None:
        for flat in self.flat:
            if flat.is_cuda:
                ops.zero_(flat)
            else:
                flat.zero_()

    def reattach(self) -> None:
        """Re-point ``param.grad`` at the buckets (after a ``zero_grad(set_to_none=True)``)."""
        for flat, members in zip(self.flat, self.members):
            offset = 0
            for p in members:
                if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * offset:
                    p.grad = flat[offset:offset + p.numel()].view(p.shape)
                offset += p.numel()


class DataParallel:
    """Hand-rolled bucketed gradient all-reduce for the two-pass training step.

    Usage per step::

        dp.begin_step()                       # zero the buckets
        dp.begin_pass(final=('rough',))       # buckets that receive their last contribution in this backward
        (rough_loss * dp.loss_scale).backward()
        dp.begin_pass(final=None)             # None: every remaining bucket
        (precise_loss * dp.loss_scale).backward()
        dp.finish_step()                      # current stream waits for all reductions
    """

    def __init__(self, model: nn.Module, process_group=None, broadcast_parameters: bool = True,
                 flatten_params: bool = False) -> None:
        self.model = model
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = GradientBuckets(model, flatten_params=flatten_params)
        self.loss_scale = 1.0 / self.world_size
        self._pending: Dict[int, int] = {}
        self._works: List = []
        self._launched: List[int] = []
        self._done = set()
        dev = self.buckets.flat[0].device
        self._cuda = dev.type == 'cuda'
        self._side = torch.cuda.Stream(device=dev) if self._cuda else None
        if broadcast_parameters and self.world_size > 1:
            for p in model.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            # an in-place write through .data does not move the version counters the packed-weight cache is keyed on
            ops.PACK.bump()
        ops.set_grad_ready_hook(self._on_ready)

    # ------------------------------------------------------------------------------------------------ step protocol
    def begin_step(self) -> None:
        self.buckets.reattach()
        self.buckets.zero()
        self._works.clear()
        self._launched.clear()
        self._done.clear()
        self._pending.clear()

    def begin_pass(self, final: Optional[Iterable[str]] = None) -> None:
        """Arm the buckets whose gradients become final during the next backward (prefix match on bucket names;
        ``None`` arms every bucket that has not been reduced yet)."""
        self._pending.clear()
        prefixes = None if final is None else tuple(final)
        for idx, name in enumerate(self.buckets.names):
            if idx in self._done:
                continue
            if prefixes is None or name.startswith(prefixes):
                self._pending[idx] = len(self.buckets.members[idx])

    def _on_ready(self, params: Sequence[Tensor]) -> None:
        for p in params:
            idx = self.buckets.bucket_of.get(id(p))
            if idx is None or idx not in self._pending:
                continue
            self._pending[idx] -= 1
            if self._pending[idx] == 0:
                del self._pending[idx]
                self._launch(idx)

    def _launch(self, idx: int) -> None:
        self._done.add(idx)
        self._launched.append(idx)
        if self.world_size == 1:
            return
        flat = self.buckets.flat[idx]
        if self._cuda:
            # the side stream picks up after everything queued so far on the compute stream (the kernels that wrote
            # this bucket), then NCCL runs there while backward continues on the compute stream
            self._side.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._side):
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish_step(self) -> List[str]:
        """Reduce whatever was armed but never signalled (parameters that got no gradient), then make the compute stream
        wait for every reduction.  Returns the bucket names in launch order (for tests / tracing)."""
        for idx in sorted(self._pending):
            self._launch(idx)
        self._pending.clear()
        for idx in range(len(self.buckets.names)):
            if idx not in self._done:
                self._launch(idx)
        for work in self._works:
            work.wait()
        if self._cuda and self.world_size > 1:
            torch.cuda.current_stream(self.buckets.flat[0].device).wait_stream(self._side)
        self._works.clear()
        return [self.buckets.names[i] for i in self._launched]

    def close(self) -> None:
        ops.set_grad_ready_hook(None)

    def grad_norm(self) -> Tensor:
        """Global L2 norm of the (reduced) gradients, from the flat buckets (clip_grad_norm_ of train.py:468-472)."""
        sq = [flat.double().square().sum() for flat in self.buckets.flat]
        return torch.stack(sq).sum().sqrt().float()
