"""The training step of the adaptive-scaling model as the reference's loop runs it
(experiment/adaptive_scaling/train.py:397-478, minus data loading and the optimizer):

    forward_rough(rough images)   -> rough loss / 2   -> backward
    forward_precise(precise imgs) -> precise loss / 2 -> backward   (gradients accumulate on the shared backbone)
    [data parallel: bucketed gradient all-reduce overlapped with the backward passes]

Losses stay on the device (the reference's ``float(loss)`` host syncs, train.py:415,453, are left to the caller).
"""
import math
from typing import Dict, Optional, Tuple

import torch

from .loss_function import AdaptiveScalingPreciseLossFunction, AdaptiveScalingRoughLossFunction
from .parallel import DataParallel

Tensor = torch.Tensor

ROUGH_KEYS = ('downsampled_mask', 'downsampled_score_map', 'downsampled_shape', 'downsampled_core_box')
PRECISE_KEYS = ('downsampled_char_prob_score_map', 'downsampled_char_mask', 'downsampled_shape', 'downsampled_core_box',
                'downsampled_label_point_y', 'downsampled_label_point_x', 'char_up_left_offsets', 'char_corner_angles',
                'char_corner_distances')


def batch_to_device(batch: Dict[str, object], device, non_blocking: bool = True) -> Dict[str, object]:
    """training/opt.py:17-27 of the reference: tensors move, everything else passes through."""
    return {k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def train_step(model, rough_loss_function: AdaptiveScalingRoughLossFunction,
               precise_loss_function: AdaptiveScalingPreciseLossFunction, rough_batch: Dict[str, object],
               precise_batch: Dict[str, object], dp: Optional[DataParallel] = None,
               label_point_forward: bool = False) -> Tuple[Tensor, Tensor]:
    """One two-pass step; returns the (un-halved) rough and precise losses as device tensors.  ``label_point_forward``
    (an extension, off by default): evaluate the precise offset / angle / distance heads at the label points only
    (``AdaptiveScaling.forward_precise(x, label_points=...)``) -- same losses and gradients, less work."""
    scale = 0.5 * (dp.loss_scale if dp is not None else 1.0)
    if dp is not None:
        dp.begin_step()
        dp.begin_pass(final=('rough',))
    mask, height = model.forward_rough(rough_batch['image'])
    rough_loss = rough_loss_function(
        rough_char_mask_feature=mask, rough_char_height_feature=height, **{k: rough_batch[k] for k in ROUGH_KEYS})
    (rough_loss * scale).backward()
    del mask, height
    if dp is not None:
        dp.begin_pass(final=None)
    lp = (precise_batch['downsampled_label_point_y'], precise_batch['downsampled_label_point_x']) if label_point_forward else None
    prob, offset, angle, distance = model.forward_precise(precise_batch['image'], label_points=lp) if lp is not None \
        else model.forward_precise(precise_batch['image'])
    precise_loss = precise_loss_function(
        precise_char_mask_feature=None, precise_char_prob_feature=prob,
        precise_char_up_left_corner_offset_feature=offset, precise_char_corner_angle_feature=angle,
        precise_char_corner_distance_feature=distance, **{k: precise_batch[k] for k in PRECISE_KEYS})
    (precise_loss * scale).backward()
    if dp is not None:
        dp.finish_step()
    return rough_loss.detach(), precise_loss.detach()


class GraphedTrainStep:
    """``train_step`` captured once into a CUDA graph and replayed: the ~1000 kernel launches of a step (every C-ABI call,
    the autograd glue, the bucket zero-fills) become one ``cudaGraphLaunch``, which removes the launch gaps between
    dependent kernels and the host time at the pass boundaries (opt-in; the eager step stays the default).

    The captured step reads its inputs from buffers this object owns: ``__call__`` copies the new batches into them
    (same keys, shapes and dtypes as the example batches; the non-tensor entries -- ``downsampled_shape``,
    ``downsampled_core_box`` -- are baked into the graph and must not change) and returns the two loss tensors, which
    are overwritten by the next replay.  Gradients land in ``dp``'s flat buckets exactly as in the eager step, stochastic
    depth draws fresh masks on every replay (torch's graph-safe Philox offsets).  With several processes the bucketed NCCL
    all-reduces are captured too: the side stream forks from the compute stream at each bucket's last gradient and joins it
    again in ``finish_step``, which is a legal cross-stream capture (measured at 2 GPUs: 87.3 -> 85.3 ms per step, gradients
    equal to the eager step's; ``close()`` the object BEFORE ``destroy_process_group()`` -- a live graph that holds captured NCCL
    kernels keeps the communicator's teardown waiting).  The parameters may change between replays (the
    fused optimizer updates them in place) only if the kernel-layout weight copies are refreshed inside the graph:
    pass ``repack_weights=True`` when an optimizer steps between replays."""

    def __init__(self, model, rough_loss_function, precise_loss_function, rough_batch: Dict[str, object],
                 precise_batch: Dict[str, object], dp: DataParallel, warmup: int = 3, label_point_forward: bool = False,
                 repack_weights: bool = False) -> None:
        from . import ops
        if dp is None:
            raise ValueError('GraphedTrainStep needs a DataParallel (its flat buckets hold the gradients)')
        own = lambda d: {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        self.rough_batch, self.precise_batch = own(rough_batch), own(precise_batch)

        def run():
            if repack_weights:
                ops.PACK.bump()
            return train_step(model, rough_loss_function, precise_loss_function, self.rough_batch, self.precise_batch, dp,
                              label_point_forward=label_point_forward)
        # warm-up on a side stream (torch's capture recipe): fills the packed-weight cache, the cached device constants and
        # the allocator, so that the capture itself records kernels only
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = run()

    @staticmethod
    def _same(a: object, b: object) -> bool:
        """Equality of the non-tensor batch entries: boxes by their four edges (the reference passes a fresh
        ``vkit.element.Box`` with every batch, whatever its ``__eq__`` does), everything else by ``==``."""
        edges = ('up', 'down', 'left', 'right')
        if all(hasattr(a, f) and hasattr(b, f) for f in edges):
            return all(int(getattr(a, f)) == int(getattr(b, f)) for f in edges)
        return bool(a == b)

    @classmethod
    def _refill(cls, own: Dict[str, object], new: Dict[str, object]) -> None:
        for k, v in own.items():
            if isinstance(v, torch.Tensor):
                if new[k] is not v:
                    if tuple(new[k].shape) != tuple(v.shape) or new[k].dtype != v.dtype:
                        raise ValueError(f'GraphedTrainStep: batch entry {k!r} is {tuple(new[k].shape)} {new[k].dtype}, the captured '
                                         f'graph reads {tuple(v.shape)} {v.dtype}')
                    v.copy_(new[k], non_blocking=True)
            elif not cls._same(new[k], v):
                raise ValueError(f'GraphedTrainStep: batch entry {k!r} is part of the captured graph and cannot change')

    def close(self) -> None:
        """Release the captured graph (its kernels, its private memory pool, the NCCL work it holds).  Call it before
        ``torch.distributed.destroy_process_group()``."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self.losses = None

    def __call__(self, rough_batch: Dict[str, object], precise_batch: Dict[str, object]) -> Tuple[Tensor, Tensor]:
        if self.graph is None:
            raise RuntimeError('GraphedTrainStep: closed')
        self._refill(self.rough_batch, rough_batch)
        self._refill(self.precise_batch, precise_batch)
        self.graph.replay()
        return self.losses


class FusedAdamW:
    """Global gradient-norm clipping + AdamW over the flat buckets of ``parallel.GradientBuckets(model, flatten_params=True)``:
    the optimizer tail of the reference loop (``clip_grad_norm_`` + ``torch.optim.AdamW.step``,
    experiment/adaptive_scaling/train.py:73-80,287-298,468-478) as one sum-of-squares pass and one update kernel per bucket,
    with the clip coefficient computed on the device (no host synchronisation).  ``step(lr=...)`` takes the learning rate of
    the caller's schedule (the reference uses CosineAnnealingWarmRestarts)."""

    def __init__(self, buckets, lr: float = 8e-4, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None) -> None:
        if not buckets.flat_params:
            raise ValueError('FusedAdamW needs GradientBuckets(model, flatten_params=True)')
        self.buckets = buckets
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        if buckets.frozen:
            raise ValueError('FusedAdamW steps whole flat buckets; parameters with requires_grad=False are not supported: '
                             + ', '.join(buckets.frozen[:4]))
        self.exp_avg = [torch.zeros_like(t) for t in buckets.flat_params]
        self.exp_avg_sq = [torch.zeros_like(t) for t in buckets.flat_params]
        self.steps = 0
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=buckets.flat_params[0].device)

    def zero_grad(self) -> None:
        self.buckets.reattach()
        self.buckets.zero()

    # ---- checkpointing in torch.optim.AdamW's layout (the reference saves / restores ``optimizer.state_dict()``,
    # experiment/adaptive_scaling/train.py:94-96,307-322,597-603): parameters are numbered in ``model.parameters()`` order
    def _param_slices(self):
        """(bucket index, offset, parameter) in the model's own parameter order."""
        where = {}
        for b, members in enumerate(self.buckets.members):
            offset = 0
            for p in members:
                where[id(p)] = (b, offset)
                offset += p.numel()
        return [(where[id(p)][0], where[id(p)][1], p) for p in self.buckets.ordered_params if id(p) in where]

    def state_dict(self) -> Dict[str, object]:
        state = {}
        for idx, (b, off, p) in enumerate(self._param_slices()):
            n = p.numel()
            state[idx] = {'step': torch.tensor(float(self.steps)),
                          'exp_avg': self.exp_avg[b][off:off + n].view(p.shape).detach().clone(),
                          'exp_avg_sq': self.exp_avg_sq[b][off:off + n].view(p.shape).detach().clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': self.weight_decay, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'decoupled_weight_decay': True, 'params': list(range(len(state)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, state_dict: Dict[str, object]) -> None:
        """Accepts ``torch.optim.AdamW.state_dict()`` of an optimizer built over ``model.parameters()`` (the reference's,
        train.py:287-292) or this class's own."""
        slices = self._param_slices()
        state = state_dict['state']
        groups = state_dict['param_groups']
        if len(groups) != 1 or len(groups[0]['params']) != len(slices):
            raise ValueError('FusedAdamW.load_state_dict expects one parameter group covering every model parameter')
        g = groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = float(g['lr']), tuple(g['betas']), float(g['eps']), float(g['weight_decay'])
        steps = set()
        for idx, (b, off, p) in zip(g['params'], slices):
            n = p.numel()
            if idx not in state:          # torch creates the per-parameter state lazily at the first step
                self.exp_avg[b][off:off + n].zero_()
                self.exp_avg_sq[b][off:off + n].zero_()
                steps.add(0)
                continue
            ent = state[idx]
            if tuple(ent['exp_avg'].shape) != tuple(p.shape):
                raise ValueError(f'optimizer state {idx}: shape {tuple(ent["exp_avg"].shape)} vs parameter {tuple(p.shape)}')
            self.exp_avg[b][off:off + n].copy_(ent['exp_avg'].reshape(-1))
            self.exp_avg_sq[b][off:off + n].copy_(ent['exp_avg_sq'].reshape(-1))
            steps.add(int(float(ent['step'])))
        if len(steps) > 1:
            raise ValueError(f'FusedAdamW keeps ONE step counter; the state holds several: {sorted(steps)}')
        self.steps = steps.pop() if steps else 0

    def grad_norm(self) -> Tensor:
        """Global L2 norm of the gradients as seen by the last ``step`` (device tensor)."""
        return self._sumsq.sqrt().float()

    def step(self, lr: Optional[float] = None, grad_scale: float = 1.0) -> None:
        from . import _lib as L
        from . import ops
        lr = self.lr if lr is None else lr
        self.steps += 1
        b1, b2 = self.betas
        bc1, bc2 = 1.0 - b1 ** self.steps, 1.0 - b2 ** self.steps
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        stream = L.stream_ptr()
        self._sumsq.zero_()
        if clip:
            for g in self.buckets.flat:
                L.check(L.LIB.vkocr_sumsq_f32(L.ptr(g), g.numel(), L.ptr(self._sumsq), stream), 'sumsq_f32')
        for p, g, m, v in zip(self.buckets.flat_params, self.buckets.flat, self.exp_avg, self.exp_avg_sq):
            L.check(L.LIB.vkocr_adamw_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, b1, b2, self.eps, self.weight_decay,
                                           bc1, bc2, L.ptr(self._sumsq) if clip else None, float(self.max_grad_norm or 0.0),
                                           grad_scale, stream), 'adamw_step')
        ops.PACK.bump()   # the parameters changed behind torch's version counters: re-pack the kernel-layout weights


class CosineWarmRestartsSchedule:
    """The learning-rate schedule of the reference loop as a plain function of the fractional epoch:
    ``torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(T_0=10, T_mult=10, eta_min=8e-6)`` stepped with
    ``epoch_idx + (batch_idx - 1) / train_num_batches`` (experiment/adaptive_scaling/train.py:73-80,293-298,474-477).
    ``FusedAdamW.step(lr=schedule.step(epoch))`` takes the result; no optimizer object is patched.  ``state_dict`` /
    ``load_state_dict`` speak the torch scheduler's keys, so the ``optimizer_scheduler_state_dict`` of a reference
    ``RestoreState`` file (train.py:91-96,323-331) resumes here and ours resumes there."""

    def __init__(self, base_lr: float = 8e-4, t0: int = 10, t_mult: int = 10, eta_min: float = 8e-6) -> None:
        if t0 <= 0 or t_mult < 1:
            raise ValueError(f'CosineWarmRestartsSchedule: T_0 {t0}, T_mult {t_mult}')
        self.base_lr, self.t0, self.t_mult, self.eta_min = float(base_lr), int(t0), int(t_mult), float(eta_min)
        self.t_i, self.t_cur, self.last_epoch = self.t0, 0.0, 0
        self.last_lr = self.lr_at(0.0)

    def _locate(self, epoch: float) -> Tuple[float, int]:
        """(position inside the current cosine period, length of that period)"""
        if epoch < 0:
            raise ValueError(f'epoch {epoch} < 0')
        if epoch < self.t0:
            return epoch, self.t0
        if self.t_mult == 1:
            return epoch % self.t0, self.t0
        n = int(math.log(epoch / self.t0 * (self.t_mult - 1) + 1, self.t_mult))
        return epoch - self.t0 * (self.t_mult ** n - 1) / (self.t_mult - 1), self.t0 * self.t_mult ** n

    def lr_at(self, epoch: float) -> float:
        t_cur, t_i = self._locate(epoch)
        return self.eta_min + (self.base_lr - self.eta_min) * (1.0 + math.cos(math.pi * t_cur / t_i)) / 2.0

    def step(self, epoch: float) -> float:
        self.t_cur, self.t_i = self._locate(epoch)
        self.last_epoch = math.floor(epoch)
        self.last_lr = self.lr_at(epoch)
        return self.last_lr

    def state_dict(self) -> Dict[str, object]:
        return {'T_0': self.t0, 'T_i': self.t_i, 'T_mult': self.t_mult, 'eta_min': self.eta_min, 'T_cur': self.t_cur,
                'base_lrs': [self.base_lr], 'last_epoch': self.last_epoch, '_last_lr': [self.last_lr]}

    def load_state_dict(self, state: Dict[str, object]) -> None:
        if len(state['base_lrs']) != 1:
            raise ValueError('CosineWarmRestartsSchedule keeps one parameter group')
        self.t0, self.t_i, self.t_mult = int(state['T_0']), int(state['T_i']), int(state['T_mult'])
        self.eta_min, self.t_cur = float(state['eta_min']), float(state['T_cur'])
        self.base_lr, self.last_epoch = float(state['base_lrs'][0]), int(state['last_epoch'])
        self.last_lr = float(state['_last_lr'][0]) if state.get('_last_lr') else \
            self.eta_min + (self.base_lr - self.eta_min) * (1.0 + math.cos(math.pi * self.t_cur / self.t_i)) / 2.0

