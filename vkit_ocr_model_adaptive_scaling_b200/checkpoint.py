"""Checkpoint / export interop with the reference's train script.

The reference saves ``torch.save(cattrs.unstructure(RestoreState(...)))`` — a plain dict with the keys ``epoch_idx``,
``model_jit_state_dict``, ``optimizer_state_dict`` (torch.optim.AdamW) and ``optimizer_scheduler_state_dict``
(CosineAnnealingWarmRestarts) — as ``state_dict_{epoch}.pt`` / ``state_dict_{epoch}_not_best.pt``
(experiment/adaptive_scaling/train.py:91-96,586-603), restores from it (:307-338) and builds the deployed TorchScript file
from it (:608-644; loaded by inferencing/adaptive_scaling.py:85-90).  The modules of this package keep the reference's
``state_dict`` layout key for key, so both directions are plain ``load_state_dict`` calls; this module adds the file
format, the optimizer-state mapping of ``training.FusedAdamW`` and the export into a reference (eager or scripted) module.
Plain host-side Python: nothing here is on the GPU hot path.
"""
from typing import Any, Dict, Mapping, Optional

import torch
from torch import nn

RESTORE_STATE_KEYS = ('epoch_idx', 'model_jit_state_dict', 'optimizer_state_dict', 'optimizer_scheduler_state_dict')


def load_restore_state(path, map_location='cpu') -> Dict[str, Any]:
    """The reference's ``RestoreState`` record (train.py:91-96) as a dict, from a ``.pt`` file written by train.py:597-603
    (or by ``save_restore_state``).  Raises ``ValueError`` when the file is not such a record."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    if not isinstance(state, Mapping) or 'model_jit_state_dict' not in state:
        raise ValueError(f'{path}: not a RestoreState checkpoint (expected the keys {RESTORE_STATE_KEYS})')
    missing = [k for k in RESTORE_STATE_KEYS if k not in state]
    if missing:
        raise ValueError(f'{path}: RestoreState checkpoint lacks {missing}')
    return dict(state)


def load_model_state(model: nn.Module, restore_state: Mapping[str, Any], strict: bool = True):
    """``model_jit.load_state_dict(restore_state.model_jit_state_dict)`` (train.py:317) for a module of this package."""
    from . import ops
    result = model.load_state_dict(restore_state['model_jit_state_dict'], strict=strict)
    ops.PACK.bump()          # load_state_dict copies in place: the kernel-layout weight copies are stale
    return result


def save_restore_state(path, epoch_idx: int, model: nn.Module, optimizer_state_dict: Optional[Mapping[str, Any]] = None,
                       optimizer_scheduler_state_dict: Optional[Mapping[str, Any]] = None) -> None:
    """Writes the reference's checkpoint layout (train.py:597-603): the file restores into the reference's train loop
    (:307-338) and into ``build_model_jit_from_state_dict_path`` (:608-633).  ``optimizer_state_dict`` is
    ``torch.optim.AdamW.state_dict()`` or ``training.FusedAdamW.state_dict()`` (same layout)."""
    record = {
        'epoch_idx': int(epoch_idx),
        'model_jit_state_dict': {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
        'optimizer_state_dict': dict(optimizer_state_dict) if optimizer_state_dict is not None else {},
        'optimizer_scheduler_state_dict': dict(optimizer_scheduler_state_dict) if optimizer_scheduler_state_dict is not None else {},
    }
    torch.save(record, path)


def export_to_reference_module(model: nn.Module, reference_module: nn.Module) -> nn.Module:
    """Loads this package's weights into a module of the REFERENCE implementation (eager ``vkit_open_model.model.
    AdaptiveScaling`` or its ``torch.jit.script``-ed form: their key sets are identical) with ``strict=True``."""
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    reference_module.load_state_dict(sd, strict=True)
    return reference_module


def build_reference_model_jit(model: nn.Module, config, output_model_jit=None):
    """The reference's ``build_model_jit_from_state_dict_path`` / ``build_and_dump_model_jit_from_state_dict_path``
    (train.py:608-644) fed from a live module of this package and its ``AdaptiveScalingConfig``: reference ``AdaptiveScaling(config)`` -> ``torch.jit.script``
    -> our weights -> ``eval()`` (-> ``torch.jit.save``), i.e. the TorchScript file ``AdaptiveScalingInferencing`` loads
    (inferencing/adaptive_scaling.py:85-90).  Needs the reference package ``vkit_open_model`` to be importable."""
    try:
        from vkit_open_model import model as ref_model   # type: ignore
    except ImportError as exc:   # pragma: no cover - depends on the deployment
        raise RuntimeError('build_reference_model_jit needs the reference package vkit_open_model on sys.path') from exc
    cfg = config                                          # this package's AdaptiveScalingConfig (same fields / enum values)
    ref_cfg = ref_model.AdaptiveScalingConfig(
        size=ref_model.AdaptiveScalingSize(cfg.size.value),
        neck_head_type=ref_model.AdaptiveScalingNeckHeadType(cfg.neck_head_type.value),
        rough_upsampling_factor=cfg.rough_upsampling_factor,
        rough_init_char_height_output_bias=cfg.rough_init_char_height_output_bias,
        precise_upsampling_factor=cfg.precise_upsampling_factor,
        precise_enable_char_mask_head=cfg.precise_enable_char_mask_head,
    )
    model_jit = torch.jit.script(ref_model.AdaptiveScaling(ref_cfg))
    export_to_reference_module(model, model_jit)
    model_jit.eval()
    if output_model_jit is not None:
        torch.jit.save(model_jit, output_model_jit)
    return model_jit
