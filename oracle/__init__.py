"""CPU oracle for the adaptive-scaling hot path (TEST INFRASTRUCTURE ONLY).

This package is a from-scratch *functional* restatement, in plain fp32 PyTorch ops driven by a
``state_dict``, of the reference's ``vkit_open_model.model`` forward pass and
``vkit_open_model.loss_function`` losses.  It exists to check the CUDA product path; it is never the
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

Parity pinning: the reference's own tests hold **no numeric golden values** for this path (they are
shape/smoke tests), so the oracle is pinned against *outputs of the reference itself*:
``oracle/make_golden.py`` imports ``/root/reference`` in the build container, runs it on seeded inputs
and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays those fixtures through this
oracle on every box (the reference cannot travel to the GPU box, the fixtures do).
"""
from .model import (  # noqa: F401
    ModelSpec,
    convnext_forward,
    neck_forward,
    head_forward,
    forward_rough,
    forward_precise,
)
from .loss import (  # noqa: F401
    Box,
    rough_loss,
    precise_loss,
    focal_with_logits,
    dice,
    l1,
    l2,
    cross_entropy_with_logits,
    weighted_bce_with_logits,
    weight_adaptive_heatmap_regression,
)
from .synth import (  # noqa: F401
    synth_state_dict,
    synth_rough_batch,
    synth_precise_batch,
)
