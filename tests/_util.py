"""Shared helpers of the GPU parity tests: differential comparison of a product module against the oracle."""
from typing import Callable, Dict, Mapping, Optional, Sequence

import torch

# north_star tolerances: fp32 mode rel 1e-4, bf16 mode rel 2e-2 — relative L2 error of every output tensor, of the
# losses, and of the gradient (all parameter gradients taken together, the vector the optimizer / global-norm clip sees).
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
GRAD_TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
# One small-shape exception, stated.  The north-star bound holds at the benchmark shapes for BOTH necks
# (tests/test_gpu_fullsize.py, B=32 / 640x640: UPERNEXT 1.1e-2, FPN 8.5e-3 against the fp32 oracle) and for the SMALL / BASE /
# LARGE configurations on two 128x192 images (test_larger_configs_against_oracle: 1.3e-2 / 1.3e-2 / 1.0e-2).  On the 2-image
# 160x224 batch of test_training_step_against_oracle (synthetic weights seed 7) the parameter gradients are sums over 45x
# fewer pixels than at the benchmark shape: their coherent part shrinks with the pixel count while the bf16 rounding noise
# of ~60 chained tensors only shrinks with its square root, and the FPN configuration (kaiming-initialised 384-channel
# laterals, forward mask-logit error 1.8e-2) measures 2.9e-2 - 3.1e-2 there; the reference's own ops under
# torch.autocast(bfloat16) measure 3.3e-2 on the same inputs (the test prints both; UPERNEXT: 1.5e-2 vs 1.5e-2).
# Anything not listed here is held to GRAD_TOL.
SMALL_SHAPE_BF16_GRAD_TOL = {'tiny/fpn': 3.5e-2}
# Individual parameter-gradient tensors are sums over up to millions of pixels: rounding noise scales with
# sqrt(sum t_i^2), not with |sum t_i|, so a tensor whose terms cancel (biases of zero-mean maps, the stem at the end of
# the longest backward chain) carries a larger *relative* error than the gradient as a whole.  Per-tensor bound:
PER_TENSOR_GRAD_TOL = {torch.float32: 3e-4, torch.bfloat16: 5e-2}
# ... and tensors whose norm is below this fraction of the global gradient norm are held to the absolute error that
# fraction implies instead (their relative error is ill-conditioned and irrelevant to the update).
NEGLIGIBLE = 1e-3


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().reshape(-1)
    b = b.detach().double().reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = float(b.norm())
    num = float((a - b).norm())
    if den == 0.0:
        return num
    return num / den


def assert_close(a: torch.Tensor, b: torch.Tensor, tol: float, what: str, atol: float = 0.0) -> None:
    assert tuple(a.shape) == tuple(b.shape), f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    assert torch.isfinite(a.detach().float()).all(), f'{what}: non-finite values'
    err = rel_err(a, b)
    if atol > 0.0 and float((a.detach().double() - b.detach().double()).abs().max()) <= atol:
        return
    assert err <= tol, f'{what}: relative L2 error {err:.3e} > {tol:.1e}'


def oracle_params(module: torch.nn.Module, dtype: torch.dtype = torch.float64) -> Dict[str, torch.Tensor]:
    """The module's weights as oracle leaves.  fp64 by default: on the GPU the oracle's own fp32 convolutions (cuDNN
    TF32 / FFT / Winograd algorithms) are noisier than the 1e-4 we are checking, fp64 is the exact arithmetic of the
    reference's formulas."""
    return {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in module.state_dict().items()}


def randomize(module: torch.nn.Module, seed: int) -> None:
    """Give every parameter an O(1) effect (block_scale is 1e-6 at init and hides errors, convnext.py:38)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            shape = tuple(p.shape)
            if name.endswith('block_scale'):
                v = torch.rand(shape, generator=g) * 0.5 + 0.5
            elif p.dim() == 1 and (name.endswith('.2.weight') or name.endswith('ln.1.weight') or name.endswith('stem.2.weight')):
                v = torch.rand(shape, generator=g) + 0.5
            elif p.dim() == 1:
                v = torch.randn(shape, generator=g) * 0.1
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                v = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
            p.copy_(v.to(p.device))


def compare_grads(module: torch.nn.Module, ref: Mapping[str, torch.Tensor], dtype: torch.dtype, what: str,
                  verbose: bool = True, grad_tol: Optional[float] = None) -> float:
    """Product parameter gradients (module.<param>.grad) against the oracle's (ref[name].grad); returns the global
    relative L2 error.  Criteria: see GRAD_TOL / PER_TENSOR_GRAD_TOL / NEGLIGIBLE above.  ``grad_tol``: an explicit bound
    for this call instead of GRAD_TOL[dtype] -- only the documented small-shape exceptions of DESIGN.md §2 pass one."""
    pairs = []
    for name, p in module.named_parameters():
        rg = ref[name].grad
        if rg is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f'{what}: {name} has a gradient, the oracle has none'
            continue
        assert p.grad is not None, f'{what}: {name} has no gradient'
        assert torch.isfinite(p.grad).all(), f'{what}: non-finite gradient of {name}'
        pairs.append((name, p.grad.detach().double().reshape(-1), rg.detach().double().reshape(-1)))
    num = sum(float((a - b).square().sum()) for _, a, b in pairs) ** 0.5
    den = sum(float(b.square().sum()) for _, a, b in pairs) ** 0.5
    global_err = num / max(den, 1e-300)
    rows = []
    for name, a, b in pairs:
        nb, nd = float(b.norm()), float((a - b).norm())
        rows.append((nd / max(nb, 1e-300), nb / max(den, 1e-300), nd, name))
    rows.sort(reverse=True)
    if verbose:
        print(f'[{what}] global gradient rel L2 error {global_err:.3e}; worst tensors:')
        for err, share, nd, name in rows[:5]:
            print(f'    {err:.3e}  (norm share {share:.2e})  {name}')
    gtol = GRAD_TOL[dtype] if grad_tol is None else grad_tol
    ptol = PER_TENSOR_GRAD_TOL[dtype] * gtol / GRAD_TOL[dtype]
    assert global_err <= gtol, f'{what}: global gradient relative L2 error {global_err:.3e} > {gtol:.1e}'
    for err, share, nd, name in rows:
        if share < NEGLIGIBLE:
            assert nd <= ptol * NEGLIGIBLE * den, \
                f'{what}: grad of {name} (negligible norm share {share:.1e}): abs L2 error {nd:.3e}'
        else:
            assert err <= ptol, f'{what}: grad of {name}: relative L2 error {err:.3e} > {ptol:.1e}'
    return global_err
