"""Shared checkers of the parity tests (test infrastructure).

`gradient_mismatches` is the one gradient comparison of the model-level tests: per parameter tensor the relative L2
error, the max-norm relative error AND the cosine similarity against a reference gradient.  A zeroed tensor reads
error 1.0 / cosine 0, a sign flip error 2.0 / cosine -1, a 5 % scale error 0.05 - all outside the tolerances used with
it (tests assert that explicitly on deliberately broken copies, so the check is known to be able to fail)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    den = a.norm() * b.norm()
    return float((a @ b) / den) if den > 0 else (1.0 if a.norm() == b.norm() else 0.0)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / ||b||_2 in fp64"""
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def gradient_mismatches(grads: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float, min_cos: float,
                        tol_max: float = 0.1, floors: Optional[Dict[str, float]] = None) -> List[Tuple[str, float, float, float]]:
    """[(name, L2 rel err, max-norm rel err, cosine)] of every tensor of `ref` whose gradient in `grads` is off.

    A tensor passes when its relative L2 error is <= tol (a zeroed tensor reads 1.0, a sign flip 2.0, a 5 % scale error
    0.05), its direction agrees (cosine >= min_cos) and no single element is off by more than tol_max of the tensor
    maximum.  The L2 norm averages the bf16 rounding flips of the stored activation gradients, whose realisation changes
    with every change of fp32 summation order in the forward pass; the max-norm alone moved between 2e-2 and 7e-2 on the
    same tensor for the same code.  floors: for tensors that are one cancelling sum (a handful of elements: the
    SelfAttention gamma, 1-channel heads) the measured sensitivity of the reference itself to sub-ulp perturbations of the
    stored gradients; such a tensor passes up to 4 x its floor."""
    bad = []
    for name, g_ref in ref.items():
        e2, em, c = rel_l2(grads[name], g_ref), rel(grads[name], g_ref), cosine(grads[name], g_ref)
        t = tol
        if floors is not None and name in floors:
            t = max(tol, 4.0 * floors[name])
            if e2 <= t:
                continue                       # a scalar has no direction beyond its sign
        if not (e2 <= t and em <= max(tol_max, t) and c >= min_cos):
            bad.append((name, e2, em, c))
    return bad


def plan_taps(net) -> Dict[str, torch.Tensor]:
    """the CUDA plan's stored activations as NCHW fp32 tensors in torch channel order, keyed by the plan's activation
    names (the rounding point names of oracle/bf16_emulation.py).  The 1x1 convolutions of PixelShuffle_ICNR keep their
    output channels in (i, j, c) order in the plan (GEMM row (2i+j)*c + c_ holds torch channel 4 c_ + 2i + j)."""
    from unet_b200.layout import shuffle_row_of_co
    out = {}
    for name, a in net.named_acts.items():
        t = a.t[..., :a.C].permute(0, 3, 1, 2).float().contiguous()
        if name.endswith(".shuf.0.out") or name == "layers.8.0.out":
            t = t[:, torch.tensor(shuffle_row_of_co(a.C), device=t.device)].contiguous()
        out[name] = t
    return out
