"""Shared checkers of the parity tests (test infrastructure).

`gradient_mismatches` is the one gradient comparison of the model-level tests: per parameter tensor the max-norm
relative error AND the cosine similarity against a reference gradient.  A zeroed tensor reads error 1.0 / cosine 0, a
sign flip error 2.0 / cosine -1 - both far outside any tolerance used with it (tests assert that explicitly on a
deliberately broken copy, so the check is known to be able to fail)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    den = a.norm() * b.norm()
    return float((a @ b) / den) if den > 0 else (1.0 if a.norm() == b.norm() else 0.0)


def gradient_mismatches(grads: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float,
                        min_cos: float) -> List[Tuple[str, float, float]]:
    """[(name, rel err, cosine)] of every tensor of `ref` whose gradient in `grads` is off by more than `tol` (max-norm
    relative) or whose direction differs (cosine < min_cos)."""
    bad = []
    for name, g_ref in ref.items():
        e, c = rel(grads[name], g_ref), cosine(grads[name], g_ref)
        if not (e <= tol and c >= min_cos):
            bad.append((name, e, c))
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
