"""Synthetic multi-band tiles (no datasets are reachable offline).

`uniform_tiles` is the BASELINE/SURVEY 8(d) benchmark input: uint8 uniform on [0,255] (seed 1234), labels uniform
(seed 4321).  `aerial_like_tiles` draws smooth low-frequency fields plus noise and derives the label from the bands,
which gives gradients a consistent direction (a well-conditioned parity problem, unlike labels of pure noise)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def uniform_tiles(n: int, c: int, h: int, w: int, n_classes: int, seed: int = 1234, label_seed: int = 4321):
    x = torch.randint(0, 256, (n, c, h, w), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)
    y = torch.randint(0, n_classes, (n, h, w), generator=torch.Generator().manual_seed(label_seed), dtype=torch.uint8)
    return x, y


def aerial_like_tiles(n: int, c: int, h: int, w: int, n_classes: int, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand((n, c, max(2, h // 32), max(2, w // 32)), generator=g)
    mid = torch.rand((n, c, max(2, h // 8), max(2, w // 8)), generator=g)
    f = 0.6 * F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) \
        + 0.3 * F.interpolate(mid, size=(h, w), mode="bilinear", align_corners=False) \
        + 0.1 * torch.rand((n, c, h, w), generator=g)
    x = (f.clamp(0, 1) * 255).round().to(torch.uint8)
    idx = f[:, -1] - f[:, 0] if c > 1 else f[:, 0] - 0.5     # "NIR minus red"-style index
    lo, hi = idx.min(), idx.max()
    y = ((idx - lo) / (hi - lo + 1e-9) * n_classes).floor().clamp(0, n_classes - 1).to(torch.uint8)
    return x, y
