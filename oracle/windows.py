"""ORACLE (test infrastructure) — sliding-window tile offsets.

PARITY UNPINNED: the reference calls the third-party package slidingwindow==0.0.14 (environment/requirements.txt:9),
which is absent from /root/reference and from this image.  This restates its published algorithm
(slidingwindow/SlidingWindow.py::generate -> generateForSize) for the one call the reference makes:

    slidingwindow.generate(numpy_image, DimOrder.HeightWidthChannel, patch_size, patch_overlap)
                                                                          create_tiles_unet.py:52-54

    windowSizeX = min(maxWindowSize, width), windowSizeY = min(maxWindowSize, height)
    windowOverlap = int(floor(windowSize * overlapPercent));  step = windowSize - windowOverlap
    lastX = width - windowSizeX;  xOffsets = range(0, lastX+1, stepSizeX), plus lastX if not already last (same for y)
    windows emitted x-outer / y-inner;  window.indices() = (slice(y, y+h), slice(x, x+w));  getRect() = (x, y, w, h)
"""
import math
from typing import List, Tuple


def axis_offsets(dim: int, patch_size: int, patch_overlap: float) -> Tuple[List[int], int]:
    win = min(patch_size, dim)
    overlap = int(math.floor(win * patch_overlap))
    step = win - overlap
    last = dim - win
    offs = list(range(0, last + 1, step))
    if len(offs) == 0 or offs[-1] != last:
        offs.append(last)
    return offs, win


def compute_windows(height: int, width: int, patch_size: int, patch_overlap: float) -> List[Tuple[int, int, int, int]]:
    """-> list of (x, y, w, h) in the order the reference enumerates them (index = tile number in file names,
    create_tiles_unet.py:408-431)."""
    if patch_overlap > 1:
        raise ValueError(f"Patch overlap {patch_overlap} must be between 0 - 1")  # create_tiles_unet.py:48-49
    xs, ww = axis_offsets(width, patch_size, patch_overlap)
    ys, wh = axis_offsets(height, patch_size, patch_overlap)
    return [(x, y, ww, wh) for x in xs for y in ys]


def keep_window(crop, max_empty: float) -> bool:
    """create_tiles_unet.py:414 — a crop is kept iff count_nonzero >= size * (1 - max_empty)."""
    import numpy as np
    return not (np.sum(crop != 0) < crop.size * (1 - max_empty))


def tile_geotransform(gt, x: int, y: int):
    """create_tiles_unet.py:224-226 — shifted GDAL geotransform of a tile (note: the reference uses the x pixel size
    for the y origin as well)."""
    return (x * gt[1] + gt[0], gt[1], gt[2], gt[3] - y * gt[1], gt[4], gt[5])
