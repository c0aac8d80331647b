"""Host-side mirror of the reference's Python seams for the hot path (same names, argument meaning and error behaviour),
running on the B200 plan instead of fastai.

    reference                                               here
    --------------------------------------------------------------------------------------------------------------
    train.unet_learner_MS(dls, arch, ...)  train.py:98      unet_learner_MS(n_in, n_classes, arch, size, batch_size, ...)
    learn.fit_one_cycle(epochs, lr_max=slice(lr/ef, lr))    Learner.fit_one_cycle(epochs, lr_max, train_batches, valid_batches)
                                           train.py:246-250
    learn.predict(tile) -> (dec, argmax, probs)             Learner.predict(tile_u8)
                                           predict.py:193
    learn.export(path) / load_learner(path) train.py:373    Learner.export(path) / load_learner(path)   (state_dict with fastai keys)
    predict.save_predictions(...)          predict.py:146   save_predictions(...) over GeoTIFF (or .npy) tiles
    train.train_func(...)                  train.py:287     train_func(...) over data_path/{trai,vali}/{img,mask}_tiles/*.tif
    create_tiles_unet.compute_windows      :30              unet_b200.tiling.compute_windows

Data enters as uint8 tiles `[B, n_in, H, W]` (the reference reads GeoTIFF bands as int32 -> float32 and fastai divides by
255, data.py:24 + IntToFloatTensor; both happen inside the cast kernel) and class-id masks `[B, H, W]`.
Tiles on disk are GeoTIFFs (`unet_b200/geotiff.py`, a numpy restatement of the rasterio/GDAL calls of the path) or `.npy`
arrays; `train_func` / `save_predictions` keep the reference's positional signatures, `predict_geotiff` is the tile-free
whole-raster variant.
"""
from __future__ import annotations

import csv
import math
import os
import time
import warnings
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .engine import Trainer, one_cycle
from .network import UNetB200
from .predict_engine import TiledPredictor
from .geotiff import GeoInfo, geotiff_info, open_mask, open_tile, read_geotiff, write_geotiff
from .tiling import colour_classes, compute_windows, placement_from_geotransform, shard_windows_by_columns

ARCHITECTURES = ("xresnet18", "xresnet34", "xresnet50", "xresnet101")   # params_and_main.py:12,99


def _arch_name(arch) -> str:
    name = arch if isinstance(arch, str) else getattr(arch, "__name__", str(arch))
    if name not in ARCHITECTURES:
        # the reference only works with xresnet bodies: body[0][0] must be a ConvLayer (train.py:130)
        raise TypeError(f"architecture {name!r} is not an xresnet body; choose from {ARCHITECTURES}")
    return name


class Learner:
    """What `unet_learner_MS` returns: model + loss + optimizer + schedule, on one GPU (or one rank of a DP job)."""

    def __init__(self, arch: str, n_in: int, n_classes: int, size: Tuple[int, int], batch_size: int,
                 class_weights: Optional[Sequence[float]] = None, opt_func: str = "adam", lr: float = 1e-3,
                 wd: float = 0.01, encoder_factor: float = 10.0, moms: Sequence[float] = (0.95, 0.85, 0.95),
                 self_attention: bool = False):
        self.arch, self.n_in, self.n_classes, self.size, self.bs = arch, n_in, n_classes, tuple(size), batch_size
        self.class_weights = list(class_weights) if class_weights is not None else None
        self.self_attention = bool(self_attention)
        self.net = UNetB200(arch, n_in, n_classes, self.size, batch_size, training=True,
                            class_weights=self.class_weights, self_attention=self.self_attention)
        self.net.init_parameters(seed=0, randomize_bn=False)     # fastai defaults: gamma 1 / 0 (BatchZero), beta 1e-3
        self.trainer = Trainer(self.net, optimizer=opt_func, lr=lr, wd=wd, encoder_factor=encoder_factor)
        self.lr, self.moms = lr, tuple(moms)
        self._eval: Optional[UNetB200] = None
        self.history: List[Dict[str, float]] = []

    # ---- training ------------------------------------------------------------------------------------------------
    def fit_one_cycle(self, epochs: int, lr_max: Optional[float] = None,
                      train_batches: Optional[Callable[[], Iterable]] = None,
                      valid_batches: Optional[Callable[[], Iterable]] = None, history_csv: Optional[str] = None,
                      monitor: str = "dice_multi", best_path: Optional[str] = None) -> List[Dict[str, float]]:
        """fastai fit_one_cycle: cosine warm-up over the first 25 % from lr/25, anneal to lr/1e5, momentum 0.95->0.85->0.95;
        CSVLogger columns epoch,train_loss,valid_loss,dice_multi,time (history.csv:1); SaveModelCallback keeps the best."""
        assert train_batches is not None, "train_batches: callable returning an iterable of (x_u8, y) batches"
        lr_max = self.lr if lr_max is None else lr_max
        n_per_epoch = sum(1 for _ in train_batches())
        total, step = max(1, epochs * n_per_epoch), 0
        best = None
        for ep in range(epochs):
            t0 = time.time()
            run, n = 0.0, 0
            it = iter(train_batches())
            cur = next(it, None)
            while cur is not None:
                nxt = next(it, None)
                x, y = cur
                lr, mom = one_cycle(step / total, lr_max, moms=self.moms)
                if self.trainer.optimizer == "adam":
                    self.trainer.set_adam_hyper(lr, mom)
                else:
                    self.trainer.lr = lr
                loss = self.trainer.step(x, y, prefetch=nxt)       # the next batch's H2D copy runs behind this step
                run += float(loss.item())
                n += 1
                step += 1
                cur = nxt
            row = {"epoch": ep, "train_loss": run / max(1, n)}
            if valid_batches is not None:
                row["valid_loss"], row["dice_multi"] = self.validate(valid_batches())
            row["time"] = time.time() - t0
            self.history.append(row)
            score = row.get(monitor)
            if best_path and score is not None and (best is None or (score < best if "loss" in monitor else score > best)):
                best = score
                torch.save(self.state_dict(), best_path)
        if best_path and best is not None and os.path.exists(best_path):
            # fastai SaveModelCallback(monitor, fname='best-model') reloads the best epoch after fit (train.py:209)
            self.load_state_dict(torch.load(best_path, map_location="cpu"))
        if history_csv:
            with open(history_csv, "w", newline="") as f:
                wr = csv.writer(f)
                wr.writerow(["epoch", "train_loss", "valid_loss", "dice_multi", "time"])
                for r in self.history:
                    m, s = divmod(int(r["time"]), 60)
                    wr.writerow([r["epoch"], r["train_loss"], r.get("valid_loss", ""), r.get("dice_multi", ""), f"{m:02d}:{s:02d}"])
        return self.history

    def _eval_net(self) -> UNetB200:
        if self._eval is None:
            self._eval = UNetB200(self.arch, self.n_in, self.n_classes, self.size, self.bs, training=False,
                                  class_weights=self.class_weights, self_attention=self.self_attention)
        self._eval.load_state_dict(self.net.state_dict())
        return self._eval

    def validate(self, batches: Iterable) -> Tuple[float, float]:
        """valid_loss (weighted CE, mean over batches) and fastai DiceMulti (macro Dice over classes present)."""
        net = self._eval_net()
        C_ = self.n_classes
        inter = torch.zeros(C_, dtype=torch.float64)
        psum = torch.zeros(C_, dtype=torch.float64)
        tsum = torch.zeros(C_, dtype=torch.float64)
        losses = []
        w = net.class_weights
        for x, y in batches:
            x, y = x.to(net.device), y.to(net.device)
            net.set_input(x.contiguous())
            net.forward()
            logits = net.logits_nchw()
            losses.append(float(torch.nn.functional.cross_entropy(logits, y.long(), weight=w)))
            pred = logits.argmax(1)
            for c in range(C_):
                p, t = pred == c, y == c
                inter[c] += float((p & t).sum())
                psum[c] += float(p.sum())
                tsum[c] += float(t.sum())
        den = psum + tsum
        dice = [2 * inter[c] / den[c] for c in range(C_) if den[c] > 0]
        return float(np.mean(losses)) if losses else float("nan"), float(np.mean(dice)) if dice else float("nan")

    # ---- inference -------------------------------------------------------------------------------------------------
    def predict(self, tile_u8):
        """fastai `Learner.predict`: returns (decoded mask, argmax [H,W], probabilities [C,H,W]) as CPU tensors."""
        t = torch.as_tensor(tile_u8)
        if t.dim() != 3 or t.shape[0] != self.n_in:
            raise ValueError(f"expected a [{self.n_in}, H, W] tile, got {tuple(t.shape)}")
        net = self._eval_net()
        probs, amax = TiledPredictor(net).predict_tiles(t[None].to(net.device).contiguous())
        return amax[0].cpu(), amax[0].cpu(), probs[0].cpu()

    # ---- persistence -------------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self.net.state_dict()

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.net.load_state_dict(sd)

    def export(self, path) -> None:
        """`learn.export` (train.py:373): a plain torch checkpoint with fastai state_dict keys + the constructor args
        (a pickled fastai Learner cannot be produced without fastai)."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        torch.save({"arch": self.arch, "n_in": self.n_in, "n_classes": self.n_classes, "size": self.size,
                    "batch_size": self.bs, "class_weights": self.class_weights, "self_attention": self.self_attention,
                    "state_dict": {k: v.cpu() for k, v in self.state_dict().items()}}, path)


def unet_learner_MS(n_in: int, n_classes: int, arch="xresnet34", size: Tuple[int, int] = (256, 256),
                    batch_size: int = 4, pretrained=None, loss_func=None, class_weights=None, opt_func: str = "adam",
                    lr: float = 1e-3, wd: float = 0.01, encoder_factor: float = 10.0, moms=(0.95, 0.85, 0.95),
                    regression: bool = False, self_attention: bool = False) -> Learner:
    """train.py:98-160.  `n_in` / `size` replace what the reference probes from `dls.train_ds` (:124-125) and
    `n_classes` replaces `len(dls.vocab)` (:140); `pretrained` may be a state_dict with fastai keys."""
    if regression:
        raise NotImplementedError("the regression variant (MSELossFlat, n_out=1) is outside the built hot path")
    if loss_func is not None and not isinstance(loss_func, str):
        warnings.warn("loss_func objects are ignored: the plan implements CrossEntropyLossFlat(axis=1) with class weights")
    learn = Learner(_arch_name(arch), n_in, n_classes, size, batch_size, class_weights, opt_func, lr, wd,
                    encoder_factor, moms, self_attention=self_attention)
    if isinstance(pretrained, dict):
        learn.load_state_dict(pretrained)
    return learn


def load_learner(path, batch_size: Optional[int] = None) -> Learner:
    """predict.py:161 / train.py:225 — loads what `Learner.export` wrote."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    learn = Learner(ck["arch"], ck["n_in"], ck["n_classes"], tuple(ck["size"]), batch_size or ck["batch_size"],
                    ck.get("class_weights"), self_attention=ck.get("self_attention", False))
    learn.load_state_dict(ck["state_dict"])
    return learn


def _load_tile(path: Path):
    """(uint8 [n_in,H,W] tile, GeoInfo or None).  `.tif` tiles carry their own georeferencing (predict.py:206-215)."""
    if path.suffix.lower() in (".tif", ".tiff"):
        arr, geo = read_geotiff(path)
        if arr.dtype != np.uint8:
            raise ValueError(f"{path}: {arr.dtype} tiles are not supported by the uint8 fast path")
        return arr, geo
    return np.load(path), None


def _store(path: Path, arr: np.ndarray, geo: Optional[GeoInfo], nodata, class_zero: bool) -> Path:
    """`store_tif` (predict.py:19-52) for `.tif` tiles; `.npy` otherwise (same class_zero un-shift)."""
    if geo is not None:
        write_geotiff(path, arr, geo, nodata=nodata, class_zero=class_zero)
        return path
    if class_zero and arr.dtype.kind in "ui":
        arr = np.where(arr == 0, 0 if nodata is None else nodata, arr - 1).astype(arr.dtype)
    path = path.with_suffix(".npy")
    np.save(path, arr)
    return path


def save_predictions(predict_model, predict_path, regression: bool = False, merge: bool = False,
                     all_classes: bool = False, specific_class: Optional[int] = None, large_file: bool = False,
                     AOI=None, year=None, validation_vision: bool = False, class_zero: bool = False,
                     geotransforms: Optional[Dict[str, Sequence[float]]] = None):
    """predict.py:146-355 over the tiles in `predict_path`: GeoTIFF tiles (`*.tif`, `[n_in,H,W]` uint8 - georeferencing
    is read from every tile as the reference does) or `.npy` tiles (then `geotransforms[name] = (ulx, xres, xskew, uly,
    yskew, yres)` must be given for `merge`).  Without `merge` one prediction per tile goes to
    `../predicted_tiles_<model>/` (argmax Byte, `specific_class` or `all_classes` probabilities Float32 - or int8 x31
    with `large_file`, predict.py:226-254); with `merge` the tiles are placed by their geotransform with the same
    python `round()` arithmetic as predict.py:294-297, overlap-averaged and arg-maxed on the device, and ONE
    `<AOI>_<year>_<model>_prediction.tif` is written next to the tile folder (`all_classes` / `specific_class` write the
    averaged probabilities instead; `large_file` uses the int8 x31 / floor-division merge).  Returns the output path(s)."""
    if regression:
        raise NotImplementedError("regression prediction is outside the built hot path")
    if validation_vision:
        warnings.warn("validation_vision (confusion-matrix plots, predict.py:56-143) is outside the built path; skipped")
    learn = predict_model if isinstance(predict_model, Learner) else load_learner(predict_model)
    path = Path(predict_path)
    model_name = "model" if isinstance(predict_model, Learner) else os.path.basename(str(predict_model)).split(".")[0]
    tiles = sorted(path.glob("*.tif")) or sorted(path.glob("*.npy"))
    if not tiles:
        raise FileNotFoundError(f"no .tif / .npy tiles in {path}")
    out_dir = path.parent if merge else path.parent / ("predicted_tiles_" + model_name)
    out_dir.mkdir(parents=True, exist_ok=True)
    net = learn._eval_net()
    pred = TiledPredictor(net)
    B, P = net.N, net.H
    dev, lib = net.device, net.lib

    def batches():
        for b0 in range(0, len(tiles), B):
            chunk = tiles[b0:b0 + B]
            loaded = [_load_tile(t) for t in chunk]
            x = torch.from_numpy(np.stack([a for a, _ in loaded]))
            if tuple(x.shape[1:]) != (net.n_in, P, P):
                raise ValueError(f"tiles must be [{net.n_in},{P},{P}], got {tuple(x.shape[1:])}")
            yield b0, chunk, [g for _, g in loaded], x

    if merge:
        # ---- pass 1 over the headers: extent of the mosaic (predict.py:257-276)
        geos: List[Optional[GeoInfo]] = []
        gts = []
        for t in tiles:
            if t.suffix.lower() in (".tif", ".tiff"):
                nb, h, w, _, g = geotiff_info(t)
                geos.append(g)
                gt = g.geotransform
            else:
                if geotransforms is None:
                    raise ValueError("merge=True over .npy tiles needs the tiles' geotransforms")
                geos.append(None)
                gt, h, w = geotransforms[t.name], P, P
            gts.append([gt[0], w, gt[1], gt[3], h, gt[5]])
        gts = np.array(gts, dtype=np.float64)
        if geos[0] is not None and any(not geos[0].same_projection(g) for g in geos[1:]):
            warnings.warn("Geoprojection is not the same for all prediction tiles.")          # predict.py:211-212
        if len(set(gts[:, 1])) != 1 or len(set(gts[:, 4])) != 1:
            warnings.warn("Not all tiles have the same resolution.")                           # predict.py:272-273
        ulx_full, uly_full = gts[:, 0].min(), gts[:, 3].max()
        xmax_r, ymin_r = int(np.argmax(gts[:, 0])), int(np.argmin(gts[:, 3]))
        x_len = round((gts[:, 0].max() + gts[xmax_r, 1] * gts[xmax_r, 2] - ulx_full) / gts[0, 2])
        y_len = round((gts[:, 3].min() + gts[ymin_r, 4] * gts[ymin_r, 5] - uly_full) / gts[0, 5])
        place = [placement_from_geotransform(g[0], P, g[2], g[3], P, g[5], ulx_full, uly_full) for g in gts]
        acc = torch.zeros((net.n_out, y_len, x_len), dtype=torch.float32, device=dev)
        cnt = torch.zeros((y_len, x_len), dtype=torch.uint8, device=dev)
        mask = torch.empty((y_len, x_len), dtype=torch.uint8, device=dev)
        accumulate = lib.b2u_stitch_accumulate_q31 if large_file else lib.b2u_stitch_accumulate
        s_ = ops.stream_ptr()
        for b0, chunk, _, x in batches():
            n = len(chunk)
            if n < B:
                x = torch.cat([x, x[:1].expand(B - n, -1, -1, -1)], 0)
            net.set_input(x.to(dev).contiguous())
            net.forward()
            wins = [(place[b0 + i][0], place[b0 + i][1], P, P) for i in range(n)]
            y0 = torch.tensor([w_[1] for w_ in wins] + [0] * (B - n), dtype=torch.int32, device=dev)
            x0 = torch.tensor([w_[0] for w_ in wins] + [0] * (B - n), dtype=torch.int32, device=dev)
            for cls in colour_classes(wins):
                sel = torch.tensor(cls, dtype=torch.int32, device=dev)
                _lib.check(accumulate(net.logits.data_ptr(), net.logits.shape[-1], net.n_out, B, P, P, y0.data_ptr(),
                                      x0.data_ptr(), sel.data_ptr(), len(cls), acc.data_ptr(), cnt.data_ptr(), y_len,
                                      x_len, 0, 0, s_), "b2u_stitch_accumulate")
            torch.cuda.synchronize()
        if all_classes or specific_class is not None:
            # averaged probabilities: the reference's own host arithmetic on the device sums (predict.py:318-337)
            m = acc.cpu().numpy()
            c8 = cnt.cpu().numpy().astype(np.int8)
            if large_file:
                m = m.astype(np.int8)
                mk = np.broadcast_to(c8 > 0, m.shape)
                m[mk] //= np.broadcast_to(c8, m.shape)[mk]
            else:
                mk = np.broadcast_to(c8 > 0, m.shape)
                m[mk] /= np.broadcast_to(c8, m.shape)[mk]
            out = m if all_classes else m[specific_class]
        else:
            finalize = lib.b2u_stitch_finalize_q31 if large_file else lib.b2u_stitch_finalize
            _lib.check(finalize(acc.data_ptr(), cnt.data_ptr(), net.n_out, y_len, x_len, mask.data_ptr(), s_),
                       "b2u_stitch_finalize")
            out = mask.cpu().numpy()
        name = "_".join([p_ for p_ in (AOI, year, model_name, "prediction") if p_])
        geo = None
        if geos[0] is not None:
            g0 = geos[0]
            geo = GeoInfo((float(ulx_full), float(gts[0, 2]), 0.0, float(uly_full), 0.0, float(gts[0, 5])), g0.geokeys,
                          g0.geodoubles, g0.geoascii, None, True)                                # predict.py:350-352
        return _store(out_dir / (name + ".tif"), out, geo, None, class_zero)

    outs = []
    for b0, chunk, geos, x in batches():
        probs, amax = pred.predict_tiles(x.to(dev).contiguous())
        probs, amax = probs.cpu().numpy(), amax.cpu().numpy()
        for i, t in enumerate(chunk):
            if all_classes:
                arr = probs[i]
            elif specific_class is None:
                arr = amax[i]                                           # decoded argmax (predict.py:232)
            else:
                arr = probs[i][specific_class]
            if large_file and (all_classes or specific_class):          # predict.py:245-249 (sic: class 0 is falsy there)
                arr = np.around(arr * ((128 / 4) - 1)).astype(np.int8)
            outs.append(_store(out_dir / t.name, arr, geos[i], None, class_zero))
    return outs


def predict_geotiff(predict_model, raster_path, out_path=None, patch_overlap: float = 0.125, large_file: bool = False,
                    class_zero: bool = False, rank: int = 0, world: int = 1, max_strip_columns: Optional[int] = None):
    """Tile-free variant of the predict path: what `split_raster` (create_tiles_unet.py:252-431) + `save_predictions(merge=
    True)` produce together - the stitched argmax mask of a whole 4-band GeoTIFF - without writing tiles to disk: the
    raster is windowed on the device with `compute_windows` offsets, predicted and stitched in HBM.
    `max_strip_columns`: rasters that should not sit in host / device memory at once are streamed as vertical strips of
    at most that many output columns - each strip reads only the file window of the tile columns it needs
    (`read_geotiff(window=)`), is predicted as an owner-computes strip and lands in its columns of the mask, so the
    result is bit-identical to the one-shot prediction.  With `world > 1` every rank writes the column strip it owns
    (`<out>.part<rank>.tif`, georeferenced to its origin)."""
    learn = predict_model if isinstance(predict_model, Learner) else load_learner(predict_model)
    net = learn._eval_net()
    bands, Y, X, dt, geo = geotiff_info(raster_path)
    if dt != np.uint8:
        raise ValueError(f"{raster_path}: {dt} rasters are not supported by the uint8 fast path")
    if bands != net.n_in:
        raise ValueError(f"raster has {bands} bands, the model expects {net.n_in}")
    P = net.H
    pred = TiledPredictor(net)
    windows = compute_windows(Y, X, P, patch_overlap)
    _, xb, xe = shard_windows_by_columns(windows, X, rank, world)
    n_strips = 1 if not max_strip_columns else max(1, -(-(xe - xb) // int(max_strip_columns)))
    mask = np.empty((Y, xe - xb), dtype=np.uint8)
    for k in range(n_strips):
        # strip k of this rank's columns: the balanced partition of [xb, xe), then the tile columns intersecting it
        base, rem = divmod(xe - xb, n_strips)
        sb = xb + k * base + min(k, rem)
        se = sb + base + (1 if k < rem else 0)
        idx = [i for i, (x, y, w, h) in enumerate(windows) if x < se and x + w > sb]
        xs0 = min(windows[i][0] for i in idx)
        xs1 = max(windows[i][0] + windows[i][2] for i in idx)
        strip, _ = read_geotiff(raster_path, window=(xs0, 0, xs1 - xs0, Y))
        dev_s = torch.from_numpy(np.ascontiguousarray(strip)).pin_memory().to(net.device, non_blocking=True)
        # the strip holds exactly the tile columns of [sb, se): a one-rank prediction of it reproduces those tiles
        m, _, _ = pred.predict_raster(dev_s, patch_overlap, 0, 1, large_file=large_file)
        mask[:, sb - xb:se - xb] = m[:, sb - xs0:se - xs0].cpu().numpy()
    out_path = Path(out_path) if out_path is not None else Path(raster_path).with_name(Path(raster_path).stem + "_prediction.tif")
    if world > 1:
        out_path = out_path.with_suffix(f".part{rank}.tif")
    write_geotiff(out_path, mask, geo.window(xb, 0), nodata=None, class_zero=class_zero)
    return out_path


# ---------------------------------------------------------------------------------------------------- training entry
def _tile_batches(files: Sequence[Path], batch_size: int, n_classes: int, class_zero: bool, shuffle_seed: Optional[int],
                  drop_last: bool):
    """Batches of (uint8 [B,n_in,H,W], uint8 [B,H,W]) from `img_tiles/*.tif` + `mask_tiles/*.tif` (data.py:100-105,
    utils.py:40-55).  The last partial batch is padded by repetition only when `drop_last` is False."""
    def gen():
        order = list(range(len(files)))
        if shuffle_seed is not None:
            np.random.default_rng(shuffle_seed + gen.epoch).shuffle(order)
            gen.epoch += 1
        for b0 in range(0, len(order), batch_size):
            idx = order[b0:b0 + batch_size]
            if len(idx) < batch_size:
                if drop_last:
                    break
                idx = idx + [idx[0]] * (batch_size - len(idx))
            xs, ys = [], []
            for i in idx:
                x = open_tile(files[i])
                if x.dtype != np.uint8:
                    raise ValueError(f"{files[i]}: only uint8 tiles are supported by the training fast path")
                y = open_mask(files[i]).astype(np.int64)
                if y.max() >= n_classes:
                    raise ValueError(f"{files[i]}: mask label {int(y.max())} >= number of classes {n_classes}")
                xs.append(x)
                ys.append(y.astype(np.uint8))
            yield torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ys))
    gen.epoch = 0
    return gen


def train_func(data_path, existing_model, model_Path, description, BATCH_SIZE, visualize_data_example=False,
               enable_regression=False, CLASS_WEIGHTS="even", ARCHITECTURE="xresnet34", EPOCHS=1, LEARNING_RATE=1e-3,
               ENCODER_FACTOR=10, LR_FINDER=None, loss_func=None, monitor="dice_multi", self_attention=False,
               VALID_SCENES=("vali",), CODES=("background", "class1"), transforms=False, split_idx=None,
               export_model_summary=False, aug_pipe=None, n_transform_imgs=0, info=False, class_zero=False):
    """`train_func` (train.py:287-375) with the reference's positional signature, on the B200 plan.  Expects
    `data_path/{trai,vali}/{img_tiles,mask_tiles}/*.tif` (utils.py:25-36, data.py:102-105); trains with fastai's recipe
    (Adam, wd 0.01, discriminative lrs `slice(lr/ENCODER_FACTOR, lr)`, one-cycle) and writes
    `<model_Path>/<description>/<description>.pkl`, `.json` and `_history.csv` (train.py:314-320, 234, 255).
    Augmentation (`transforms`, `aug_pipe`), the LR finder, plots and the model summary are outside the built path and
    are refused or skipped with a warning.  Returns the trained Learner."""
    import json
    if enable_regression:
        raise NotImplementedError("the regression variant is outside the built hot path")
    if LR_FINDER:
        warnings.warn("LR_FINDER is outside the built path; training proceeds with LEARNING_RATE")
    if transforms or aug_pipe:
        warnings.warn("albumentations pipelines are outside the built path; training proceeds without augmentation")
    data_path = Path(data_path)
    valid_scenes = [VALID_SCENES] if isinstance(VALID_SCENES, str) else list(VALID_SCENES)
    train_files, valid_files = [], []
    if not data_path.exists():
        raise FileNotFoundError(data_path)                                         # train.py:54
    for scene in sorted(p for p in data_path.iterdir() if p.is_dir()):
        files = sorted((scene / "img_tiles").glob("*.tif"))
        (valid_files if scene.name in valid_scenes else train_files).extend(files)
    if not train_files:
        raise FileNotFoundError(f"no training tiles under {data_path}/*/img_tiles")
    nb, h, w, _, _ = geotiff_info(train_files[0])                                  # train.py:124-125 probes the first item
    codes = list(CODES)
    n_classes = len(codes)
    if isinstance(CLASS_WEIGHTS, str):
        if CLASS_WEIGHTS == "even":
            cw = [1.0 / n_classes] * n_classes                                     # train.py:338-339
        elif CLASS_WEIGHTS == "weighted":
            # utils.py:105-117 get_class_weights: total / count per class over (up to 1200) training masks.  The
            # reference takes the counts from `unique()`, which silently drops absent classes; here an absent class
            # keeps its slot (count clamped to 1).  The weighted-mean CE is invariant to the common scale.
            counts = np.zeros(n_classes, dtype=np.float64)
            for f in train_files[:1200]:
                counts += np.bincount(open_mask(f).ravel(), minlength=n_classes)[:n_classes]
            cw = list(counts.sum() / np.maximum(counts, 1.0))
        else:
            raise ValueError(f"CLASS_WEIGHTS {CLASS_WEIGHTS!r} not understood")
    else:
        cw = list(CLASS_WEIGHTS)
    learn = unet_learner_MS(nb, n_classes, ARCHITECTURE, (h, w), BATCH_SIZE, class_weights=cw, lr=LEARNING_RATE,
                            encoder_factor=ENCODER_FACTOR, self_attention=self_attention)
    if existing_model:
        if not os.path.exists(existing_model):
            raise FileNotFoundError(existing_model)
        learn.load_state_dict(torch.load(existing_model, map_location="cpu", weights_only=False)["state_dict"])
    out_dir = Path(model_Path) / description
    out_dir.mkdir(parents=True, exist_ok=True)
    tb = _tile_batches(train_files, BATCH_SIZE, n_classes, class_zero, shuffle_seed=0, drop_last=len(train_files) >= BATCH_SIZE)
    vb = _tile_batches(valid_files, BATCH_SIZE, n_classes, class_zero, None, False) if valid_files else None
    if monitor not in (None, "train_loss", "valid_loss", "r2_score", "dice_multi"):
        raise ValueError("Monitor must be one of ['train_loss', 'valid_loss', 'r2_score', 'dice_multi']")    # train.py:207-208
    mon = monitor if monitor in ("dice_multi", "valid_loss", "train_loss") else "dice_multi"                   # train.py:198-201
    if vb is None and mon != "train_loss":
        mon = "train_loss"
    learn.fit_one_cycle(EPOCHS, LEARNING_RATE, tb, vb, history_csv=str(out_dir / f"{description}_history.csv"),
                        monitor=mon, best_path=str(out_dir / "best-model.pth"))
    learn.export(out_dir / f"{description}.pkl")
    with open(out_dir / f"{description}.json", "w") as f:
        json.dump({"description": description, "architecture": learn.arch, "bands": nb, "tile": [h, w], "codes": codes,
                   "class_weights": cw, "batch_size": BATCH_SIZE, "epochs": EPOCHS, "learning_rate": LEARNING_RATE,
                   "encoder_factor": ENCODER_FACTOR, "train_tiles": len(train_files), "valid_tiles": len(valid_files),
                   "class_zero": bool(class_zero), "history": learn.history}, f, indent=1)
    return learn
