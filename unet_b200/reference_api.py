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
    predict.save_predictions(...)          predict.py:146   save_predictions(...) over in-memory tiles / .npy tiles
    create_tiles_unet.compute_windows      :30              unet_b200.tiling.compute_windows

Data enters as uint8 tiles `[B, n_in, H, W]` (the reference reads GeoTIFF bands as int32 -> float32 and fastai divides by
255, data.py:24 + IntToFloatTensor; both happen inside the cast kernel) and class-id masks `[B, H, W]`.
GeoTIFF file I/O itself is out of the measured path; tiles on disk are read from `.npy`.
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
from .tiling import compute_windows, placement_from_geotransform

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
                 wd: float = 0.01, encoder_factor: float = 10.0, moms: Sequence[float] = (0.95, 0.85, 0.95)):
        self.arch, self.n_in, self.n_classes, self.size, self.bs = arch, n_in, n_classes, tuple(size), batch_size
        self.class_weights = list(class_weights) if class_weights is not None else None
        self.net = UNetB200(arch, n_in, n_classes, self.size, batch_size, training=True,
                            class_weights=self.class_weights)
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
            for x, y in train_batches():
                lr, mom = one_cycle(step / total, lr_max, moms=self.moms)
                if self.trainer.optimizer == "adam":
                    self.trainer.set_adam_hyper(lr, mom)
                else:
                    self.trainer.lr = lr
                loss = self.trainer.step(x, y)
                run += float(loss.item())
                n += 1
                step += 1
            row = {"epoch": ep, "train_loss": run / max(1, n)}
            if valid_batches is not None:
                row["valid_loss"], row["dice_multi"] = self.validate(valid_batches())
            row["time"] = time.time() - t0
            self.history.append(row)
            score = row.get(monitor)
            if best_path and score is not None and (best is None or (score < best if "loss" in monitor else score > best)):
                best = score
                torch.save(self.state_dict(), best_path)
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
                                  class_weights=self.class_weights)
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
                    "batch_size": self.bs, "class_weights": self.class_weights,
                    "state_dict": {k: v.cpu() for k, v in self.state_dict().items()}}, path)


def unet_learner_MS(n_in: int, n_classes: int, arch="xresnet34", size: Tuple[int, int] = (256, 256),
                    batch_size: int = 4, pretrained=None, loss_func=None, class_weights=None, opt_func: str = "adam",
                    lr: float = 1e-3, wd: float = 0.01, encoder_factor: float = 10.0, moms=(0.95, 0.85, 0.95),
                    regression: bool = False, self_attention: bool = False) -> Learner:
    """train.py:98-160.  `n_in` / `size` replace what the reference probes from `dls.train_ds` (:124-125) and
    `n_classes` replaces `len(dls.vocab)` (:140); `pretrained` may be a state_dict with fastai keys."""
    if regression:
        raise NotImplementedError("the regression variant (MSELossFlat, n_out=1) is outside the built hot path")
    if self_attention:
        raise NotImplementedError("SelfAttention on UnetBlock #1 is not built yet (SURVEY.md 8(f) rank 2)")
    if loss_func is not None and not isinstance(loss_func, str):
        warnings.warn("loss_func objects are ignored: the plan implements CrossEntropyLossFlat(axis=1) with class weights")
    learn = Learner(_arch_name(arch), n_in, n_classes, size, batch_size, class_weights, opt_func, lr, wd,
                    encoder_factor, moms)
    if isinstance(pretrained, dict):
        learn.load_state_dict(pretrained)
    return learn


def load_learner(path, batch_size: Optional[int] = None) -> Learner:
    """predict.py:161 / train.py:225 — loads what `Learner.export` wrote."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    learn = Learner(ck["arch"], ck["n_in"], ck["n_classes"], tuple(ck["size"]), batch_size or ck["batch_size"],
                    ck.get("class_weights"))
    learn.load_state_dict(ck["state_dict"])
    return learn


def save_predictions(predict_model, predict_path, regression: bool = False, merge: bool = False,
                     all_classes: bool = False, specific_class: Optional[int] = None, large_file: bool = False,
                     AOI=None, year=None, validation_vision: bool = False, class_zero: bool = False,
                     geotransforms: Optional[Dict[str, Sequence[float]]] = None):
    """predict.py:146-355 over `.npy` tiles ([n_in,H,W] uint8) in `predict_path`.  Without `merge` one `<tile>.npy`
    prediction per tile is written to `../predicted_tiles_<model>/` (argmax uint8, `specific_class` probabilities or all
    probabilities); with `merge` the tiles are placed by their geotransform (`geotransforms[name] = (ulx, xres, xskew,
    uly, yskew, yres)`, GDAL order) exactly as predict.py:294-297 does, averaged over overlaps and arg-maxed, and ONE
    `<AOI>_<year>_<model>_prediction.npy` is written next to the tile folder.  Returns the output path(s)."""
    if regression:
        raise NotImplementedError("regression prediction is outside the built hot path")
    if large_file:
        raise NotImplementedError("int8 'large_file' accumulation is not built yet (SURVEY.md 8(f) rank 4)")
    learn = predict_model if isinstance(predict_model, Learner) else load_learner(predict_model)
    path = Path(predict_path)
    model_name = "model" if isinstance(predict_model, Learner) else os.path.basename(str(predict_model)).split(".")[0]
    tiles = sorted(path.glob("*.npy"))
    if not tiles:
        raise FileNotFoundError(f"no .npy tiles in {path}")
    out_dir = path.parent if merge else path.parent / ("predicted_tiles_" + model_name)
    out_dir.mkdir(parents=True, exist_ok=True)
    net = learn._eval_net()
    pred = TiledPredictor(net)
    B = net.N
    if merge:
        # predict.py:257-337 on the device: placement from the geotransforms (same python round() arithmetic as
        # predict.py:294-297), softmax + overlap accumulate + normalise + argmax in the stitch kernels
        if geotransforms is None:
            raise ValueError("merge=True needs the tiles' geotransforms")
        if all_classes or specific_class is not None:
            raise NotImplementedError("merged probability outputs are not built yet (SURVEY.md 8(f) rank 4); argmax only")
        names = [t.name for t in tiles]
        P = net.H
        gts = np.array([[geotransforms[n][0], P, geotransforms[n][1], geotransforms[n][3], P, geotransforms[n][5]]
                        for n in names], dtype=np.float64)
        ulx_full, uly_full = gts[:, 0].min(), gts[:, 3].max()
        xmax_r, ymin_r = int(np.argmax(gts[:, 0])), int(np.argmin(gts[:, 3]))
        x_len = round((gts[:, 0].max() + gts[xmax_r, 1] * gts[xmax_r, 2] - ulx_full) / gts[0, 2])
        y_len = round((gts[:, 3].min() + gts[ymin_r, 4] * gts[ymin_r, 5] - uly_full) / gts[0, 5])
        place = [placement_from_geotransform(g[0], P, g[2], g[3], P, g[5], ulx_full, uly_full) for g in gts]
        dev, lib = net.device, net.lib
        acc = torch.zeros((net.n_out, y_len, x_len), dtype=torch.float32, device=dev)
        cnt = torch.zeros((y_len, x_len), dtype=torch.uint8, device=dev)
        mask = torch.empty((y_len, x_len), dtype=torch.uint8, device=dev)
        from .tiling import colour_classes
        s_ = ops.stream_ptr()
        for b0 in range(0, len(tiles), B):
            chunk = tiles[b0:b0 + B]
            n = len(chunk)
            x = torch.from_numpy(np.stack([np.load(t) for t in chunk]))
            if n < B:
                x = torch.cat([x, x[:1].expand(B - n, -1, -1, -1)], 0)
            net.set_input(x.to(dev).contiguous())
            net.forward()
            wins = [(place[b0 + i][0], place[b0 + i][1], P, P) for i in range(n)]
            y0 = torch.tensor([w[1] for w in wins] + [0] * (B - n), dtype=torch.int32, device=dev)
            x0 = torch.tensor([w[0] for w in wins] + [0] * (B - n), dtype=torch.int32, device=dev)
            for cls in colour_classes(wins):
                sel = torch.tensor(cls, dtype=torch.int32, device=dev)
                _lib.check(lib.b2u_stitch_accumulate(net.logits.data_ptr(), net.logits.shape[-1], net.n_out, B, P, P,
                                                     y0.data_ptr(), x0.data_ptr(), sel.data_ptr(), len(cls),
                                                     acc.data_ptr(), cnt.data_ptr(), y_len, x_len, 0, 0, s_),
                           "b2u_stitch_accumulate")
            torch.cuda.synchronize()
        _lib.check(lib.b2u_stitch_finalize(acc.data_ptr(), cnt.data_ptr(), net.n_out, y_len, x_len, mask.data_ptr(), s_),
                   "b2u_stitch_finalize")
        name = "_".join([p_ for p_ in (AOI, year, model_name, "prediction") if p_]) + ".npy"
        out = mask.cpu().numpy()
        if class_zero:
            out = out + 1
        np.save(out_dir / name, out)
        return out_dir / name
    results, names = [], []
    for b0 in range(0, len(tiles), B):
        chunk = tiles[b0:b0 + B]
        n = len(chunk)
        x = torch.from_numpy(np.stack([np.load(t) for t in chunk])).to(net.device).contiguous()
        probs, amax = pred.predict_tiles(x)
        for i, t in enumerate(chunk):
            names.append(t.name)
            results.append((probs[i].cpu().numpy(), amax[i].cpu().numpy()))
    outs = []
    for name, (pr, am) in zip(names, results):
        arr = pr if all_classes else (am if specific_class is None else pr[specific_class])
        if class_zero and arr.dtype == np.uint8:
            arr = arr + 1          # store_tif un-shifts the class_zero label shift (predict.py:19-52)
        np.save(out_dir / name, arr)
        outs.append(out_dir / name)
    return outs
