"""The reference-facing Python seams (unet_learner_MS / fit_one_cycle / predict / export / load_learner /
save_predictions) running on the B200 plan: training reduces the loss on a learnable synthetic task, the exported
checkpoint round-trips, and the merged prediction equals the tile-wise reference merge."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_learner_trains_predicts_and_round_trips(tmp_path):
    from unet_b200.reference_api import load_learner, save_predictions, unet_learner_MS
    from unet_b200.synth import aerial_like_tiles
    from unet_b200.tiling import compute_windows
    with pytest.raises(TypeError):
        unet_learner_MS(4, 2, arch="resnet34")                 # only xresnet bodies work in the reference too
    learn = unet_learner_MS(4, 2, arch="xresnet18", size=(64, 64), batch_size=8, lr=2e-3, opt_func="adam")
    x, y = aerial_like_tiles(32, 4, 64, 64, 2, seed=3)
    batches = lambda: [(x[i:i + 8], y[i:i + 8]) for i in range(0, 32, 8)]
    hist = learn.fit_one_cycle(4, 2e-3, batches, batches, history_csv=str(tmp_path / "history.csv"))
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]      # the Adam / one-cycle step actually learns
    assert 0.0 <= hist[-1]["dice_multi"] <= 1.0
    assert open(tmp_path / "history.csv").readline().strip() == "epoch,train_loss,valid_loss,dice_multi,time"
    dec, amax, probs = learn.predict(x[0])
    assert amax.shape == (64, 64) and probs.shape == (2, 64, 64)
    assert torch.allclose(probs.sum(0), torch.ones(64, 64), atol=1e-5)
    learn.export(tmp_path / "m" / "model.pkl")
    again = load_learner(tmp_path / "m" / "model.pkl")
    _, amax2, probs2 = again.predict(x[0])
    assert torch.equal(amax, amax2) and torch.equal(probs, probs2)

    # tiles on disk + geotransforms -> merged mask, against the numpy merge of the per-tile probabilities
    H, W, P = 150, 200, 64
    raster, _ = aerial_like_tiles(1, 4, H, W, 2, seed=9)
    raster = raster[0].numpy()
    wins = compute_windows(H, W, P, 0.25)
    tdir = tmp_path / "tiles"
    tdir.mkdir()
    gt0 = (500000.0, 0.5, 0.0, 6000000.0, 0.0, -0.5)
    gts = {}
    for i, (wx, wy, ww, wh) in enumerate(wins):
        np.save(tdir / f"img_{i}.npy", raster[:, wy:wy + wh, wx:wx + ww])
        gts[f"img_{i}.npy"] = (gt0[0] + wx * gt0[1], gt0[1], 0.0, gt0[3] + wy * gt0[5], 0.0, gt0[5])
    out = save_predictions(again, str(tdir), False, merge=True, AOI="aoi", year="2024", geotransforms=gts)
    merged = np.load(out)
    assert merged.shape == (H, W)
    from oracle.stitch import merge_pixel_windows
    names = sorted(p.name for p in tdir.glob("*.npy"))
    probs_list, win_list = [], []
    for n in names:
        i = int(n.split("_")[1].split(".")[0])
        _, _, pr = again.predict(torch.from_numpy(np.load(tdir / n)))
        probs_list.append(pr.numpy())
        win_list.append(wins[i])
    ref = merge_pixel_windows(probs_list, win_list, H, W)
    assert (merged == ref).mean() >= 0.9999
    outs = save_predictions(again, str(tdir), False, merge=False)
    assert len(outs) == len(wins) and np.load(outs[0]).shape == (P, P)


def _write_tiles(tdir, raster, wins, gt0, geo_keys):
    from unet_b200.geotiff import GeoInfo, write_geotiff
    tdir.mkdir(parents=True, exist_ok=True)
    base = GeoInfo(gt0, geo_keys, georeferenced=True)
    for i, (wx, wy, ww, wh) in enumerate(wins):
        write_geotiff(tdir / f"img_{i:03d}.tif", raster[:, wy:wy + wh, wx:wx + ww], base.window(wx, wy))


def test_geotiff_tiles_merge_modes_and_whole_raster(tmp_path):
    """save_predictions over GeoTIFF tiles (placement from each tile's own geotransform, predict.py:206-222, 292-305):
    merged argmax, `large_file` int8 merge (bit-exact against the numpy restatement fed with the SAME per-tile
    probabilities), averaged probabilities, per-tile outputs with georeferencing, and the tile-free `predict_geotiff`."""
    from oracle.stitch import merge_tiles
    from unet_b200.geotiff import read_geotiff, write_geotiff, GeoInfo
    from unet_b200.reference_api import predict_geotiff, save_predictions, unet_learner_MS
    from unet_b200.synth import aerial_like_tiles
    from unet_b200.tiling import compute_windows
    learn = unet_learner_MS(4, 3, arch="xresnet18", size=(64, 64), batch_size=8)
    learn.net.init_parameters(seed=1, randomize_bn=True)
    H, W, P, ov = 150, 200, 64, 0.25
    raster, _ = aerial_like_tiles(1, 4, H, W, 2, seed=11)
    raster = raster[0].numpy()
    wins = compute_windows(H, W, P, ov)
    gt0 = (383000.0, 0.2, 0.0, 5819000.0, 0.0, -0.2)
    keys = (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 25833)
    tdir = tmp_path / "aoi" / "tiles"
    _write_tiles(tdir, raster, wins, gt0, keys)
    names = sorted(p.name for p in tdir.glob("*.tif"))
    probs, gts = [], []
    for n in names:
        a, g = read_geotiff(tdir / n)
        probs.append(learn.predict(torch.from_numpy(a))[2].numpy())
        gt = g.geotransform
        gts.append([gt[0], P, gt[1], gt[3], P, gt[5]])

    out = save_predictions(learn, str(tdir), False, merge=True, AOI="aoi", year="2024")
    assert out.name == "aoi_2024_model_prediction.tif"
    merged, geo = read_geotiff(out)
    ref, origin = merge_tiles(probs, gts)
    assert merged.shape == (1, H, W) and merged.dtype == np.uint8
    assert (merged[0] == ref).mean() >= 0.9999
    assert geo.geotransform == (gt0[0], gt0[1], 0.0, gt0[3], 0.0, gt0[5]) and geo.geokeys == keys   # predict.py:350-352

    out8 = save_predictions(learn, str(tdir), False, merge=True, large_file=True, AOI="aoi8")
    ref8, _ = merge_tiles(probs, gts, large_file=True)
    m8, _ = read_geotiff(out8)
    assert np.array_equal(m8[0], ref8)           # integer arithmetic on identical probabilities: bit-exact

    outp = save_predictions(learn, str(tdir), False, merge=True, all_classes=True, AOI="aoip")
    mp, _ = read_geotiff(outp)
    refp, _ = merge_tiles(probs, gts, all_classes=True)
    assert mp.dtype == np.float32 and np.abs(mp - refp).max() <= 1e-5

    outs = save_predictions(learn, str(tdir), False, merge=False, class_zero=True)
    assert len(outs) == len(wins)
    t0, g0 = read_geotiff(outs[0])
    am = probs[0].argmax(0)
    assert np.array_equal(t0[0], np.where(am == 0, 0, am - 1))       # class_zero un-shift, predict.py:34-36
    assert g0.geotransform == read_geotiff(tdir / names[0])[1].geotransform

    # whole raster in, mask GeoTIFF out: same mask as the tile-wise merge
    rpath = tmp_path / "aoi" / "raster.tif"
    write_geotiff(rpath, raster, GeoInfo(gt0, keys, georeferenced=True))
    mpath = predict_geotiff(learn, rpath, patch_overlap=ov)
    mm, gm = read_geotiff(mpath)
    assert np.array_equal(mm[0], merged[0]) and gm.geotransform == geo.geotransform
    # streamed as vertical strips (window reads of the file): bit-identical to the one-shot prediction
    spath = predict_geotiff(learn, rpath, tmp_path / "aoi" / "streamed.tif", patch_overlap=ov, max_strip_columns=70)
    assert np.array_equal(read_geotiff(spath)[0], mm)


def test_train_func_on_geotiff_tiles(tmp_path):
    """train.py:287-375 entry point with the reference's positional arguments over data_path/{trai,vali}/..."""
    from unet_b200.geotiff import GeoInfo, write_geotiff
    from unet_b200.reference_api import load_learner, train_func
    from unet_b200.synth import aerial_like_tiles
    x, y = aerial_like_tiles(20, 4, 64, 64, 2, seed=5)
    for scene, sl in (("trai", slice(0, 16)), ("vali", slice(16, 20))):
        for sub in ("img_tiles", "mask_tiles"):
            (tmp_path / "data" / scene / sub).mkdir(parents=True)
        for i in range(sl.start, sl.stop):
            write_geotiff(tmp_path / "data" / scene / "img_tiles" / f"t{i}.tif", x[i].numpy(), GeoInfo())
            write_geotiff(tmp_path / "data" / scene / "mask_tiles" / f"t{i}.tif", y[i].numpy(), GeoInfo())
    with pytest.raises(FileNotFoundError):
        train_func(tmp_path / "nope", None, tmp_path / "models", "run", 8)
    learn = train_func(tmp_path / "data", None, tmp_path / "models", "run", 8, False, False, "even", "xresnet18", 3, 2e-3,
                       10, None, None, "dice_multi", False, "vali", ["background", "forest"])
    d = tmp_path / "models" / "run"
    assert (d / "run.pkl").exists() and (d / "run.json").exists()
    assert open(d / "run_history.csv").readline().strip() == "epoch,train_loss,valid_loss,dice_multi,time"
    # (train_loss is fastai's smoothed loss: with six optimizer steps it is dominated by the warm-up spike of the fresh
    # BatchZero network, so learning is read off the validation loss)
    assert len(learn.history) == 3 and learn.history[-1]["valid_loss"] < learn.history[0]["valid_loss"]
    assert all(np.isfinite(r["train_loss"]) for r in learn.history)
    again = load_learner(d / "run.pkl")
    assert torch.equal(again.predict(x[0])[1], learn.predict(x[0])[1])


def test_learner_with_self_attention(tmp_path):
    """params_and_main.py:83 default self_attention=True through the reference-facing API: the CUDA-graph training step
    (attention products captured with the rest of the plan) learns, gamma leaves its zero init, export / load round-trips."""
    from unet_b200.reference_api import load_learner, unet_learner_MS
    from unet_b200.synth import aerial_like_tiles
    learn = unet_learner_MS(4, 2, arch="xresnet18", size=(64, 64), batch_size=8, lr=2e-3, self_attention=True)
    assert float(learn.net.param("layers.5.conv2.2.gamma")) == 0.0          # fastai init: the block starts as identity
    x, y = aerial_like_tiles(32, 4, 64, 64, 2, seed=3)
    batches = lambda: [(x[i:i + 8], y[i:i + 8]) for i in range(0, 32, 8)]
    hist = learn.fit_one_cycle(4, 2e-3, batches, batches)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"] and all(np.isfinite(h["train_loss"]) for h in hist)
    assert float(learn.net.param("layers.5.conv2.2.gamma")) != 0.0
    learn.export(tmp_path / "sa.pkl")
    again = load_learner(tmp_path / "sa.pkl")
    assert again.self_attention and "layers.5.conv2.2.query.0.weight_u" in again.state_dict()
    _, a1, p1 = learn.predict(x[0])
    _, a2, p2 = again.predict(x[0])
    assert torch.equal(a1, a2) and torch.equal(p1, p2)


def test_fastai_initial_state():
    """What `unet_learner_MS` hands to the first optimizer step: BatchNorm gamma 0 on the last BatchNorm of
    every ResBlock convpath (NormType.BatchZero - the oracle's `zero_bn` pattern) and 1 elsewhere, beta 1e-3, running
    statistics (0, 1); decoder biases N(0, 0.01) except the middle conv and the final ResBlock (apply_init: 0); ICNR on
    the PixelShuffle convolutions; attention gamma 0."""
    from oracle.unet_oracle import DynamicUnetOracle
    from unet_b200.reference_api import unet_learner_MS
    learn = unet_learner_MS(4, 2, arch="xresnet34", size=(64, 64), batch_size=2)
    sd = learn.state_dict()
    ref = DynamicUnetOracle("xresnet34", 4, 2).state_dict()       # constructed, not randomised: fastai's BN pattern
    n_zero = 0
    for k, v in ref.items():
        if v.dim() == 1 and (k.endswith(".1.weight") or k.endswith(".1.bias") or ".bn." in k or k.startswith("layers.1.")) \
                and not k.endswith("num_batches_tracked"):
            assert torch.equal(sd[k].cpu(), v), k
            n_zero += int(k.endswith(".weight") and float(v.abs().max()) == 0.0)
    assert n_zero == 16                                          # one BatchZero per ResBlock of xresnet34
    for k in ("layers.3.0.0.bias", "layers.3.1.0.bias", "layers.11.convpath.0.0.bias", "layers.11.convpath.1.0.bias"):
        assert float(sd[k].abs().max()) == 0.0
    b = torch.cat([sd[f"layers.{i}.conv{j}.0.bias"].flatten() for i in (4, 5, 6, 7) for j in (1, 2)])
    assert 0.008 < float(b.std()) < 0.012 and abs(float(b.mean())) < 2e-3
    w = sd["layers.4.shuf.0.0.weight"]
    assert torch.equal(w[0::4], w[1::4]) and torch.equal(w[0::4], w[3::4])       # icnr_init


def test_validate_on_device_matches_torch_and_ignores_padding():
    """Learner.validate: AvgLoss weighted by the REAL samples of every batch and fastai DiceMulti, reduced on the device -
    against torch on the plan's own logits (restricted to the real samples) and the oracle's dice_multi."""
    import torch.nn.functional as F
    from oracle.unet_oracle import dice_multi
    from unet_b200.reference_api import unet_learner_MS
    from unet_b200.synth import aerial_like_tiles
    C_ = 3
    learn = unet_learner_MS(4, C_, arch="xresnet18", size=(64, 64), batch_size=8, class_weights=[0.2, 0.5, 0.3])
    learn.net.init_parameters(seed=4, randomize_bn=True)
    learn._version += 1
    x, y = aerial_like_tiles(13, 4, 64, 64, C_, seed=2)
    # 13 samples in batches of 8: the second batch holds 5 real samples padded with copies of its first one
    pad = lambda t: torch.cat([t[8:], t[8:9].expand(3, *t.shape[1:])], 0)
    batches = [(x[:8], y[:8], 8), (pad(x), pad(y), 5)]
    got = learn.validate(batches)
    net = learn._eval_net()
    w = torch.tensor([0.2, 0.5, 0.3], device="cuda")
    losses, preds = [], []
    for xb, yb, n in batches:
        net.set_input(xb.cuda().contiguous())
        net.forward()
        lg = net.logits_nchw()[:n]
        losses.append((F.cross_entropy(lg, yb[:n].cuda().long(), weight=w).item(), n))
        preds.append(lg.argmax(1).cpu())
    want_loss = sum(l * n for l, n in losses) / 13
    assert abs(got["valid_loss"] - want_loss) <= 1e-5 * abs(want_loss)
    assert abs(got["dice_multi"] - dice_multi(torch.cat(preds), y.long(), C_)) <= 1e-12
    # counting the padding would change both numbers
    biased = dice_multi(torch.cat([preds[0], learn_pred_all(net, batches[1][0])]), torch.cat([y[:8], batches[1][1]]).long(), C_)
    assert abs(biased - got["dice_multi"]) > 1e-6


def learn_pred_all(net, xb):
    net.set_input(xb.cuda().contiguous())
    net.forward()
    return net.logits_nchw().argmax(1).cpu()


def test_regression_variant(tmp_path):
    """enable_regression (train.py:87-95, 137-138, 189-192; predict.py:196-198, 307-316): n_out 1, MSELossFlat, rmse / R2Score,
    Learner_adjust.predict, the regression merge with nodata -9999 - through train_func / save_predictions over GeoTIFF tiles."""
    from oracle.stitch import merge_tiles
    from unet_b200.geotiff import GeoInfo, read_geotiff, write_geotiff
    from unet_b200.reference_api import load_learner, save_predictions, train_func, unet_learner_MS
    from unet_b200.synth import aerial_like_tiles
    from unet_b200.tiling import compute_windows
    x, _ = aerial_like_tiles(24, 4, 64, 64, 2, seed=6)
    target = (x[:, 3].float() - x[:, 0].float()) / 64.0 + 1.0          # a continuous "index" of the bands, O(1) values
    for scene, sl in (("trai", slice(0, 16)), ("vali", slice(16, 24))):
        for sub in ("img_tiles", "mask_tiles"):
            (tmp_path / "data" / scene / sub).mkdir(parents=True)
        for i in range(sl.start, sl.stop):
            write_geotiff(tmp_path / "data" / scene / "img_tiles" / f"t{i}.tif", x[i].numpy(), GeoInfo())
            write_geotiff(tmp_path / "data" / scene / "mask_tiles" / f"t{i}.tif", target[i].numpy(), GeoInfo())
    learn = train_func(tmp_path / "data", None, tmp_path / "models", "reg", 8, False, True, "even", "xresnet18", 10, 3e-3,
                       10, None, None, None, False, "vali", ["value"])
    assert learn.regression and learn.net.n_out == 1 and learn.input_div == (1.0, 1.0)     # no IntToFloatTensor (data.py:98)
    d = tmp_path / "models" / "reg"
    assert open(d / "reg_history.csv").readline().strip() == "epoch,train_loss,valid_loss,_rmse,r2_score,time"
    h = learn.history
    # raw 0..255 band values enter the network (the reference's regression DataBlock has no IntToFloatTensor), so the
    # first steps are rough; the best epoch (which SaveModelCallback keeps) must beat the first one
    assert min(r["valid_loss"] for r in h[1:]) < h[0]["valid_loss"] and all(np.isfinite(r["valid_loss"]) for r in h)
    assert all(abs(r["_rmse"] ** 2 - r["valid_loss"]) < 1e-3 * r["valid_loss"] + 1e-6 for r in h)
    # metrics against torch on the plan's own predictions
    net = learn._eval_net()
    preds = []
    for i in range(16, 24, 8):
        net.set_input(x[i:i + 8].cuda().contiguous())
        net.forward()
        preds.append(net.logits_nchw()[:, 0].cpu())
    p_, t_ = torch.cat(preds).double().flatten(), target[16:24].double().flatten()
    r2 = 1 - ((p_ - t_) ** 2).sum() / ((t_ - t_.mean()) ** 2).sum()
    again = load_learner(d / "reg.pkl")
    m = again.validate([(x[16:24], target[16:24], 8)])
    assert abs(m["r2_score"] - r2.item()) < 1e-6 and abs(m["_rmse"] - ((p_ - t_) ** 2).mean().sqrt().item()) < 1e-6
    val, val2 = again.predict(x[16])
    assert val.shape == (1, 64, 64) and torch.equal(val, val2)
    with pytest.raises(ValueError):
        save_predictions(again, str(tmp_path), False, merge=True)           # classification call on a regression model
    # regression merge over georeferenced tiles with a gap (nodata)
    H, W, P = 150, 200, 64
    raster, _ = aerial_like_tiles(1, 4, H, W, 2, seed=12)
    raster = raster[0].numpy()
    wins = [w_ for w_ in compute_windows(H, W, P, 0.25) if not (w_[0] == 48 and w_[1] == 48)]
    gt0 = (383000.0, 0.2, 0.0, 5819000.0, 0.0, -0.2)
    tdir = tmp_path / "aoi" / "tiles"
    _write_tiles(tdir, raster, wins, gt0, (1, 1, 0, 0))
    names = sorted(p.name for p in tdir.glob("*.tif"))
    vals, gts = [], []
    for n in names:
        a, g = read_geotiff(tdir / n)
        vals.append(again.predict(torch.from_numpy(a))[1].numpy())
        gt = g.geotransform
        gts.append([gt[0], P, gt[1], gt[3], P, gt[5]])
    out = save_predictions(again, str(tdir), True, merge=True, AOI="aoi")
    got, geo = read_geotiff(out)
    ref, _ = merge_tiles(vals, gts, regression=True)
    assert got.dtype == np.float32 and geo.nodata == -9999
    assert np.array_equal(got[0] == -9999, ref == -9999)
    assert np.allclose(got[0], ref, rtol=1e-5, atol=1e-5)
    outs = save_predictions(again, str(tdir), True, merge=False)
    assert np.allclose(read_geotiff(outs[0])[0], vals[0], atol=1e-6)


def test_sixteen_bit_tiles_through_the_api(tmp_path):
    """A0 for 16-bit imagery through train_func / predict: uint16 tiles whose values exceed 8 bits make the dataset the
    reference's 'int16' kind (utils.py:72-89) - the raw values are divided by 255 twice on the device; the plan's logits
    for raw uint16 input equal its logits for the same values pre-scaled in fp32."""
    from unet_b200.geotiff import GeoInfo, write_geotiff
    from unet_b200.reference_api import load_learner, train_func
    from unet_b200.synth import aerial_like_tiles
    x8, y = aerial_like_tiles(12, 4, 64, 64, 2, seed=8)
    x16 = (x8.to(torch.int32) * 200 + 17).numpy().astype(np.uint16)            # up to 51 017
    for scene, sl in (("trai", slice(0, 8)), ("vali", slice(8, 12))):
        for sub in ("img_tiles", "mask_tiles"):
            (tmp_path / "data" / scene / sub).mkdir(parents=True)
        for i in range(sl.start, sl.stop):
            write_geotiff(tmp_path / "data" / scene / "img_tiles" / f"t{i}.tif", x16[i], GeoInfo())
            write_geotiff(tmp_path / "data" / scene / "mask_tiles" / f"t{i}.tif", y[i].numpy(), GeoInfo())
    learn = train_func(tmp_path / "data", None, tmp_path / "models", "u16", 4, False, False, "even", "xresnet18", 2, 2e-3,
                       10, None, None, "dice_multi", False, "vali", ["a", "b"])
    assert learn.sixteen_bit and learn.input_dtype == torch.uint16 and learn.input_div == (255.0, 255.0)
    assert all(np.isfinite(r["train_loss"]) for r in learn.history)
    again = load_learner(tmp_path / "models" / "u16" / "u16.pkl")
    assert again.sixteen_bit and again.input_dtype == torch.uint16
    net = again._eval_net()
    xb = torch.from_numpy(x16[:4])
    net.set_input(xb.cuda().contiguous())
    net.forward()
    a = net.logits_nchw().clone()
    net.set_input(torch.from_numpy(x16[:4].astype(np.int32)).cuda().float().div_(255).div_(255).contiguous())
    net.forward()
    assert torch.equal(a, net.logits_nchw())
    _, amax, probs = again.predict(xb[0])
    assert amax.shape == (64, 64) and torch.allclose(probs.sum(0), torch.ones(64, 64), atol=1e-5)
