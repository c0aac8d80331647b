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
    assert len(learn.history) == 3 and learn.history[-1]["train_loss"] < learn.history[0]["train_loss"]
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
