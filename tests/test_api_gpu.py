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
