"""GeoTIFF reader / writer (unet_b200/geotiff.py) against Pillow's independent TIFF codec and round trips.
The reference reaches these files through rasterio / GDAL (data.py:18-28, utils.py:40-55, predict.py:19-52)."""
import numpy as np
import pytest

from unet_b200.geotiff import GeoInfo, geotiff_info, open_mask, open_tile, read_geotiff, write_geotiff

PIL = pytest.importorskip("PIL.Image")

GEO = GeoInfo(geotransform=(383000.5, 0.2, 0.0, 5819000.25, 0.0, -0.2),
              geokeys=(1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 25833), georeferenced=True)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
@pytest.mark.parametrize("bands", [1, 4])
def test_round_trip_and_window(tmp_path, dtype, bands):
    rng = np.random.default_rng(0)
    a = (rng.random((bands, 70, 93)) * 200).astype(dtype)
    p = tmp_path / "t.tif"
    write_geotiff(p, a, GEO, nodata=-9999 if dtype == np.float32 else None)
    b, g = read_geotiff(p)
    assert b.dtype == a.dtype and np.array_equal(a, b)
    assert g.geotransform == GEO.geotransform and g.geokeys == GEO.geokeys and g.same_projection(GEO)
    assert g.nodata == (-9999 if dtype == np.float32 else None)
    nb, h, w, dt, g2 = geotiff_info(p)
    assert (nb, h, w, dt) == (bands, 70, 93, np.dtype(dtype)) and g2.geotransform == GEO.geotransform
    win, gw = read_geotiff(p, window=(10, 20, 33, 41))
    assert np.array_equal(win, a[:, 20:61, 10:43])
    # upper-left corner of the window in map units (create_tiles_unet.py:224-226)
    assert gw.geotransform[0] == pytest.approx(383000.5 + 10 * 0.2) and gw.geotransform[3] == pytest.approx(5819000.25 - 20 * 0.2)


def test_compressed_and_bigtiff_layouts(tmp_path):
    a = (np.arange(3 * 300 * 257).reshape(3, 300, 257) % 251).astype(np.uint8)
    p = tmp_path / "z.tif"
    write_geotiff(p, a, GEO, compress=True)
    b, _ = read_geotiff(p)
    assert np.array_equal(a, b)
    # Pillow reads what we write (single band and RGB; Pillow has no 4-band uint8 mode without alpha semantics)
    write_geotiff(tmp_path / "g.tif", a[0], GEO)
    assert np.array_equal(np.asarray(PIL.open(tmp_path / "g.tif")), a[0])
    write_geotiff(tmp_path / "rgb.tif", a, None)
    assert np.array_equal(np.asarray(PIL.open(tmp_path / "rgb.tif")), np.moveaxis(a, 0, 2))


@pytest.mark.parametrize("tile,planar,compress", [(64, False, False), (32, True, True), (None, True, False), (48, False, True)])
def test_tiled_and_planar_layouts(tmp_path, tile, planar, compress):
    """The layouts large source rasters come in (GDAL TILED=YES / INTERLEAVE=BAND): written here, read back whole and by
    window, and decoded by Pillow's independent reader (single band and RGB)."""
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, size=(3, 150, 203), dtype=np.uint8)
    p = tmp_path / "l.tif"
    write_geotiff(p, a, GEO, tile=tile, planar=planar, compress=compress)
    b, g = read_geotiff(p)
    assert np.array_equal(a, b) and g.geotransform == GEO.geotransform
    w, gw = read_geotiff(p, window=(60, 70, 100, 61))
    assert np.array_equal(w, a[:, 70:131, 60:160])
    if not planar:      # Pillow's own decoder has no raw mode for several band-sequential layouts
        assert np.array_equal(np.asarray(PIL.open(p)), np.moveaxis(a, 0, 2))
    a16 = rng.integers(0, 60000, size=(1, 97, 130), dtype=np.uint16)
    write_geotiff(p, a16, None, tile=tile, planar=planar, compress=compress)
    assert np.array_equal(read_geotiff(p)[0], a16) and (planar or np.array_equal(np.asarray(PIL.open(p)), a16[0]))
    with pytest.raises(ValueError):
        write_geotiff(p, a, GEO, tile=50)


@pytest.mark.parametrize("compression", [None, "tiff_lzw", "tiff_adobe_deflate", "packbits"])
def test_reads_pillow_files(tmp_path, compression):
    rng = np.random.default_rng(1)
    # smooth-ish content so that LZW builds long strings and crosses the 9/10/11/12-bit code widths
    base = np.cumsum(rng.integers(-2, 3, size=(180, 211, 3)), axis=1).astype(np.int64)
    rgb = (base % 256).astype(np.uint8)
    p = tmp_path / "p.tif"
    PIL.fromarray(rgb).save(p, compression=compression)
    a, g = read_geotiff(p)
    assert a.shape == (3, 180, 211) and np.array_equal(np.moveaxis(a, 0, 2), rgb) and not g.georeferenced
    g16 = (base[:, :, 0] % 65536).astype(np.uint16)
    PIL.fromarray(g16).save(p, compression=compression)
    a, _ = read_geotiff(p)
    assert a.dtype == np.uint16 and np.array_equal(a[0], g16)
    w, _ = read_geotiff(p, window=(200, 170, 11, 10))
    assert np.array_equal(w[0], g16[170:180, 200:211])


def test_tile_helpers_and_class_zero(tmp_path):
    (tmp_path / "trai" / "img_tiles").mkdir(parents=True)
    (tmp_path / "trai" / "mask_tiles").mkdir(parents=True)
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(4, 32, 32), dtype=np.uint8)
    msk = rng.integers(0, 3, size=(32, 32), dtype=np.uint8)
    fn = tmp_path / "trai" / "img_tiles" / "a.tif"
    write_geotiff(fn, img, GEO)
    write_geotiff(tmp_path / "trai" / "mask_tiles" / "a.tif", msk, GEO)
    assert np.array_equal(open_tile(fn), img)                      # uint8 stays raw (divided by 255 on the device)
    assert np.array_equal(open_mask(fn), msk)                      # utils.py:51-55 get_y: band 1 of mask_tiles/<name>
    write_geotiff(fn, img.astype(np.uint16) * 3, GEO)
    x = open_tile(fn, chnls=[0, 2])                                # 16-bit tiles stay raw too (A0 input contract on the device)
    assert x.dtype == np.uint16 and np.array_equal(x, img[[0, 2]].astype(np.uint16) * 3)
    write_geotiff(fn, img.astype(np.float32) * 3, GEO)
    x = open_tile(fn, chnls=[0, 2])
    assert x.dtype == np.float32 and np.allclose(x, img[[0, 2]].astype(np.float32) * 3 / 255.0)
    # predict.py:34-36: class 0 -> nodata, other classes decremented
    out = tmp_path / "cz.tif"
    write_geotiff(out, msk, GEO, nodata=255, class_zero=True)
    b, g = read_geotiff(out)
    assert np.array_equal(b[0], np.where(msk == 0, 255, msk - 1)) and g.nodata == 255
    with pytest.raises(FileNotFoundError):
        read_geotiff(tmp_path / "missing.tif")


def test_split_raster_tiles_filter_and_split(tmp_path):
    """create_tiles_unet.py:252-431: window order / offsets (slidingwindow semantics), empty-tile filter, per-tile
    georeferencing, class_zero shift and the trai/vali/test distribution."""
    from unet_b200.create_tiles import compute_windows, split_raster
    rng = np.random.default_rng(3)
    H, W, P, ov = 300, 420, 128, 0.25
    img = rng.integers(1, 256, size=(4, H, W), dtype=np.uint8)
    img[:, :, :140] = 0                                 # an empty band on the left: tiles there are dropped
    msk = rng.integers(0, 2, size=(H, W), dtype=np.uint8)
    write_geotiff(tmp_path / "scene.tif", img, GEO)
    write_geotiff(tmp_path / "scene_mask.tif", msk, GEO)
    wins = compute_windows(np.zeros((H, W, 4)), P, ov)
    assert wins[0] == (0, 0, P, P) and wins[1][0] == 0 and wins[1][1] == 96     # x-outer, y-inner; step = 128 - 32
    out = split_raster(str(tmp_path / "scene.tif"), str(tmp_path / "scene_mask.tif"), str(tmp_path / "ds"), P, ov,
                       [0.5, 0.5, 0.0], 0.5, True)
    kept = [i for i, (x, y, w, h) in enumerate(wins)
            if np.sum(img[:, y:y + h, x:x + w] != 0) >= 4 * w * h * 0.5]
    assert sorted(int(p.stem.split("_")[-1]) for p in out) == kept and 0 < len(kept) < len(wins)
    assert {p.parent.parent.name for p in out} == {"trai", "vali"} and not (tmp_path / "ds" / "img_tiles").exists()
    p0 = out[0]
    i0 = int(p0.stem.split("_")[-1])
    x, y, w, h = wins[i0]
    tile, g = read_geotiff(p0)
    assert np.array_equal(tile, img[:, y:y + h, x:x + w])
    assert g.geotransform[0] == pytest.approx(GEO.geotransform[0] + x * 0.2) and g.geotransform[3] == pytest.approx(GEO.geotransform[3] - y * 0.2)
    m, _ = read_geotiff(str(p0).replace("img_tiles", "mask_tiles"))
    assert np.array_equal(m[0], msk[y:y + h, x:x + w] + 1)              # class_zero: labels shifted by one
    with pytest.raises(ValueError):
        split_raster(str(tmp_path / "scene.tif"), None, str(tmp_path / "ds2"), 512, ov)


def test_params_and_main_dispatch(tmp_path):
    """params_and_main.py:121-181: defaults, the forced resets without enable_extra_parameters, unknown names, and the
    Create_tiles stage end to end (the Train / Predict stages call the GPU entry points tested in test_api_gpu.py)."""
    from unet_b200.params_and_main import main, resolve
    p = resolve({"patch_size": 64, "self_attention": True, "large_file": True})
    assert p["self_attention"] is False and p["large_file"] is False and p["ARCHITECTURE"] == "xresnet34"   # :134-147
    with pytest.warns(UserWarning):
        q = resolve({"enable_extra_parameters": True, "self_attention": True, "ENCODER_FACTOR": 5})
    assert q["self_attention"] is True and q["ENCODER_FACTOR"] == 5 and q["data_path"] == q["base_dir"]
    with pytest.raises(KeyError):
        resolve({"patchsize": 64})
    rng = np.random.default_rng(5)
    img = rng.integers(1, 256, size=(4, 160, 160), dtype=np.uint8)
    msk = rng.integers(0, 3, size=(160, 160), dtype=np.uint8)
    write_geotiff(tmp_path / "scene.tif", img, GEO)
    write_geotiff(tmp_path / "scene_mask.tif", msk, GEO)
    js = tmp_path / "params.json"
    js.write_text(__import__("json").dumps({"image_path": str(tmp_path / "scene.tif"), "mask_path": str(tmp_path / "scene_mask.tif"),
                                            "base_dir": str(tmp_path / "ds"), "patch_size": 64, "patch_overlap": 0.25,
                                            "split": [0.75, 0.25]}))
    out = main(str(js))
    assert len(out["tiles"]) == 9 and {t.parent.parent.name for t in out["tiles"]} == {"trai", "vali"}
    assert "learner" not in out and "predictions" not in out
