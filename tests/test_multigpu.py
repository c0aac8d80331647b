"""Hardware multi-rank parity (SURVEY 4 tier 4): needs >= 2 GPUs on the box (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).  tools/ddp_check.py holds the assertions."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_data_parallel_step_and_sharded_prediction_match_one_gpu():
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tools", "ddp_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert out.count("-> OK") == 3 and "FAIL" not in out, out[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_params_and_main_trains_data_parallel(tmp_path):
    """The reference's switchboard with N_GPUS = 2 (params_and_main.py:121-177 + the new parameter): main() re-enters Train
    under torch.distributed.run, every rank trains on its share of the batches, rank 0 writes the model files."""
    import json
    import numpy as np
    sys.path.insert(0, ROOT)
    from unet_b200.geotiff import GeoInfo, write_geotiff
    from unet_b200.synth import aerial_like_tiles
    x, y = aerial_like_tiles(40, 4, 64, 64, 2, seed=5)
    for scene, sl in (("trai", slice(0, 32)), ("vali", slice(32, 40))):
        for sub in ("img_tiles", "mask_tiles"):
            (tmp_path / "data" / scene / sub).mkdir(parents=True)
        for i in range(sl.start, sl.stop):
            write_geotiff(tmp_path / "data" / scene / "img_tiles" / f"t{i}.tif", x[i].numpy(), GeoInfo())
            write_geotiff(tmp_path / "data" / scene / "mask_tiles" / f"t{i}.tif", y[i].numpy(), GeoInfo())
    params = {"Create_tiles": False, "Train": True, "Predict": False, "data_path": str(tmp_path / "data"),
              "model_path": str(tmp_path / "models"), "description": "dp", "BATCH_SIZE": 4, "EPOCHS": 2, "LEARNING_RATE": 1e-3,
              "CODES": ["a", "b"], "transforms": False, "enable_extra_parameters": True, "ARCHITECTURE": "xresnet18",
              "self_attention": False, "N_GPUS": 2, "visualize_data_example": False, "export_model_summary": False}
    pj = tmp_path / "params.json"
    pj.write_text(json.dumps(params))
    r = subprocess.run([sys.executable, "-m", "unet_b200.params_and_main", str(pj)], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    d = tmp_path / "models" / "dp"
    meta = json.loads((d / "dp.json").read_text())
    assert (d / "dp.pkl").exists() and meta["n_gpus"] == 2 and len(meta["history"]) == 2
    assert all(np.isfinite(h["train_loss"]) and np.isfinite(h["valid_loss"]) for h in meta["history"])
    assert meta["train_tiles"] == 32                  # 8 global batches of 4 -> 4 optimizer steps per rank and epoch
