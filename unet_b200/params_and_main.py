"""The reference's top-level switchboard (`params_and_main.py`): the same parameter names, the same three stages
(Create_tiles -> Train -> Predict, params_and_main.py:155-177) and the same positional calls into `split_raster`,
`train_func` and `save_predictions` - dispatched to the B200 implementations of this package.

The reference keeps its parameters as module globals edited in place; here they come from a dict or a JSON file
(`python -m unet_b200.params_and_main params.json`), with the reference's defaults (params_and_main.py:20-118).
Without `enable_extra_parameters` the extra parameters are reset exactly as `main()` does (params_and_main.py:131-147).
Deliberately not replicated: the `pathlib.PosixPath = pathlib.WindowsPath` monkey-patch (:127-128).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time
import warnings
from typing import Any, Dict, Optional, Union

DEFAULTS: Dict[str, Any] = {
    # stages
    "Create_tiles": True, "Train": False, "Predict": False,
    # create tiles (params_and_main.py:29-38)
    "image_path": None, "mask_path": None, "base_dir": ".", "patch_size": 400, "patch_overlap": 0, "split": [0.8, 0.2],
    # training (:45-60)
    "data_path": None, "model_path": None, "description": "model", "info": "", "existing_model": None, "BATCH_SIZE": 4,
    "EPOCHS": 15, "LEARNING_RATE": 1e-4, "enable_regression": False, "visualize_data_example": True,
    "export_model_summary": True, "CODES": ["NO_Data", "Background", "Beschirmung"], "CLASS_WEIGHTS": "even",
    "transforms": True, "split_idx": 0, "n_transform_imgs": 1, "aug_pipe": None,
    # prediction (:74-85)
    "predict_path": None, "predict_model": None, "AOI": None, "year": None, "merge": False, "regression": False,
    "validation_vision": False, "class_zero": False,
    # extra parameters (:87-100), only honoured with enable_extra_parameters
    "enable_extra_parameters": False, "self_attention": True, "ENCODER_FACTOR": 10, "LR_FINDER": None,
    "VALID_SCENES": ["vali"], "loss_func": None, "monitor": "dice_multi", "all_classes": False, "specific_class": None,
    "large_file": False, "max_empty": 0.9, "ARCHITECTURE": "xresnet34",
    # not in the reference (single GPU): data-parallel training / sharded prediction over the GPUs of one box
    "N_GPUS": 1,
}

# what main() forces when enable_extra_parameters is off (params_and_main.py:134-147)
_FORCED = {"ENCODER_FACTOR": 10, "LR_FINDER": None, "VALID_SCENES": ["vali"], "loss_func": None, "monitor": None,
           "all_classes": False, "specific_class": None, "enable_regression": False, "large_file": False,
           "max_empty": 0.9, "ARCHITECTURE": "xresnet34", "self_attention": False}


def resolve(params: Union[str, Dict[str, Any], None]) -> Dict[str, Any]:
    """defaults <- user parameters (dict or JSON path) <- the resets of main() when extra parameters are disabled."""
    if isinstance(params, str):
        with open(params) as f:
            params = json.load(f)
    p = dict(DEFAULTS)
    unknown = sorted(set(params or {}) - set(DEFAULTS))
    if unknown:
        raise KeyError(f"unknown parameter(s) {unknown}; the reference's names are {sorted(DEFAULTS)}")
    p.update(params or {})
    if p["enable_extra_parameters"]:
        warnings.warn("Extra parameters are enabled. Code may behave in unexpected ways. "
                      "Please disable unless experienced with the code.")                     # params_and_main.py:130-132
    else:
        p.update(_FORCED)
    if p["data_path"] is None:
        p["data_path"] = p["base_dir"]                                                        # :44 data_path = base_dir
    return p


def main(params: Union[str, Dict[str, Any], None] = None) -> Dict[str, Any]:
    """params_and_main.py:121-181.  Returns what each stage produced (tile paths, learner, prediction paths)."""
    p = resolve(params)
    out: Dict[str, Any] = {}
    t0 = time.time()
    n_gpus = int(p.get("N_GPUS") or 1)
    if n_gpus > 1 and "RANK" not in os.environ and (p["Train"] or p["Predict"]):
        # one process per GPU: tile creation runs here once, then Train / Predict are re-entered under torch.distributed.run
        # (NCCL over NVLink); every rank trains on its share of the batches and rank 0 writes the files
        if p["Create_tiles"]:
            main({**p, "Train": False, "Predict": False, "N_GPUS": 1, "enable_extra_parameters": True})
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump({**p, "Create_tiles": False, "enable_extra_parameters": True}, f)
        port = 29500 + os.getpid() % 2000
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "unet_b200.params_and_main",
                        f.name], check=True)
        os.unlink(f.name)
        return out
    if p["Create_tiles"]:
        from .create_tiles import split_raster
        out["tiles"] = split_raster(path_to_raster=p["image_path"], path_to_mask=p["mask_path"], patch_size=p["patch_size"],
                                    patch_overlap=p["patch_overlap"], base_dir=p["base_dir"], split=p["split"],
                                    max_empty=p["max_empty"], class_zero=p["class_zero"])
    if p["Train"]:
        from .reference_api import train_func
        out["learner"] = train_func(p["data_path"], p["existing_model"], p["model_path"], p["description"], p["BATCH_SIZE"],
                                    p["visualize_data_example"], p["enable_regression"], p["CLASS_WEIGHTS"],
                                    p["ARCHITECTURE"], p["EPOCHS"], p["LEARNING_RATE"], p["ENCODER_FACTOR"], p["LR_FINDER"],
                                    p["loss_func"], p["monitor"] or "dice_multi", p["self_attention"], p["VALID_SCENES"],
                                    p["CODES"], p["transforms"], p["split_idx"], p["export_model_summary"], p["aug_pipe"],
                                    p["n_transform_imgs"], p["info"], p["class_zero"])
    if p["Predict"]:
        from .reference_api import save_predictions
        out["predictions"] = save_predictions(p["predict_model"], p["predict_path"], p["regression"], p["merge"],
                                              p["all_classes"], p["specific_class"], p["large_file"], p["AOI"], p["year"],
                                              p["validation_vision"], class_zero=p["class_zero"])
    dt = time.time() - t0
    print(f"The operation took {dt:.2f} seconds or {dt / 60:.2f} minutes")                   # params_and_main.py:179-180
    return out


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
