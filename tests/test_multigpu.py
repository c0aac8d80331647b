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
