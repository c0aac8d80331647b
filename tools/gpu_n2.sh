# 2-GPU evidence: the hardware multi-rank parity tests (gradient equality, gathered mask, N_GPUS through the switchboard)
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q ) > gpurun_out/r02_pytest_multigpu.log 2>&1
tail -n 14 gpurun_out/r02_pytest_multigpu.log | cut -c1-300
