# 2-GPU evidence: the hardware multi-rank parity test and the 2-GPU bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q ) > gpurun_out/r02_pytest_multigpu.log 2>&1
tail -n 12 gpurun_out/r02_pytest_multigpu.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py 2>&1 | grep ddp_check
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
cut -c1-260 gpurun_out/r02_bench_2gpu.json
