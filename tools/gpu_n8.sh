# 8-GPU evidence: the bench line at N=8 (train weak scaling, 20000^2 prediction strong scaling incl. the NCCL gather)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --no-extra > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
cut -c1-260 gpurun_out/r02_bench_8gpu.json; tail -2 gpurun_out/r02_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --no-extra --no-profile > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
cut -c1-260 gpurun_out/r02_bench_4gpu.json
