mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_api_gpu.py tests/test_elementwise_gpu.py tests/test_multigpu.py -m gpu -q ) > gpurun_out/r2_pytest_gpu2.log 2>&1
tail -n 30 gpurun_out/r2_pytest_gpu2.log | cut -c1-250
timeout 300 python tools/predict_profile.py 8192 64 > gpurun_out/r2_predict_profile.txt 2>&1; cat gpurun_out/r2_predict_profile.txt | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-extra > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
cut -c1-260 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
B2U_GRAD_FP32=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-extra --no-predict --no-profile > gpurun_out/r2_bench_n2_fp32.json 2> gpurun_out/r2_bench_n2_fp32.err
cut -c1-260 gpurun_out/r2_bench_n2_fp32.json
