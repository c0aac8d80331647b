mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
( time timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_elementwise_gpu.py -m gpu -q -x ) > gpurun_out/r2_pytest_ops.log 2>&1
tail -n 5 gpurun_out/r2_pytest_ops.log
( time timeout 1500 python -m pytest tests/test_network_gpu.py -m gpu -q ) > gpurun_out/r2_pytest_network.log 2>&1
tail -n 30 gpurun_out/r2_pytest_network.log; cat gpurun_out/tf_parity.txt
timeout 300 python tools/layer_profile.py 64 > gpurun_out/r2_layer_profile.txt 2>&1
head -3 gpurun_out/r2_layer_profile.txt
timeout 300 python bench.py --no-cpu-baseline --no-predict > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
cut -c1-300 gpurun_out/r2_bench_a.json
