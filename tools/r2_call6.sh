mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
( time timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x ) > gpurun_out/r2_pytest_gemm.log 2>&1
tail -n 15 gpurun_out/r2_pytest_gemm.log | cut -c1-200
timeout 300 python tools/grad_bisect.py xresnet18 3 2 128 8 > gpurun_out/r2_grad_bisect.log 2>&1
head -50 gpurun_out/r2_grad_bisect.log | cut -c1-150
timeout 300 python tools/layer_profile.py 64 > gpurun_out/r2_layer_profile.log 2>&1
head -3 gpurun_out/r2_layer_profile.log
( time timeout 900 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "eval or raster or golden or attention" ) > gpurun_out/r2_pytest_network2.log 2>&1
tail -n 5 gpurun_out/r2_pytest_network2.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-predict > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
cut -c1-200 gpurun_out/r2_bench_b.json
