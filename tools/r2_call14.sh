mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gemm_gpu.py tests/test_elementwise_gpu.py -m gpu -q -x -k "batched or softmax_dim1 or crop or layout" ) > gpurun_out/r2_pytest_attn.log 2>&1
tail -n 15 gpurun_out/r2_pytest_attn.log | cut -c1-250
( time timeout 1800 python -m pytest tests/test_network_gpu.py tests/test_api_gpu.py -m gpu -q -k "attention or predict or raster" ) > gpurun_out/r2_pytest_attn2.log 2>&1
tail -n 15 gpurun_out/r2_pytest_attn2.log | cut -c1-250
