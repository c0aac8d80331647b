mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x ) > gpurun_out/r2_pytest_gemm.log 2>&1
tail -n 6 gpurun_out/r2_pytest_gemm.log | cut -c1-200
for c in c100_100 res100 c64 c96_96 shuf96 c256; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
for c in c100_100 res100; do timeout 120 python tools/conv_timeline.py $c > gpurun_out/r2_tl3_$c.txt 2>&1; done
timeout 300 python tools/layer_profile.py 64 > gpurun_out/r2_layer_profile.log 2>&1
head -3 gpurun_out/r2_layer_profile.log
( time timeout 900 python -m pytest tests/test_network_gpu.py -m gpu -q -x ) > gpurun_out/r2_pytest_network.log 2>&1
tail -n 5 gpurun_out/r2_pytest_network.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-predict > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
cut -c1-200 gpurun_out/r2_bench_c.json
