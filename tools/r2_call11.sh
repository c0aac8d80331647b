mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -n 25 gpurun_out/r2_pytest_gpu.log | cut -c1-250
for c in c100_100 res100 c96_96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
timeout 300 python tools/layer_profile.py 64 > gpurun_out/r2_layer_profile.log 2>&1
head -3 gpurun_out/r2_layer_profile.log
( time timeout 900 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err ) 2> gpurun_out/r2_bench_d.time
tail -3 gpurun_out/r2_bench_d.err; cat gpurun_out/r2_bench_d.time; cut -c1-300 gpurun_out/r2_bench_d.json
