mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
( time timeout 1500 python -m pytest tests/test_network_gpu.py -m gpu -q ) > gpurun_out/r2_pytest_network.log 2>&1
tail -n 30 gpurun_out/r2_pytest_network.log | cut -c1-250; cat gpurun_out/tf_parity.txt
for c in c64 res100 c100_100; do timeout 120 python tools/conv_timeline.py $c > gpurun_out/r2_timeline_$c.txt 2>&1; done
