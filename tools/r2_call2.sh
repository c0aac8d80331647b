mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
for c in res100 c100_100 shuf96 c256 c96_96; do timeout 120 python tools/conv_timeline.py $c > gpurun_out/r2_timeline_$c.txt 2>&1; done
( time timeout 1200 python -m pytest tests/test_network_gpu.py -m gpu -q -x ) > gpurun_out/r2_pytest_network.log 2>&1
tail -n 25 gpurun_out/r2_pytest_network.log; cat gpurun_out/tf_parity.txt
