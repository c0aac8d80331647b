mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
timeout 300 python tools/grad_bisect.py xresnet18 3 2 128 8 > gpurun_out/r2_grad_bisect.log 2>&1
head -45 gpurun_out/r2_grad_bisect.log | cut -c1-150
for c in c64 res100 c100_100; do timeout 120 python tools/conv_timeline.py $c > gpurun_out/r2_timeline_$c.txt 2>&1; head -6 gpurun_out/r2_timeline_$c.txt | cut -c1-200; done
( time timeout 1500 python -m pytest tests/test_network_gpu.py -m gpu -q ) > gpurun_out/r2_pytest_network.log 2>&1
tail -n 30 gpurun_out/r2_pytest_network.log | cut -c1-250; cat gpurun_out/tf_parity.txt
