mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time
timeout 200 python tools/layer_profile.py 64 > gpurun_out/layer_profile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 1300 --csv --log-file gpurun_out/step_metrics.csv python tools/step_eager.py 2 64 > gpurun_out/ncu_step.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 2 -c 1 -f -o gpurun_out/conv_res100 python tools/one_conv.py res100 3 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-600; tail -2 gpurun_out/layer_profile.log
