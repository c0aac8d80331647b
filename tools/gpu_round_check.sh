# Round-end evidence run on ONE B200: GPU tests, the bench line, per-op profile, the ncu launch list of bench.py,
# DRAM/L2/tensor-pipe metrics of every GEMM launch of one step, and one ncu --set full capture of the heaviest conv.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time
cut -c1-400 gpurun_out/bench.json
timeout 200 python tools/layer_profile.py 64 > gpurun_out/layer_profile.log 2>&1
B2U_NO_SIDE_STREAM=1 timeout 200 python tools/op_profile.py 64 > gpurun_out/op_profile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 800 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --no-predict > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"conv_gemm|wgrad_gemm" --launch-skip 368 -c 184 --csv --log-file gpurun_out/step_metrics.csv python tools/step_eager.py 3 64 > gpurun_out/ncu_step.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 2 -c 1 -f -o gpurun_out/conv_res100 python tools/one_conv.py res100 3 > gpurun_out/ncu_conv.log 2>&1
for f in gpurun_out/ncu_step.log gpurun_out/ncu_conv.log; do tail -n 2 $f; done
