# Final evidence run of the round on ONE B200 (trimmed gpu_round_check.sh: no torch baseline, no --set full captures)
R=${1:-r02}
mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${R}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/${R}_pytest_gpu.log; cp gpurun_out/tf_parity.txt gpurun_out/${R}_tf_parity.txt 2>/dev/null
( time python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench.err ) 2> gpurun_out/${R}_bench.time
cut -c1-300 gpurun_out/${R}_bench_1gpu.json
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python tools/layer_profile.py 64 > gpurun_out/layer_profile.log 2>&1; cp gpurun_out/layer_profile.txt gpurun_out/${R}_layer_profile.txt
B2U_NO_SIDE_STREAM=1 timeout 200 python tools/op_profile.py 64 > gpurun_out/op_profile.log 2>&1; cp gpurun_out/op_profile.txt gpurun_out/${R}_op_profile.txt
timeout 200 python tools/predict_profile.py 8192 64 > gpurun_out/${R}_predict_profile.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 800 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --no-predict --no-extra > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"conv_gemm|wgrad_gemm" --launch-skip 368 -c 184 --csv --log-file gpurun_out/${R}_gemm_kernel_metrics.csv python tools/step_eager.py 3 64 > gpurun_out/ncu_step.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bn_|shuffle|maxpool|ce_|pointwise|copy_lanes|nchw|im2col|sgd|stage_weights|wgrad_reduce" --launch-skip 340 -c 170 --csv --log-file gpurun_out/${R}_mem_kernel_metrics.csv python tools/step_eager.py 3 64 > gpurun_out/ncu_step_mem.log 2>&1
for f in gpurun_out/ncu_bench.log gpurun_out/ncu_step.log gpurun_out/ncu_step_mem.log; do tail -n 1 $f | cut -c1-200; done
