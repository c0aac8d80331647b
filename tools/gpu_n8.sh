# 8-GPU evidence: the bench line at N=8 (train weak scaling, 20000^2 prediction strong scaling incl. the NCCL gather), and
# what the node itself costs: the same step with the all-reduce skipped (diagnostic), as 8 unrelated single-GPU
# processes at once, and on one GPU with the other seven idle.
mkdir -p gpurun_out
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d.get("clocks",{}); print("%-28s %.3f ms/step  %.0f tiles/s  sm %s MHz  %s W  %s" % (sys.argv[1], d["ms_per_step"], d["value"], c.get("sm_mhz"), c.get("power_w"), c.get("reasons")))'
LIGHT="--no-extra --no-predict --no-profile --no-cpu-baseline --steps 40 --warmup 8"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
python bench.py $LIGHT 2>/dev/null | python -c "$J" "N=1, seven GPUs idle" | tee gpurun_out/r02_n8_experiments.txt
timeout 900 $TR --master-port 29621 bench.py --gpus 8 --no-extra > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
python -c "$J" "N=8 default (full line)" < gpurun_out/r02_bench_8gpu.json | tee -a gpurun_out/r02_n8_experiments.txt
timeout 600 $TR --master-port 29622 bench.py --gpus 8 $LIGHT 2>/dev/null | python -c "$J" "N=8 default" | tee -a gpurun_out/r02_n8_experiments.txt
B2U_DIAG_SKIP_ALLREDUCE=1 timeout 600 $TR --master-port 29623 bench.py --gpus 8 $LIGHT 2>/dev/null | python -c "$J" "N=8 all-reduce skipped" | tee -a gpurun_out/r02_n8_experiments.txt
for i in 0 1 2 3 4 5 6 7; do
  ( CUDA_VISIBLE_DEVICES=$i python bench.py $LIGHT 2>/dev/null | python -c "$J" "8 single-GPU jobs, GPU $i" > gpurun_out/solo_$i.txt ) &
done
wait
cat gpurun_out/solo_?.txt | tee -a gpurun_out/r02_n8_experiments.txt
python bench.py $LIGHT 2>/dev/null | python -c "$J" "N=1 again, seven GPUs idle" | tee -a gpurun_out/r02_n8_experiments.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_8gpu.json").read().strip().splitlines()[-1])
print("predict:", {k: d["predict"][k] for k in ("value", "seconds", "tiles_run_max_rank", "ownership_grid")})
PY
