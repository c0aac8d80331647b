mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n$N.json"))
    print("n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "predict", round(d["predict"]["value"],1), "predict e2e", round(d["predict"]["e2e"]["value"],1), d["predict"]["e2e"]["strip_mask_equals_sharded_mask"])
except Exception as e: print("failed", e, open("gpurun_out/bench_n$N.err").read()[-1500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/predict_bench.py 20000 64 > gpurun_out/predict_n$N.json 2> gpurun_out/predict_n$N.err; echo "predict n$N rc=$?"; tail -n 1 gpurun_out/predict_n$N.json
