mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_smoke.py > gpurun_out/ddp_smoke.log 2>&1; echo "ddp_smoke rc=$?"; grep -E "identical|done" gpurun_out/ddp_smoke.log | head -4
i=0
for seg in 24 45 1000; do
  i=$((i+1))
  B2U_AR_MIN_SEG_MB=$seg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --steps 20 --warmup 5 --no-profile --no-predict > gpurun_out/bench_n2_$seg.json 2> gpurun_out/bench_n2_$seg.err; echo "bench n2 seg=$seg rc=$?"
done
python - <<'PY'
import json
for f in ("24","45","1000"):
    try:
        d=json.load(open(f"gpurun_out/bench_n2_{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "loss", d["final_loss"])
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_n2_{f}.err").read()[-1500:])
PY
