mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_elementwise_gpu.py tests/test_network_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in a b; do
  if [ $v = a ]; then export B2U_SHUFFLE_PER_PIXEL=1; else unset B2U_SHUFFLE_PER_PIXEL; fi
  timeout 200 python tools/eval_op_profile.py 64 2>&1 | head -4 | cut -c1-150
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-predict --no-profile > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python - <<'PY'
import json
for f in ("a","b"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "loss", d["final_loss"])
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-800:])
PY
