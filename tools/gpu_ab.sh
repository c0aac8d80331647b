mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -5
for v in a b; do
  if [ $v = a ]; then export B2U_WGRAD_UNEVEN=1; else unset B2U_WGRAD_UNEVEN; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-predict > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python - <<'PY'
import json
for f in ("a","b"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv ms", round(d["kernels"]["conv"]["ms"],2), "wgrad ms", round(d["kernels"]["wgrad"]["ms"],2), "loss", d["final_loss"])
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-800:])
PY
