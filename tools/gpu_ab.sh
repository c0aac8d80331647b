mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -5
B2U_NO_FUSED_SHUFFLE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
for f in ("a","b"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv ms", round(d["kernels"]["conv"]["ms"],2), "wgrad ms", round(d["kernels"]["wgrad"]["ms"],2), "loss", d["final_loss"], "predict", round(d["predict"]["value"],1), round(d["predict"]["e2e"]["value"],1))
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-800:])
PY
