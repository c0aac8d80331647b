mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -8
tail -n 25 gpurun_out/pytest_gpu.log | grep -E "^E|assert" | head -10
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-predict --no-profile > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_b.json"))
    print("ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "loss", d["final_loss"])
except Exception as e: print("failed", e, open("gpurun_out/bench_b.err").read()[-800:])
PY
