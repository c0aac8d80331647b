mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-predict > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
for f in ("b",):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"],3), "tiles/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv ms", round(d["kernels"]["conv"]["ms"],2), "wgrad ms", round(d["kernels"]["wgrad"]["ms"],2), "loss", d["final_loss"])
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-800:])
PY
timeout 200 python tools/layer_profile.py 64 > gpurun_out/layer_profile.log 2>&1; grep wgrad gpurun_out/layer_profile.txt | sort -k2 -n -r | head -12
