mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_network_gpu.py -m gpu -q -k "raster or eval" ) > gpurun_out/r2_pytest_pred.log 2>&1
tail -n 6 gpurun_out/r2_pytest_pred.log | cut -c1-250
timeout 600 python bench.py --no-extra --no-cpu-baseline --no-profile > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
tail -2 gpurun_out/r2_bench_f.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_f.json"))
print("train", d["value"], "predict", d["predict"]["value"], d["predict"]["seconds"], d["predict"]["e2e"], d["predict"]["workload"][-140:])
PY
