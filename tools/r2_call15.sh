mkdir -p gpurun_out; rm -f gpurun_out/tf_parity.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu3.log 2>&1
tail -n 15 gpurun_out/r2_pytest_gpu3.log | cut -c1-250
( time timeout 900 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err ) 2> gpurun_out/r2_bench_e.time
tail -2 gpurun_out/r2_bench_e.err | cut -c1-200; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_e.json"))
print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["frac_burst"])
print("predict", d["predict"]["value"], d["predict"]["seconds"], d["predict"]["e2e"]["value"])
for k, v in d["extra"].items(): print(k, v.get("value"), v.get("error"))
PY
timeout 200 python tools/predict_profile.py 8192 64 2>&1 | tail -7
