#!/bin/bash
# last sanity run of the round's final commit + one --set full capture of the new statistics epilogue
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2 )
python bench.py --no-extra > gpurun_out/r02_bench_1gpu_last.json 2>/dev/null; cut -c1-260 gpurun_out/r02_bench_1gpu_last.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu_last.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["frac_burst"], d["roofline"]["launches_per_step"], d["clocks"], d["e2e"]["value"], d["predict"]["value"])
PY
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 2 -c 1 -f -o gpurun_out/r02_conv_e64f python tools/one_conv.py e64f 3 > gpurun_out/ncu_e64f.log 2>&1; tail -1 gpurun_out/ncu_e64f.log | cut -c1-150
