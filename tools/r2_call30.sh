#!/bin/bash
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_elementwise_gpu.py tests/test_network_gpu.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-300 | head -30
for c in shuf96 e64 e64f s32 s32f c100_100 res100 c96_96 c256; do python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60; done
echo "== two barriers, one staging buffer"
for c in shuf96 e64 s32 c100_100; do B2U_CONV_MAX_STG=1 python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60; done
echo "== bench"
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["roofline"]["frac_burst"], "predict", d["predict"]["value"], d["e2e"]["value"])'
python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" default
for env in "A=1" "B2U_NO_STEM_IM2COL=1"; do
  echo "== $env"; env $env timeout 200 python tools/predict_profile.py 8192 64 2>&1 | grep -E "predict_raster|forward" | cut -c1-150
done
python tools/eval_op_profile.py 2>&1 | head -14 | cut -c1-160
tail -9 gpurun_out/tf_parity.txt | cut -c1-250
