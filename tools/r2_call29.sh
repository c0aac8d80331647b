#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-300 | head -30
for env in "A=1" "B2U_NO_STEM_IM2COL=1" "B2U_CONV_NO_SOLO=1"; do
  echo "== $env"; env $env timeout 200 python tools/predict_profile.py 8192 64 2>&1 | grep -E "predict_raster|forward" | cut -c1-150
done
cat gpurun_out/tf_parity.txt | tail -12 | cut -c1-260
