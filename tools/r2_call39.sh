#!/bin/bash
# the library on fewer SMs than the device has (B2U_SM_LIMIT): workspace query and plan creation agree, results unchanged
for n in 140 100; do
  echo "== B2U_SM_LIMIT=$n"
  B2U_SM_LIMIT=$n timeout 150 python bench.py --no-extra --no-predict --no-cpu-baseline --no-profile --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-230
done
B2U_SM_LIMIT=132 timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_network_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -1
