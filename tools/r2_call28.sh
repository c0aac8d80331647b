#!/bin/bash
T='tests/test_network_gpu.py::test_train_step_parity[xresnet34-4-2-256-2-uniform] tests/test_network_gpu.py::test_train_step_parity[xresnet50-4-8-64-2-aerial] tests/test_network_gpu.py::test_self_attention_parity'
for env in "A=1" "B2U_CONV_NO_SOLO=1" "B2U_CONV_MAX_STG=2" "B2U_NO_STEM_IM2COL=1" "B2U_CONV_NO_SOLO=1 B2U_CONV_MAX_STG=2 B2U_NO_STEM_IM2COL=1"; do
  echo "== $env"
  env $env timeout 600 python -m pytest $T -m gpu -q 2>&1 | grep -E "^E       AssertionError|passed|failed" | cut -c1-200
done
