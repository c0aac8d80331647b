#!/bin/bash
# small-channel layers: which planner mode is the fastest
for c in e64 s32 s32f s32_64f s64_32; do
  for env in "" "B2U_CONV_NO_PAIR=1" "B2U_CONV_NO_WRES=1" "B2U_CONV_NO_WRES=1 B2U_CONV_NO_ROWMODE=1" "B2U_CONV_NO_HALO=1" "B2U_CONV_NO_PAIR=1 B2U_CONV_NO_WRES=1 B2U_CONV_NO_ROWMODE=1"; do
    echo -n "[$env] "; env $env python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60
  done
done
for c in s32 s32f; do echo "== timeline $c"; python tools/conv_timeline.py $c --rebuild 2>&1 | grep -E "mean period|first 12|first 24" | cut -c1-300; done
timeout 300 python -m pytest "tests/test_network_gpu.py" -m gpu -x -q -k "xresnet50" 2>&1 | tail -2
