#!/bin/bash
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_elementwise_gpu.py tests/test_network_gpu.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-250 | head -30
for c in e64 e64f s32 s32f s32_64f s64_32 e128f c96_96; do python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60; done
echo "== bench"
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["roofline"]["frac_burst"], "predict", d["predict"]["value"], d["e2e"]["value"])'
python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" default
B2U_NO_STEM_IM2COL=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" no_im2col
B2U_CONV_NO_SOLO=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" no_solo
B2U_CONV_NO_SOLO=1 B2U_NO_STEM_IM2COL=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" neither
