#!/bin/bash
# staging ring (up to 4 buffers) + statistics from the staged tile: correctness first, then A/B
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_network_gpu.py -m gpu -x -q 2>&1 | tail -4
for c in e64 e64s e64f e128 e128f e256 e256f e512 e512f c100_100 c96_96; do python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-70; done
echo "== max 2 staging buffers"
for c in e64 e64f e128f c96_96; do B2U_CONV_MAX_STG=2 python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-70; done
echo "== bench"
python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['roofline']['frac_burst'], 'predict', d['predict']['value'] if 'predict' in d else [k for k in d])"
B2U_CONV_MAX_STG=2 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train(stg2)', d['ms_per_step'], d['value'], d['roofline']['frac_burst'])"
python tools/conv_timeline.py e64f --rebuild 2>&1 | grep -E "mean period" | cut -c1-200
