#!/bin/bash
# encoder-shaped convolutions: cost of the BN-statistics epilogue and of the fused finalize; role timelines
for c in e64 e64s e64f e128 e128f e256 e256s e256f e512 e512f; do python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-110; done
for c in e64 e64f e256 e256f e512f; do echo "== timeline $c"; python tools/conv_timeline.py $c 2>&1 | grep -E "mean period|first 12|first 24" | cut -c1-400; done
echo "== n-split A/B"
B2U_CONV_NO_NSPLIT=1 python tools/one_conv.py e512f 20 2>&1 | tail -1 | cut -c1-200
python tools/one_conv.py e512f 20 2>&1 | tail -1 | cut -c1-200
B2U_CONV_NO_NSPLIT=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nosplit', d['ms_per_step'], d['value'], d['roofline']['frac_burst'])"
python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split  ', d['ms_per_step'], d['value'], d['roofline']['frac_burst'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
