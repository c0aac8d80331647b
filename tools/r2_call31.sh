#!/bin/bash
OLD=$PWD/unet_b200/_obj_old/libb2u_oldconv.so
for c in c100_100 res100 c96_96 c256 e128f; do
  echo -n "old: "; B2U_LIB=$OLD python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60
  echo -n "new: "; python tools/one_conv.py $c 20 2>&1 | tail -1 | cut -c1-60
done
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["roofline"]["frac_burst"], "predict", d["predict"]["value"], d["e2e"]["value"])'
B2U_LIB=$OLD B2U_NO_STEM_IM2COL=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" old_kernel
B2U_NO_STEM_IM2COL=1 B2U_CONV_NO_SOLO=1 B2U_CONV_MAX_STG=2 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" new_kernel_old_modes
python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" new_default
B2U_LIB=$OLD B2U_NO_STEM_IM2COL=1 python bench.py --no-extra --steps 20 --warmup 5 2>/dev/null | python -c "$J" old_kernel_again
timeout 600 python -m pytest tests/test_network_gpu.py -m gpu -q -k "self_attention" 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-400 | head
