mkdir -p gpurun_out
for it in 1024 2048 4096 8192; do echo "--- B2U_BN_BWD_ITEMS=$it"; B2U_BN_BWD_ITEMS=$it B2U_NO_SIDE_STREAM=1 timeout 200 python tools/op_profile.py 64 2>&1 | grep -E "sum of per-op|_bn_bwd"; B2U_BN_BWD_ITEMS=$it timeout 200 python bench.py --no-cpu-baseline --no-predict --no-extra --no-profile 2>/dev/null | cut -c85-190; done
