mkdir -p gpurun_out
echo "--- pitch 64-rounding (default)"; for c in c100_100 res100 c96_96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
timeout 60 python tools/one_wgrad.py 2>&1 | tail -2 | cut -c1-120
echo "--- pitch 16"; export B2U_PITCH16=1; for c in c100_100 res100 c96_96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
timeout 60 python tools/one_wgrad.py 2>&1 | tail -2 | cut -c1-120
