mkdir -p gpurun_out
for c in c100_100 res100; do timeout 120 python tools/conv_timeline.py $c > gpurun_out/r2_tl2_$c.txt 2>&1; done
echo "--- default order"; for c in c100_100 res100 c64 c96_96 shuf96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
echo "--- SA2 full NB first"; export B2U_CONV_WRES_TRY="2,0;3,1;2,1;2,1"; for c in c100_100 res100 c64 c96_96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
echo "--- no wres"; unset B2U_CONV_WRES_TRY; export B2U_CONV_NO_WRES=1; for c in c100_100 res100 c64 c96_96; do timeout 60 python tools/one_conv.py $c 10 2>&1 | cut -c1-60; done
