mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
for b in 64 96 128 192; do timeout 200 python tools/predict_profile.py 8192 $b 2>&1 | grep -E "predict_raster|forward"; done
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
