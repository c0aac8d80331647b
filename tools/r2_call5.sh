mkdir -p gpurun_out
timeout 300 python tools/parity_probe.py xresnet18 3 2 128 8 aerial > gpurun_out/r2_probe18.txt 2>&1
tail -3 gpurun_out/r2_probe18.txt
