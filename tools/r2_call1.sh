# round-2 diagnostic call 1: role timelines of the 100-channel conv, parity probes (incl. ours-vs-emulation gradients),
# the cuDNN secondary bar
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 300 python tools/conv_timeline.py res100 > gpurun_out/r2_timeline_res100.txt 2>&1
timeout 120 python tools/conv_timeline.py c100_100 > gpurun_out/r2_timeline_c100.txt 2>&1
timeout 120 python tools/conv_timeline.py shuf96 > gpurun_out/r2_timeline_shuf96.txt 2>&1
timeout 120 python tools/conv_timeline.py c256 > gpurun_out/r2_timeline_c256.txt 2>&1
timeout 600 python tools/parity_probe.py > gpurun_out/r2_parity_probe.txt 2>&1
timeout 200 python tools/parity_probe.py xresnet34 4 2 128 16 aerial >> gpurun_out/r2_parity_probe.txt 2>&1
timeout 300 python tools/torch_baseline.py 64 > gpurun_out/r2_torch_baseline.txt 2>&1
tail -n 3 gpurun_out/r2_timeline_res100.txt gpurun_out/r2_torch_baseline.txt
