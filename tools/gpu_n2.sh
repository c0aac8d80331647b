mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_smoke.py > gpurun_out/ddp_smoke.log 2>&1; echo "ddp_smoke rc=$?"; tail -8 gpurun_out/ddp_smoke.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/predict_bench.py 20000 64 > gpurun_out/predict_n2.json 2> gpurun_out/predict_n2.err; echo "predict n2 rc=$?"; tail -3 gpurun_out/predict_n2.json
timeout 300 python tools/predict_bench.py 20000 64 > gpurun_out/predict_n1.json 2> gpurun_out/predict_n1.err; echo "predict n1 rc=$?"; tail -3 gpurun_out/predict_n1.json
