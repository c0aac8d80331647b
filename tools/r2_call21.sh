mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --no-extra --no-predict --no-profile --no-cpu-baseline 2>gpurun_out/n8_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['reasons'])" || tail -5 gpurun_out/n8_$1.err; }
echo "N=1:"; timeout 300 python bench.py --no-extra --no-predict --no-profile --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3))"
echo "N=8 default (segment overlap):"; run 29801
echo "N=8 no overlap:"; B2U_NO_AR_OVERLAP=1 run 29802
echo "N=8 no overlap, SM_LIMIT=144:"; B2U_NO_AR_OVERLAP=1 B2U_SM_LIMIT=144 run 29803
echo "N=8 overlap, SM_LIMIT=144:"; B2U_SM_LIMIT=144 run 29804
echo "N=8 no overlap, fp32 wire:"; B2U_NO_AR_OVERLAP=1 B2U_GRAD_FP32=1 run 29805
echo "N=8 predict (grid):"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29806 bench.py --gpus 8 --no-extra --no-profile --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['predict']; print(round(p['value']), p['seconds'], p['tiles_run_max_rank'], p['ownership_grid'], p['e2e'])"
