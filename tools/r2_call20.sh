mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --no-extra --no-predict --no-profile --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['reasons'])"; }
echo "N=1:"; timeout 300 python bench.py --no-extra --no-predict --no-profile --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3))"
echo "default (bf16 wire, NCCL_MAX_CTAS=8):"; run 29701
echo "SM_LIMIT=140:"; B2U_SM_LIMIT=140 run 29702
echo "SM_LIMIT=144:"; B2U_SM_LIMIT=144 run 29703
echo "NCCL_MAX_CTAS=4:"; B2U_NCCL_MAX_CTAS=4 NCCL_MAX_CTAS=4 run 29704
echo "NCCL_MAX_CTAS=32:"; NCCL_MAX_CTAS=32 run 29705
echo "SM_LIMIT=140 + NCCL_MAX_CTAS=4:"; B2U_SM_LIMIT=140 NCCL_MAX_CTAS=4 run 29706
echo "no overlap:"; B2U_NO_AR_OVERLAP=1 run 29707
echo "N=1 SM_LIMIT=140:"; B2U_SM_LIMIT=140 timeout 300 python bench.py --no-extra --no-predict --no-profile --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3))"
