#!/bin/bash
# N=8: how many CTAs NCCL gets, with and without overlapping the all-reduce with the backward segments
mkdir -p gpurun_out
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d.get("clocks",{}); print("%-34s %.3f ms/step  %.0f tiles/s  sm %s MHz" % (sys.argv[1], d["ms_per_step"], d["value"], c.get("sm_mhz")))'
LIGHT="--no-extra --no-predict --no-profile --no-cpu-baseline --steps 30 --warmup 6"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29700
run() { name=$1; shift; port=$((port+1)); env "$@" timeout 300 $TR --master-port $port bench.py --gpus 8 $LIGHT 2>/dev/null | python -c "$J" "$name" | tee -a gpurun_out/r02_n8_nccl_ctas.txt; }
rm -f gpurun_out/r02_n8_nccl_ctas.txt
run "skip all-reduce (diagnostic)" B2U_DIAG_SKIP_ALLREDUCE=1
run "8 CTAs, overlap (default)" A=1
run "16 CTAs, overlap" NCCL_MAX_CTAS=16
run "32 CTAs, overlap" NCCL_MAX_CTAS=32
run "32 CTAs, no overlap" NCCL_MAX_CTAS=32 B2U_NO_AR_OVERLAP=1
run "64 CTAs, no overlap" NCCL_MAX_CTAS=64 B2U_NO_AR_OVERLAP=1
run "16 CTAs, no overlap" NCCL_MAX_CTAS=16 B2U_NO_AR_OVERLAP=1
