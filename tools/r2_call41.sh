#!/bin/bash
ROUNDS=3 bash tools/ab_step.sh "back_to_back=A=1" "sync_each_step=B2U_BENCH_SYNC_EACH_STEP=1"
