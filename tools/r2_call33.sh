#!/bin/bash
ROUNDS=3 bash tools/ab_step.sh "default=A=1" "pitch16=B2U_PITCH16=1" "no_pair=B2U_CONV_NO_PAIR=1" "no_pdl=B2U_NO_PDL=1"
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm,clocks.max.mem --format=csv
