#!/bin/bash
# interleaved A/B of training-step configurations on ONE box: every configuration runs $ROUNDS times, alternating, so
# that box-to-box and warm-up drift cancel.  usage: ab_step.sh "NAME=ENV..." ...   (ENV words are passed to env)
ROUNDS=${ROUNDS:-3}
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%-22s %.3f ms  frac_burst %.4f  sm %s MHz" % (sys.argv[1], d["ms_per_step"], d["roofline"]["frac_burst"], d["clocks"]["sm_mhz"]))'
for r in $(seq $ROUNDS); do
  for cfg in "$@"; do
    name=${cfg%%=*}; envs=${cfg#*=}
    env $envs python bench.py --no-extra --no-predict --no-cpu-baseline --steps 40 --warmup 8 2>/dev/null | python -c "$J" $name
  done
done
