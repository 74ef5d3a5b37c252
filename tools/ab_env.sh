#!/bin/bash
# A/B of environment switches on the default bench: each argument is one "VAR=value VAR=value" set (use - for none).
# usage: REPEAT=2 bash tools/ab_env.sh "-" "LP_WALK_CTAS=4" "LP_WALK_CTAS=2 LP_RASTER_CTAS=3"
for rep in $(seq 1 ${REPEAT:-1}); do for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  env $envs python bench.py --steps 200 --warmup 10 --no-e2e --no-strong --cpu-views 0 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$envs]', round(1e3*d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['kernels_us'].items()})"
done; done
