#!/bin/bash
# N-GPU A/B of environment switches on the default bench (weak scaling of config 2): each argument is one "VAR=value ..." set
N=${N:-2}
for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-strong --cpu-views 0 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N [$envs]', round(d['value']), 'views/s', round(1e3*d['ms_per_step'],1), 'us/step', d['config'].get('allreduce'))"
done
