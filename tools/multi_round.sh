#!/bin/bash
# N-GPU bench lines (torchrun, one rank per GPU): usage N=4 bash tools/multi_round.sh [extra bench args]
mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --cpu-views 0 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N', round(d['value']), 'views/s', round(1e3*d['ms_per_step'],1), 'us/step', d['config'].get('allreduce'), {k:round(v,1) for k,v in d['roofline']['kernels_us'].items()})
s=d.get('strong_scaling')
if s: print('  strong c3:', round(s['value']), 'views/s', round(1e3*s['ms_per_step'],1), 'us/step', s.get('allreduce'))
" || tail -c 1500 gpurun_out/bench_n$N.err
