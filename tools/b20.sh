#!/bin/bash
# the driver's form of the bench (20 steps, 3 warm-up), a few times: run-to-run spread of a 1.3 ms timed region
for rep in 1 2 3 4 5; do python bench.py --steps 20 --warmup 3 --no-e2e --no-strong --cpu-views 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20 steps:', round(1e3*d['ms_per_step'],1), 'us/step', d['clocks'])"; done
