#!/bin/bash
# One GPU call: parity tests, then short bench runs of the variants named in $MODES, then the per-kernel profiles.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for mode in ${MODES:-geometry deep}; do
  python bench.py --steps 400 --warmup 20 --no-e2e --cpu-views 0 --pipeline $mode > gpurun_out/bench_$mode.log 2>gpurun_out/bench_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$mode.log").read().strip().splitlines()[-1])
    print("$mode", round(d["value"]), round(1e3*d["ms_per_step"],1), {k:round(v,1) for k,v in d["roofline"]["kernels_us"].items()})
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench_$mode.err").read()[-800:])
PY
done
if [ -n "$SPLIT" ]; then python tools/split_profile.py 2>&1 | tail -1 | tee gpurun_out/split_profile.log; fi
if [ -n "$C4" ]; then python tools/c4_profile.py 5 2>&1 | tail -2 | tee gpurun_out/c4_profile.log; fi
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
