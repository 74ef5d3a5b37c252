#!/bin/bash
# One GPU call: parity tests, then short bench runs of the pipeline variants, then the c4 per-kernel profile.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for mode in off geometry raster deep x1; do
  if [ $mode = x1 ]; then export LP_BWD_X1=1; m=geometry; else m=$mode; fi
  python bench.py --steps 400 --warmup 20 --no-e2e --cpu-views 0 --pipeline $m > gpurun_out/bench_$mode.log 2>gpurun_out/bench_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$mode.log").read().strip().splitlines()[-1])
    print("$mode", round(d["value"]), round(1e3*d["ms_per_step"],1), {k:round(v,1) for k,v in d["roofline"]["kernels_us"].items()})
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench_$mode.err").read()[-800:])
PY
done
python tools/split_profile.py 2>&1 | tail -1 | tee gpurun_out/split_profile.log
if [ -n "$C4" ]; then python tools/c4_profile.py 5 2>&1 | tail -2 | tee gpurun_out/c4_profile.log; fi
