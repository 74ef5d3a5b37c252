#!/bin/bash
# Quick GPU check of a kernel change: parity tests (production and -DLP_CHECKED library when built), then the default
# bench without the host / CPU legs, then whatever $EXTRA names.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
if [ -f latent-nerf-test_b200/liblp_b200_checked.so ] && [ -n "$CHECKED" ]; then
  LP_B200_LIB=$PWD/latent-nerf-test_b200/liblp_b200_checked.so python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_checked.log
fi
python bench.py --steps 200 --warmup 10 --no-e2e --no-strong --cpu-views 0 $BENCH_ARGS > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
    print(round(d["value"]), "views/s", round(1e3*d["ms_per_step"],1), "us/step", {k:round(v,1) for k,v in d["roofline"]["kernels_us"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_quick.err").read()[-1500:])
PY
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
