#!/bin/bash
# A/B sweep of the bench (one GPU): pipeline form x resident footprint-kernel CTAs per SM x library build.
# usage: gpu_sweep.sh tag "pipes" "ctas" "libs(paths or -)"
tag=$1; pipes=${2:-"deep off"}; ctas_list=${3:-"0 3"}; libs=${4:-"-"}
for lib in $libs; do for pipe in $pipes; do for ctas in $ctas_list; do
  name=${tag}_$(basename $lib .so)_${pipe}_c${ctas}
  if [ "$lib" = "-" ]; then unset LP_B200_LIB; else export LP_B200_LIB=$lib; fi
  LP_RASTER_CTAS=$ctas python bench.py --steps 200 --warmup 10 --cpu-views 0 --no-e2e --no-strong --pipeline $pipe $SWEEP_ARGS \
     > gpurun_out/sweep_${name}.json 2> gpurun_out/sweep_${name}.err
done; done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/sweep_${tag}_*.json")):
    try:
        d=json.load(open(f)); print(f.split("sweep_")[1], round(d["ms_per_step"]*1e3,1), "us", {k: round(v,1) for k,v in d["roofline"]["kernels_us"].items()})
    except Exception as e: print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
