#!/bin/bash
# One --set full capture (with source) of the kernels matching $KERNEL (regex) in the eager bench; report -> gpurun_out/$OUT.ncu-rep
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"${KERNEL:-k_raster_shade}" --launch-skip ${SKIP:-4} -c ${COUNT:-1} -f -o gpurun_out/${OUT:-prof_one} \
    python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --no-strong --cpu-views 0 ${BENCH_ARGS} > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
