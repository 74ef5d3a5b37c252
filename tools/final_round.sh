#!/bin/bash
# End-of-round evidence on one GPU: parity tests, the default bench line, the ncu launch list of the same command
# (eager launches) and one --set full capture of every kernel of the step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.log
python bench.py --pipeline off --no-e2e --cpu-views 0 > gpurun_out/bench_serial.log 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --cpu-views 0 > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_" --launch-skip 42 -c 14 -f -o gpurun_out/prof_final \
    python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --cpu-views 0 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
