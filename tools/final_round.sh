#!/bin/bash
# End-of-round evidence on one GPU: parity tests (production and -DLP_CHECKED library), smoke, the bench lines of every
# config, the ncu launch list of the default command (eager launches) and one --set full capture of every kernel of a step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
if [ -f latent-nerf-test_b200/liblp_b200_checked.so ]; then
  LP_B200_LIB=$PWD/latent-nerf-test_b200/liblp_b200_checked.so python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_checked.log
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_20.log 2> gpurun_out/bench_20.err; tail -c 300 gpurun_out/bench_20.err
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 1200 gpurun_out/bench_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>/dev/null
python bench.py --pipeline off --no-e2e --no-strong --cpu-views 0 > gpurun_out/bench_serial.log 2>/dev/null
for w in c3 c4; do python bench.py --workload $w --no-e2e --no-strong --cpu-views 0 --steps 80 > gpurun_out/bench_$w.log 2>gpurun_out/bench_$w.err; done
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --no-strong --cpu-views 0 > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_" --launch-skip ${NCU_SKIP:-30} -c ${NCU_COUNT:-12} -f -o gpurun_out/prof_final \
    python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --no-strong --cpu-views 0 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
