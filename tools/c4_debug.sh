#!/bin/bash
# config 4 through the default (graph, deep pipeline) bench under a few switches; prints value or the first error line
run() { echo "--- $1"; shift; env "$@" python bench.py --workload c4 --steps 50 --no-e2e --no-strong --cpu-views 0 $ARGS 2>&1 | grep -v CUDAEvent | grep -E "^\{|Error|error|violated" | head -2 | cut -c1-160; }
run "checked library" LP_B200_LIB=$PWD/latent-nerf-test_b200/liblp_b200_checked.so
run "production library (guarded store)" A=1
