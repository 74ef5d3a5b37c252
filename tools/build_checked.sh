#!/bin/bash
# Builds the library with -DLP_CHECKED (device-side index / capacity invariants, counted; see tests/test_zz_checked.py)
# next to the production one; run the GPU tests against it with LP_B200_LIB=$PWD/latent-nerf-test_b200/liblp_b200_checked.so
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DLP_CHECKED \
     -I include -o latent-nerf-test_b200/liblp_b200_checked.so latent-nerf-test_b200/csrc/lp_b200.cu && echo latent-nerf-test_b200/liblp_b200_checked.so
