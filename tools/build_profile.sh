#!/bin/bash
# Builds the library with -DLP_PROFILE (stop-after-stage switches, per-warp trace of the footprint kernel) next to the
# production one: latent-nerf-test_b200/liblp_b200_profile.so; use it with LP_B200_LIB=<that path>.
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DLP_PROFILE $EXTRA_NVCC \
     -I include -o latent-nerf-test_b200/liblp_b200_profile.so latent-nerf-test_b200/csrc/lp_b200.cu && echo latent-nerf-test_b200/liblp_b200_profile.so
