#!/bin/bash
# usage: build_variant.sh NAME [-DMACRO=VAL ...]  -> gpurun_out/variants/libmoire_NAME.so (tuning experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -I include -o variants/libmoire_$name.so moire_b200/csrc/engine.cu moire_b200/csrc/mcmc_driver.cpp
