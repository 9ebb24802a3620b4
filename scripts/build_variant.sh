#!/bin/bash
# usage: build_variant.sh NAME [-DMACRO=VAL ...]  -> vbuild/libmoire_NAME.so (tuning experiments;
# vbuild/ is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p vbuild
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -I include -o vbuild/libmoire_$name.so moire_b200/csrc/engine.cu moire_b200/csrc/ladder.cu moire_b200/csrc/mcmc_driver.cpp
