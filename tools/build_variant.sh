#!/bin/bash
# Build a library variant for A/B runs: tools/build_variant.sh <name> [-DMACRO=VALUE ...]  ->  ab_<name>.so at the repo root
# (POLARIS_CUDA_LIB=$PWD/ab_<name>.so selects it)
name=$1; shift
cd "$(dirname "$0")/../polaris_b200/csrc" || exit 1
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -Xcompiler -fopenmp --expt-relaxed-constexpr "$@" -shared -o ../../ab_${name}.so pc_host.cu pc_shade.cu pc_bvh_build.cu -lcudart -lgomp 2>&1 | grep -E "error" ; ls -la ../../ab_${name}.so
