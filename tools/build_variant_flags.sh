#!/bin/bash
# Build a library variant with different floating-point flags (upper-bound experiments, NOT parity-clean):
#   tools/build_variant_flags.sh <name> "<fp flags>" [-DMACRO=VALUE ...]  ->  ab_<name>.so at the repo root
name=$1; fp=$2; shift 2
cd "$(dirname "$0")/../polaris_b200/csrc" || exit 1
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $fp -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -shared -o ../../ab_${name}.so pc_host.cu pc_shade.cu -lcudart 2>&1 | grep -E "error"; ls -la ../../ab_${name}.so
