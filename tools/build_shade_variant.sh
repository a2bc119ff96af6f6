#!/bin/bash
# Library variant that differs only in the SHADING translation unit's floating-point flags:
#   tools/build_shade_variant.sh <name> "<fp flags>"   ->  ab_<name>.so (pc_host.o is the default exact build)
name=$1; fp=$2; shift 2
cd "$(dirname "$0")/../polaris_b200/csrc" || exit 1
[ -f pc_host.o ] || make -s pc_host.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $fp -DPC_SHADE_FP_MODE="\"$name\"" -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c -o /tmp/pc_shade_$name.o pc_shade.cu 2>&1 | grep -E "error"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab_${name}.so pc_host.o /tmp/pc_shade_$name.o -lcudart; ls -la ../../ab_${name}.so
