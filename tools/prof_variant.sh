#!/bin/bash
# ncu --set full capture of one kernel class for a library variant: tools/prof_variant.sh <variant|default> <kernel regex> <tag>
v=$1; k=$2; tag=$3
lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
mkdir -p gpurun_out
POLARIS_CUDA_LIB=$lib ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/prof_${tag} python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu > gpurun_out/prof_${tag}.out 2>&1
