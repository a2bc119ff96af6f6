#!/bin/bash
# A/B timing of library variants on ONE GPU: tools/ab.sh name1 name2 ...  (ab_<name>.so at the repo root,
# "default" = polaris_b200/libpolaris_cuda.so).  Prints the timed line and the per-kernel-class table.
for v in "$@"; do
  echo "== variant $v"
  lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 2 --spp ${AB_SPP:-64} --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-420
done
