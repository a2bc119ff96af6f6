#!/bin/bash
# round-1 batch H: division-free triangle pre-test on/off (both bit-exact), then the round's ncu captures
mkdir -p gpurun_out
out=gpurun_out/ab_r01h.txt
: > $out
run() {  # name args...
  name=$1; shift
  lib=""; [ "$name" != default ] && lib=$PWD/ab_$name.so
  echo "== $name $*" >> $out
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 2 --spp 128 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for v in nopre default nopre default; do run $v; done
for v in nopre default; do run $v --config c3 --spp 16; run $v --config c5 --spp 32; done
cat $out
