#!/bin/bash
# round-1 batch L: prefetch of the far child at push time
mkdir -p gpurun_out
out=gpurun_out/ab_r01l.txt
: > $out
run() {  # name args...
  name=$1; shift
  lib=""; [ "$name" != default ] && lib=$PWD/ab_$name.so
  echo "== $name $*" >> $out
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for v in default prefetch default prefetch; do run $v --spp 128; done
for v in default prefetch; do run $v --config c3 --spp 16; run $v --config c5 --spp 32; done
cat $out
