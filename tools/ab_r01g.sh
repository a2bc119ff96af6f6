#!/bin/bash
# round-1 batch G: shading-TU floating-point flags (traversal stays IEEE exact); parity tests run against each variant
mkdir -p gpurun_out
out=gpurun_out/ab_r01g.txt
: > $out
run() {  # name args...
  name=$1; shift
  lib=""; [ "$name" != default ] && lib=$PWD/ab_$name.so
  echo "== $name $*" >> $out
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 2 --spp 128 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for v in default sh_nodiv sh_fast sh_fmad; do run $v; done
for v in sh_nodiv sh_fast; do
  echo "== parity tests with $v" >> $out
  POLARIS_CUDA_LIB=$PWD/ab_$v.so timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "worst in-tolerance|pixels beyond|passed|failed|FAILED|sigma|beyond 1e-4" | cut -c1-260 >> $out
done
cat $out
