#!/bin/bash
# round-1 batch D: shade tile size x fused trace x chains, one GPU.  Output: gpurun_out/ab_r01d.txt
mkdir -p gpurun_out
out=gpurun_out/ab_r01d.txt
: > $out
run() {  # name lib args...
  name=$1; lib=$2; shift 2
  echo "== $name $*" >> $out
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 2 --spp 128 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for v in default rpt1 rpt2 rpt4 rpt4b2; do
  lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
  run $v "$lib" --opt FUSE_TRACE=0
done
run default "" --opt FUSE_TRACE=1
for c in 1 2 3 6 8; do run default "" --opt FUSE_TRACE=1 --chains $c; done
for c in 2 6 8; do run default "" --opt FUSE_TRACE=0 --chains $c; done
run default "" --config c5 --spp 32 --opt FUSE_TRACE=0
run default "" --config c5 --spp 32 --opt FUSE_TRACE=1
run rpt1 "$PWD/ab_rpt1.so" --config c5 --spp 32 --opt FUSE_TRACE=0
run default "" --config c3 --spp 16 --opt FUSE_TRACE=0
run default "" --config c3 --spp 16 --opt FUSE_TRACE=1
cat $out
