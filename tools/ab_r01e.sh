#!/bin/bash
# round-1 batch E: shade CTA geometry, one GPU.  Output: gpurun_out/ab_r01e.txt (+ ncu capture of k_shade for rpt4b2)
mkdir -p gpurun_out
out=gpurun_out/ab_r01e.txt
: > $out
run() {  # name args...
  name=$1; shift
  lib=""; [ "$name" != default ] && lib=$PWD/ab_$name.so
  echo "== $name $*" >> $out
  POLARIS_CUDA_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 2 --spp 128 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for v in rpt4b2 rpt2b2 rpt3b2 t128r4b4 t128r4b5 t128r8b3 t512r2b1 rpt4b2; do run $v --opt FUSE_TRACE=1; done
for v in rpt4b2 t128r4b4 rpt2; do run $v --opt FUSE_TRACE=1 --config c5 --spp 32; done
for v in default rpt4b2; do run $v --opt FUSE_TRACE=1 --config c4 --spp 8; done
SKIP=4 COUNT=2 tools/prof_variant.sh rpt4b2 k_shade shade_rpt4b2
cat $out
