#!/bin/bash
# r02o: k_shade's sort phase prefetches each ray's path record into L2 (PC_SHADE_PREFETCH); k_trace at 9 CTAs per SM
# (56 registers, 10-entry shared stack).  The extended ragged-frame test runs first on the default build.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or slots or golden" 2>&1 ) | tail -2
( POLARIS_CUDA_LIB=$PWD/ab_shadepf.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_depth or golden or deterministic" 2>&1 ) | tail -2
run() {  # config variant
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2"
  POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
}
for c in c2 c5 c3; do
  for v in default shadepf trav9; do run $c $v; done
done 2>&1 | tee gpurun_out/ab_r02o.txt
