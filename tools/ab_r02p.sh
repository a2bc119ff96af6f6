#!/bin/bash
# r02p: k_shade with 320-thread CTAs (10 warps, 96 registers, 960-ray tiles, 2 CTAs per SM = 20 warps) against 256 threads / 128 registers.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( POLARIS_CUDA_LIB=$PWD/ab_shade320.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_depth or golden or deterministic or ragged" 2>&1 ) | tail -2
run() {  # config variant
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2"
  POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
}
for c in c2 c5; do
  for v in default shade320; do run $c $v; done
done 2>&1 | tee gpurun_out/ab_r02p.txt
