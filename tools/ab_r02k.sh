#!/bin/bash
# r02k: (default = 8x4 primary tiles + minimal traversal carve-out) against 16x8 super-tiles, a 5-/6-entry shared stack with a
# 32 KB carve-out (224 KB L1), and the L2 fetch granularity hint.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hit_records or golden or ragged or bounce0 or deterministic" 2>&1 ) | tail -3
run() {  # config variant [env]
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2 $3"
  env $3 POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
}
for c in c3 c4 c2; do
  for v in default tiles2 stack5c14 stack6c14; do run $c $v X=1; done
done 2>&1 | tee gpurun_out/ab_r02k.txt
for c in c4 c3; do
  for g in 32 128; do run $c default PC_L2_FETCH=$g; done
done 2>&1 | tee -a gpurun_out/ab_r02k.txt
