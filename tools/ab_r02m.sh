#!/bin/bash
# r02m: k_shade's sort key foresees the Russian-roulette victims (PC_SORT_RR, default build) and, as a variant, the LEAF of a
# layered material (PC_SORT_LEAF), against the plain material-root key.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
for v in default sortleaf; do
  lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
  ( POLARIS_CUDA_LIB=$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_depth or golden or slots or deterministic or bounce_count or chains" 2>&1 ) | tail -2
done
run() {  # config variant
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2"
  POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
}
for c in c2 c5 c4 c3; do
  for v in nosortrr default sortleaf; do run $c $v; done
done 2>&1 | tee gpurun_out/ab_r02m.txt
