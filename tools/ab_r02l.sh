#!/bin/bash
# r02l: pop-time culling of stacked far children by their entry distance (PC_POP_CULL, two stack words per closest-hit entry),
# with the 12-word and a 16-word shared stack, against the default build.  Parity subset runs on the variant first.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( POLARIS_CUDA_LIB=$PWD/ab_popcull.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hit_records or golden or ragged or bounce0 or refill" 2>&1 ) | tail -3
run() {  # config variant
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2"
  POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
}
for c in c3 c4 c2; do
  for v in default popcull popcull16; do run $c $v; done
done 2>&1 | tee gpurun_out/ab_r02l.txt
