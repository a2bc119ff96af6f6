#!/bin/bash
# r02n: automatic sample slots aiming at 16 M / 32 M paths per launch instead of 8 M, with 8 and 16 slots at most.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( POLARIS_CUDA_LIB=$PWD/ab_t16m16.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slots or golden or deterministic or full_depth or chains" 2>&1 ) | tail -2
run() {  # config variant
  lib=""; [ "$2" != default ] && lib=$PWD/ab_$2.so
  echo "== $1 $2"
  POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|Error|error" | cut -c1-200
}
for c in c3 c5 c2 c1 c4; do
  for v in default t16 t16m16 t32m16; do run $c $v; done
done 2>&1 | tee gpurun_out/ab_r02n.txt
