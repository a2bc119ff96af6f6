#!/bin/bash
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
bash tools/run_gpu_tests.sh r02h
for c in c3 c4; do
  for v in default rf16_8 rf12_8 rf20_12 rf24_8; do
    lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
    echo "== $c $v"
    POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|Error|error" | cut -c1-200
  done
done 2>&1 | tee gpurun_out/ab_r02h.txt
