#!/bin/bash
# round 2, GPU session A (ONE GPU): GPU tests of the refactored host + A/B of the traversal-stack / wide-BVH variants on c2 c3 c4 c5
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/gpu_tests_r02a.txt 2>&1
tail -5 gpurun_out/gpu_tests_r02a.txt
for c in c2 c5 c3 c4; do
  for v in default smem8 smem12 smem16 wide widesmem16; do
    lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
    echo "== $c $v"
    POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $c --steps 2 --warmup 2 --spp 64 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's///g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-400
  done
done 2>&1 | tee gpurun_out/ab_r02a.txt
