#!/bin/bash
# round 2, GPU session B (ONE GPU): full GPU test suite, A/B of the deferred last occlusion launch, ncu captures for c2 c3 c4
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/gpu_tests_r02b.txt 2>&1
tail -15 gpurun_out/gpu_tests_r02b.txt
for c in c2 c5 c1 c3 c4; do
  for v in "DEFER_OCCLUSION=1" "DEFER_OCCLUSION=0"; do
    echo "== $c $v"
    timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu --opt $v 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's///g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-400
  done
done 2>&1 | tee gpurun_out/ab_r02b.txt
for c in c2 c3 c4; do bash tools/profile_gpu.sh r02 $c 4; done
