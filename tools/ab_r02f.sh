#!/bin/bash
# round 2, GPU session F (ONE GPU): sample slots -- parity tests, then slots 1 / auto on every config
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh r02f -x
for c in c2 c1 c5 c3 c4; do
  for v in "SAMPLE_SLOTS=1" "SAMPLE_SLOTS=0" "SAMPLE_SLOTS=2"; do
    echo "== $c $v"
    timeout 600 python bench.py --config $c --steps 4 --warmup 2 --no-cpu --opt $v 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
  done
done 2>&1 | tee gpurun_out/ab_r02f.txt
