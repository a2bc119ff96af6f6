#!/bin/bash
# round 2, GPU session E (ONE GPU): full GPU tests, then every config with the automatic k_trace schedule vs forced off
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh r02e
for c in c2 c5 c1 c3 c4; do
  for v in "TRACE_REFILL=-1" "TRACE_REFILL=0"; do
    echo "== $c $v"
    timeout 600 python bench.py --config $c --steps 4 --warmup 2 --no-cpu --opt $v 2>&1 | grep -E "timed|kernel classes|cold start|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's///g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-600
  done
done 2>&1 | tee gpurun_out/ab_r02e.txt
