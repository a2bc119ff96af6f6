#!/bin/bash
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slots or variants or deferred or chains or full_depth or bounce0 or golden" 2>&1 ) | tail -3
for c in c5 c2 c4; do
  for v in "DEFER_OCCLUSION=1" "DEFER_OCCLUSION=0"; do
    echo "== $c $v"
    timeout 600 python bench.py --config $c --steps 4 --warmup 2 --no-cpu --opt $v 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
  done
done 2>&1 | tee gpurun_out/ab_r02i.txt
