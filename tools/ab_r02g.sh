#!/bin/bash
# chains x slots grid on config 2 and config 1
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
for c in c2 c1 c3; do
  for ch in 1 2 4; do for sl in 2 4 8; do
    echo "== $c chains=$ch slots=$sl"
    timeout 600 python bench.py --config $c --steps 4 --warmup 2 --no-cpu --chains $ch --opt SAMPLE_SLOTS=$sl 2>&1 | grep -E "timed|Error|error" | cut -c1-200
  done; done
done 2>&1 | tee gpurun_out/ab_r02g.txt
