#!/bin/bash
# r02j: 8x4 pixel tiles for the primary rays' work queue (PC_PRIMARY_TILES, default build) against the linear order, and the
# traversal kernels' shared-memory carve-out pinned to k_shade's (72 %) / to the minimum (28 %) against the driver's choice.
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
bash tools/run_gpu_tests.sh r02j
for c in c3 c4 c2; do
  for v in default notiles carve72 carve28; do
    lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
    echo "== $c $v"
    POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches_per_batch": [0-9]*, //g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-420
  done
done 2>&1 | tee gpurun_out/ab_r02j.txt
