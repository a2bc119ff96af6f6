#!/bin/bash
# round 2, GPU session D (ONE GPU): device BVH build tests + timing, refill variants vs default on c2 c5 c3 c4
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_bvh_build.py tests/test_gpu_parity.py -m gpu -q -s -k "bvh or compile or variants or deferred or hit_records or chains" 2>&1 ) > gpurun_out/gpu_tests_r02d.txt 2>&1
grep -E "passed|failed|^FAILED|device build" gpurun_out/gpu_tests_r02d.txt | tail -20
python - <<'PY' 2>&1 | tee gpurun_out/bvh_timing_r02d.txt
import time
from polaris_b200 import scenes
from polaris_b200.scene import compile_scene
for name in ("c3_instancing", "c4_terrain"):
    raw = scenes.raw_scene(name)
    for b in ("cuda", "cuda", "host"):
        t = time.time(); sc = compile_scene(raw, 16 / 9, builder=b); dt = time.time() - t
        print(name, b, "compile_scene %.2fs" % dt, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sc.compile_timing.items()})
PY
for c in c2 c5 c3 c4; do
  for v in default refill refill20 refill26s4; do
    lib=""; [ "$v" != default ] && lib=$PWD/ab_$v.so
    echo "== $c $v"
    POLARIS_CUDA_LIB=$lib timeout 600 python bench.py --config $c --steps 2 --warmup 2 --spp 64 --no-cpu 2>&1 | grep -E "timed|kernel classes|Error|error" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's///g' -e 's/"mean_avg_us": [0-9.]*, //g' | cut -c1-400
  done
done 2>&1 | tee gpurun_out/ab_r02d.txt
