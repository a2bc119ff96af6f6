#!/bin/bash
# ncu captures (whole batches) + all five configs
mkdir -p gpurun_out
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
export PROFILES_OUT=$PWD/gpurun_out/profiles_out
mkdir -p $PROFILES_OUT; cp profiles/traffic.json $PROFILES_OUT/ 2>/dev/null
for c in c2 c3 c4; do
  bash tools/profile_gpu.sh r02 $c 8
  python tools/summarize_ncu.py r02 $c > /dev/null 2>&1
  [ "$c" != c2 ] && rm -f gpurun_out/prof_*_r02_$c.ncu-rep
done
rm -f gpurun_out/prof_k_primary_r02_c2.ncu-rep
cp $PROFILES_OUT/traffic.json profiles/traffic.json
unset POLARIS_SCENE_CACHE
bash tools/bench_all.sh r02
