#!/bin/bash
# round 2, closing measurement session (ONE GPU): GPU tests, sanitizers over smoke(), ncu captures of c2 c3 c4, all five configs
# with the full driver contract.  Tag r02b = the kernels after the 8x4 primary tiles, the minimal carve-out, the roulette-aware
# shade sort and the 16 M-path slot target.
mkdir -p gpurun_out
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
bash tools/run_gpu_tests.sh r02b
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_r02b.txt 2>&1
  echo "$tool rc=$?"; tail -1 gpurun_out/sanitizer_${tool}_r02b.txt
done
export PROFILES_OUT=$PWD/gpurun_out/profiles_out
mkdir -p $PROFILES_OUT; cp profiles/traffic.json $PROFILES_OUT/ 2>/dev/null
for c in c2 c3 c4; do
  bash tools/profile_gpu.sh r02b $c 8
  python tools/summarize_ncu.py r02b $c > /dev/null 2>&1
  rm -f gpurun_out/prof_*_r02b_$c.ncu-rep
done
cp $PROFILES_OUT/traffic.json profiles/traffic.json
unset POLARIS_SCENE_CACHE
bash tools/bench_all.sh r02b
