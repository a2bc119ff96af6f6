#!/bin/bash
# round 2 measurement session (ONE GPU): ncu captures of c2 c3 c4, all five configs with the full driver contract, sanitizers
mkdir -p gpurun_out
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
export PROFILES_OUT=$PWD/gpurun_out/profiles_out
mkdir -p $PROFILES_OUT; cp profiles/traffic.json $PROFILES_OUT/ 2>/dev/null
for c in c2 c3 c4; do
  bash tools/profile_gpu.sh r02 $c 4
  python tools/summarize_ncu.py r02 $c > /dev/null 2>&1   # on the box: the .ncu-rep files are too big to travel back
  [ "$c" != c2 ] && rm -f gpurun_out/prof_*_r02_$c.ncu-rep
done
unset POLARIS_SCENE_CACHE   # bench_all: raw scenes stay in memory -> the cold-start (device scene compile) measurement runs
bash tools/bench_all.sh r02
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_r02.txt 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_r02.txt
done
