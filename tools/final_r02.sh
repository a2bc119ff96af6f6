#!/bin/bash
# round 2 measurement session (ONE GPU): GPU tests, sanitizers, ncu captures of c2 c3 c4, all five configs with the full driver contract
mkdir -p gpurun_out
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
bash tools/run_gpu_tests.sh r02_final
# compute-sanitizer: smoke() under memcheck / racecheck / synccheck, and memcheck over the tests that exercise this round's new paths
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_r02.txt 2>&1
  echo "$tool rc=$?"; tail -1 gpurun_out/sanitizer_${tool}_r02.txt
done
POLARIS_SKIP_FULL_C4=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
  -k "slots or refilling or deferred or variants_bit_identical or known_answers or random_volumes or small_configs or c_host_worker or ipc_rows or merge_blocks" > gpurun_out/sanitizer_memcheck_tests_r02.txt 2>&1
echo "memcheck tests rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_tests_r02.txt
export PROFILES_OUT=$PWD/gpurun_out/profiles_out
mkdir -p $PROFILES_OUT; cp profiles/traffic.json $PROFILES_OUT/ 2>/dev/null
for c in c2 c3 c4; do
  bash tools/profile_gpu.sh r02 $c 8    # 8 spp with one chain: the same samples per launch as bench.py's roofline pass
  python tools/summarize_ncu.py r02 $c > /dev/null 2>&1   # on the box: the .ncu-rep files are too big to travel back
  [ "$c" != c2 ] && rm -f gpurun_out/prof_*_r02_$c.ncu-rep
done
cp $PROFILES_OUT/traffic.json profiles/traffic.json   # bench.py below reports against these captures
unset POLARIS_SCENE_CACHE   # bench_all: raw scenes stay in memory -> the cold-start (device scene compile) measurement runs
bash tools/bench_all.sh r02
