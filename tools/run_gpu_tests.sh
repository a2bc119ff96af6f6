#!/bin/bash
# full GPU test suite under gpurun: tools/run_gpu_tests.sh <tag> [pytest args]
tag=${1:-r02}; shift
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -s "$@" 2>&1 ) > gpurun_out/gpu_tests_${tag}.txt 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/gpu_tests_${tag}.txt | tail -40
