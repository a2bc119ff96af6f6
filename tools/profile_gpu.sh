#!/bin/bash
# Run under gpurun (ONE GPU): launch list + one `ncu --set full` capture per hot kernel class of a bench command.
#   tools/profile_gpu.sh <tag> [config=c2] [spp=4]
# Outputs land in gpurun_out/ (launches_<tag>_<config>.csv, prof_<kernel>_<tag>_<config>.ncu-rep);
# tools/summarize_ncu.py <tag> <config> turns them into profiles/*.  Numbers printed by bench.py under ncu are never bench values.
set -u
mkdir -p gpurun_out
R=${1:-r02}
C=${2:-c2}
SPP=${3:-4}
export POLARIS_SCENE_CACHE=${POLARIS_SCENE_CACHE:-/tmp/polaris_scenes}
CMD="python bench.py --config $C --steps 1 --warmup 1 --spp $SPP --no-cpu --chains 1"
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}_${C}.csv $CMD > gpurun_out/launches_${R}_${C}.out 2>&1
# every launch of a class in ONE batch of samples, the third batch of the run (warm caches): k_trace has 4 launches per batch,
# k_shade 5, k_primary 1 (2 captured) -- the per-launch means then describe the same launches bench.py's roofline pass times
for KS in k_trace:8:4 k_shade:10:5 k_primary:2:2; do
  K=${KS%%:*}; REST=${KS#*:}; SKIP=${REST%%:*}; CNT=${REST##*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/prof_${K}_${R}_${C} $CMD > gpurun_out/prof_${K}_${R}_${C}.out 2>&1
done
ls -la gpurun_out | grep "${R}_${C}"
