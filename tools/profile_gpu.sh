#!/bin/bash
# Run under gpurun (ONE GPU): launch list + one `ncu --set full` capture per hot kernel class of the default
# bench command (config 2).  Outputs land in gpurun_out/; tools/summarize_ncu.py turns them into profiles/*.
set -u
mkdir -p gpurun_out
R=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu --chains 1"
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/launches_${R}.out 2>&1
for K in k_shade k_trace k_occlusion k_primary; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o gpurun_out/prof_${K}_${R} $CMD > gpurun_out/prof_${K}_${R}.out 2>&1
done
ls -la gpurun_out
