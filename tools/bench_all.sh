#!/bin/bash
# All five BASELINE configs on ONE GPU, full driver contract (cpu_baseline + reference_on_gpu included): tools/bench_all.sh <tag>
tag=${1:-r01e}
mkdir -p gpurun_out
for c in c2 c1 c5 c3 c4; do
  timeout 1500 python bench.py --config $c > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.log
  tail -c 1500 gpurun_out/bench_${tag}_$c.json | cut -c1-300
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.log
