#!/bin/bash
# bench.py at N GPUs (gpurun --gpus N): tools/bench_multi.sh <N> <tag> [exchange ...]
N=${1:-2}; TAG=${2:-r02}; shift; shift
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
for x in ${@:-ipc}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 4 --exchange $x --verbose $BENCH_EXTRA > gpurun_out/bench_${TAG}_n${N}_$x.json 2> gpurun_out/bench_${TAG}_n${N}_$x.log
  echo "== bench N=$N exchange=$x rc=$?"; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_${TAG}_n${N}_$x.json"))
    print({k: j[k] for k in ("value", "ms_per_step", "device_ms_per_step", "rows_last_step", "one_gpu_same_workload", "speedup_vs_one_gpu_same_workload", "e2e", "exchange") if k in j})
except Exception as e:
    print("no JSON line:", e)
PY
  grep -E "Error|error|Traceback" gpurun_out/bench_${TAG}_n${N}_$x.log | head -5; grep -E "pass 1[0-9]:" gpurun_out/bench_${TAG}_n${N}_$x.log | tail -3
done
