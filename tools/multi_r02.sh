#!/bin/bash
# round 2, multi-GPU session (gpurun --gpus N): cross-device merge tests, the C host with worker threads on N devices, and bench.py at N
N=${1:-2}; TAG=${2:-r02}
export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 ) > gpurun_out/gpu_tests_multi_${TAG}_n$N.txt 2>&1
grep -E "passed|failed|^FAILED|skipped|render_multi|^frame|^total" gpurun_out/gpu_tests_multi_${TAG}_n$N.txt | tail -30
# the reference renderer's structure (one worker thread per tracer, concurrent MergeOutput) on the 4K frame, N devices
python - <<PY 2>&1 | tee gpurun_out/render_multi_${TAG}_n$N.txt
import subprocess, tempfile, os, pathlib, numpy as np
from polaris_b200 import scenes
from tests.test_cpu_abi import _build_c_client
from tests.test_gpu_multi import _cam_args
w, h = 3840, 2160
sc = scenes.build("c5_cornell_4k", w, h)[0]
d = pathlib.Path(tempfile.mkdtemp())
sc.save(str(d / "c5.plrscn"))
exe = _build_c_client(d, "render_multi")
for n in sorted({1, $N}):
    r = subprocess.run([exe, str(d / "c5.plrscn"), str(w), str(h), "64", str(n), "6", str(d / "out.bin")] + _cam_args(sc), capture_output=True, text=True)
    print(r.stdout, r.stderr[-500:])
PY
for x in ipc nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 4 --exchange $x --verbose > gpurun_out/bench_${TAG}_n${N}_$x.json 2> gpurun_out/bench_${TAG}_n${N}_$x.log
  echo "== bench N=$N exchange=$x rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}_n${N}_$x.json | cut -c1-900; grep -E "Error|error|Traceback" gpurun_out/bench_${TAG}_n${N}_$x.log | head -5
done
