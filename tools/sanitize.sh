#!/bin/bash
# compute-sanitizer over the hot path (run under gpurun, ONE GPU): memcheck / racecheck / synccheck / initcheck on smoke(),
# memcheck on the GPU tests that exercise every launch variant.  Logs: gpurun_out/sanitizer_*.txt (copies in profiles/).
mkdir -p gpurun_out
for t in memcheck racecheck synccheck initcheck; do
  timeout 400 compute-sanitizer --tool $t --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$t.txt 2>&1
  echo "$t rc=$?"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "variants_bit_identical or debug_stages_vs_oracle or merge_blocks or chains" > gpurun_out/sanitizer_memcheck_tests.txt 2>&1
echo "memcheck tests rc=$?"
