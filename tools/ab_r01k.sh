#!/bin/bash
# round-1 batch K: traversal-order sort of the bounce rays (PC_OPT_SORT_RAYS) on/off
mkdir -p gpurun_out
out=gpurun_out/ab_r01k.txt
: > $out
run() {
  echo "== $*" >> $out
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu "$@" 2>&1 | grep -E "timed|kernel classes|Error|error|Traceback" | sed -e 's/"alg_GBps": [0-9.]*//g' -e 's/"launches": [0-9]*, //g' | cut -c1-520 >> $out
}
for o in 0 1 0 1; do run --spp 128 --opt SORT_RAYS=$o; done
for o in 0 1; do run --config c5 --spp 32 --opt SORT_RAYS=$o; run --config c1 --opt SORT_RAYS=$o; run --config c3 --spp 16 --opt SORT_RAYS=$o; done
cat $out
