#!/bin/bash
# round 2, GPU call U: which part of bench.py stalls single steps?  (sampler on/off, page-locked columns on/off)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PJ_BENCH_TRACE=1
for v in "" ""; do
  for p in c2 c5; do
    echo "== $p [$v]"
    env $v timeout 600 python bench.py --preset $p --steps 20 --warmup 5 --resident-only 2>&1 >/dev/null | grep "steps (wall" | cut -c1-900
  done
done 2>&1 | tee gpurun_out/r2u_trace.txt
