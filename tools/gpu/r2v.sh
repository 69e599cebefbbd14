#!/bin/bash
# round 2, GPU call V: bisect of the stalled first steps (tools/stall_probe.py with PJ_TRACE_HOST=1), then bench traces and the GPU tests with the arena allocator
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PJ_TRACE_HOST=1
for f in "" tswfr ""; do timeout 300 python tools/stall_probe.py c5 "$f" 2>&1 | grep -v "^\[pj\] \|Warning" | tail -12; done | tee gpurun_out/r2v_stall.txt
unset PJ_TRACE_HOST
export PJ_BENCH_TRACE=1
for p in c2 c5; do timeout 600 python bench.py --preset $p --steps 20 --warmup 5 --resident-only 2>&1 >gpurun_out/r2v_bench_$p.json | grep "steps (wall" | cut -c1-700; done | tee -a gpurun_out/r2v_stall.txt
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2v_tests.txt
