#!/bin/bash
# round 2, GPU call A: new driver (range plan, segments, rank mode) parity + first c3 bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi -L >> gpurun_out/r2a_host.txt
timeout 1200 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "one_process or full_size or presets" > gpurun_out/r2a_scale_tests.log 2>&1
echo "scale tests rc=$?" >> gpurun_out/r2a_scale_tests.log
tail -5 gpurun_out/r2a_scale_tests.log
timeout 600 python bench.py --preset c3 --scale 0.1 --steps 5 > gpurun_out/r2a_bench_c3_0.1.json 2> gpurun_out/r2a_bench_c3_0.1.err
echo "bench c3x0.1 rc=$?"; tail -c 600 gpurun_out/r2a_bench_c3_0.1.err
timeout 1500 python bench.py --steps 10 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
echo "bench c3 rc=$?"; tail -c 600 gpurun_out/r2a_bench_c3.err
cut -c1-1500 gpurun_out/r2a_bench_c3.json
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_scale.py > gpurun_out/r2a_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2a_gpu_tests.log
tail -5 gpurun_out/r2a_gpu_tests.log
