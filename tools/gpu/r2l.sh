#!/bin/bash
# round 2, GPU call L: --extra paths (CTA-wide cap replay, batched coverage) parity + timing on c4; NUMA layout of the box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2l}
( nvidia-smi topo -m; echo; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist)"; done; echo "allowed cpus: $(grep Cpus_allowed_list /proc/self/status)"; echo "allowed mems: $(grep Mems_allowed_list /proc/self/status)"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa_node $(cat $d/numa_node)"; fi; done; which numactl ) > gpurun_out/${TAG}_numa.txt 2>&1
cat gpurun_out/${TAG}_numa.txt | head -40
timeout 900 python -m pytest tests/test_gpu_extra.py tests/test_oracle_extra.py -x -q > gpurun_out/${TAG}_extra_tests.log 2>&1
echo "extra tests rc=$?" >> gpurun_out/${TAG}_extra_tests.log; tail -4 gpurun_out/${TAG}_extra_tests.log
timeout 900 python tools/scale_check.py --preset c4 --scale 1.0 --extra --threads 16 > gpurun_out/${TAG}_c4_extra.json 2> gpurun_out/${TAG}_c4_extra.err; cat gpurun_out/${TAG}_c4_extra.json | cut -c1-1200
timeout 300 python tools/run_extra_once.py --preset c4 --scale 1.0 > gpurun_out/${TAG}_c4_extra_stages.txt 2>&1; tail -12 gpurun_out/${TAG}_c4_extra_stages.txt
