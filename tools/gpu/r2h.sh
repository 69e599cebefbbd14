#!/bin/bash
# round 2, GPU call H: filt feature extraction parity; effect of the nvidia-smi sampler on the e2e arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 900 python -m pytest tests/test_gpu_features.py -x -q > gpurun_out/${TAG}_features.log 2>&1
echo "features rc=$?" >> gpurun_out/${TAG}_features.log; tail -25 gpurun_out/${TAG}_features.log
for IV in 0.2 2.0; do
  PJ_BENCH_SMI_INTERVAL=$IV timeout 600 python bench.py --preset c2 --steps 20 --no-bam --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_iv$IV.json 2> gpurun_out/${TAG}_bench_c2_iv$IV.err
  python - "$IV" <<PY
import json,sys
d=json.loads(open("gpurun_out/${TAG}_bench_c2_iv%s.json"%sys.argv[1]).read().strip().split("\n")[-1])
print("smi interval", sys.argv[1], "value ms %.3f"%d["ms_per_step"], "dev %.3f"%d["device_ms_per_step"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], "clock samples", d["clocks"]["samples"])
PY
done
