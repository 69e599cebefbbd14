#!/bin/bash
# round 2, GPU call E: k_match without staging, register budget sweep (junction order), parity
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_sweep.txt
for P in c2 c5 c4; do
for V in 4 5 3; do
  PJ_MATCH_CTAS=$V timeout 300 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - "$P" "$V" >> gpurun_out/${TAG}_sweep.txt <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().split("\n")[-1])
    print(sys.argv[1], "ctas", sys.argv[2], "dev %.3f"%d["device_ms_per_step"], " ".join("%s=%.3f"%(k,v["ms"]) for k,v in d["roofline"]["stages"].items()))
except Exception as e:
    print(sys.argv[1:], "failed", e)
PY
done; done
cat gpurun_out/${TAG}_sweep.txt
