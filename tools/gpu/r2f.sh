#!/bin/bash
# round 2, GPU call F: 2-bit read stream + 32-base compare (classic batches repacked on the device): parity, then stage times
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_extra.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -15 gpurun_out/${TAG}_tests.log
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "presets or one_process" > gpurun_out/${TAG}_scale.log 2>&1
echo "scale rc=$?" >> gpurun_out/${TAG}_scale.log; tail -5 gpurun_out/${TAG}_scale.log
: > gpurun_out/${TAG}_sweep.txt
for P in c2 c5 c4; do
for V in 4 5; do
  PJ_MATCH_CTAS=$V timeout 300 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - "$P" "$V" >> gpurun_out/${TAG}_sweep.txt <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().split("\n")[-1])
    print(sys.argv[1], "ctas", sys.argv[2], "dev %.3f"%d["device_ms_per_step"], " ".join("%s=%.3f"%(k,v["ms"]) for k,v in d["roofline"]["stages"].items()))
except Exception as e:
    print(sys.argv[1:], "failed", e, open("gpurun_out/${TAG}_tmp.err").read()[-400:])
PY
done; done
cat gpurun_out/${TAG}_sweep.txt
