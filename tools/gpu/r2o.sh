#!/bin/bash
# round 2, GPU call O: dense phase B of k_scan_emit (A/B), parity; where the teardown time of a c3 run goes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_lean.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_sweep.txt
for P in c2 c5 c4; do
for V in 0 1; do
  PJ_SE_DENSE=$V timeout 300 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - "$P" "$V" >> gpurun_out/${TAG}_sweep.txt <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().split("\n")[-1])
    print(sys.argv[1], "dense", sys.argv[2], "dev %.3f"%d["device_ms_per_step"], " ".join("%s=%.3f"%(k,v["ms"]) for k,v in d["roofline"]["stages"].items()))
except Exception as e:
    print(sys.argv[1:], "failed", e, open("gpurun_out/${TAG}_tmp.err").read()[-300:])
PY
done; done
cat gpurun_out/${TAG}_sweep.txt
portcullis_b200/bin/pjsynth --preset c3 --scale 1 --out /tmp/c3full > /dev/null 2>&1
for i in 1 2; do PJ_TRACE=1 portcullis_b200/bin/portcullis junc -t 16 --gpus 1 -o /tmp/pj_cli/p /tmp/c3full > gpurun_out/${TAG}_cli_c3_$i.log 2>&1; grep -E "pj_destroy|Total runtime" gpurun_out/${TAG}_cli_c3_$i.log; done
