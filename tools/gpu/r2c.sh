#!/bin/bash
# round 2, GPU call C: k_match with cp.async staging; junction order vs emit order
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
PJ_MATCH_ORDER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${TAG}_tests_emit.log 2>&1
echo "tests(emit order) rc=$?" >> gpurun_out/${TAG}_tests_emit.log
tail -2 gpurun_out/${TAG}_tests_emit.log
for ORD in 0 1; do
for P in c2 c5 c4; do
  PJ_MATCH_ORDER=$ORD timeout 600 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_bench_${P}_o$ORD.json 2> gpurun_out/${TAG}_bench_${P}_o$ORD.err
  echo "bench $P order $ORD rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_${P}_o$ORD.err
done; done
python - <<PY
import json
for o in (0,1):
  for p in ("c2","c5","c4"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s_o%d.json"%(p,o)).read().strip().split("\n")[-1])
        r=d["roofline"]
        print(p, "order",o, "dev ms %.3f"%d["device_ms_per_step"], "pipe frac %.3f"%r["pipeline_frac"], " ".join("%s=%.3f"%(k,v["ms"]) for k,v in r["stages"].items()))
    except Exception as e:
        print(p, o, "failed", e)
PY
