#!/bin/bash
# round 2, GPU call B: kernel changes (64-B pair record, warp look-back, k_match in emit order) — parity + stage times on every preset
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_extra.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
for P in c2 c4 c5; do
  timeout 600 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_bench_$P.json 2> gpurun_out/${TAG}_bench_$P.err
  echo "bench $P rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_$P.err
done
timeout 600 python bench.py --preset c3 --scale 0.25 --steps 10 --resident-only > gpurun_out/${TAG}_bench_c3q.json 2> gpurun_out/${TAG}_bench_c3q.err
echo "bench c3 x0.25 rc=$?"
python - <<PY
import json
for p in ("c2","c4","c5","c3q"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%p).read().strip().split("\n")[-1])
        r=d["roofline"]
        print(p, "dev ms %.3f"%d["device_ms_per_step"], "pipe frac %.3f"%r["pipeline_frac"], " ".join("%s=%.3f"%(k,v["ms"]) for k,v in r["stages"].items()))
    except Exception as e:
        print(p, "failed", e)
PY
