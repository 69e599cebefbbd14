#!/bin/bash
# round 2, GPU call K: genome exception summary bitmap + background pool trim: parity, c3 / c2 bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_lean.py tests/test_gpu_features.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_tests.log
timeout 1500 python bench.py --steps 10 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench c3 rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_c3.err
timeout 600 python bench.py --preset c2 --steps 20 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
python - <<PY
import json
for p in ("c3","c2"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%p).read().strip().split("\n")[-1])
        r=d["roofline"]
        print(p, "value %.3g dev ms %.3f"%(d["value"], d["device_ms_per_step"]), "pipe frac %.3f"%r["pipeline_frac"], "dom", r["kernel"], "%.3f"%r["frac"])
        print("   ", " ".join("%s=%.3f(%.2f)"%(k,v["ms"],v["frac"]) for k,v in r["stages"].items()))
        print("    e2e %.3g ms %.2f h2d %.0f MB"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]/1e6))
        print("    e2e_bam", json.dumps(d.get("e2e_bam"))[:600])
    except Exception as e:
        print(p, "failed", e)
PY
