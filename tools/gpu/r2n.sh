#!/bin/bash
# round 2, GPU call N: genome upload overlapped with later segments; driver parity; c3 / c2 bench lines
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "presets or one_process or full_size" > gpurun_out/${TAG}_scale.log 2>&1
echo "scale rc=$?" >> gpurun_out/${TAG}_scale.log; tail -4 gpurun_out/${TAG}_scale.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench c3 rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_c3.json 2> gpurun_out/${TAG}_ref_c3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().split("\n")[-1])
r=d["roofline"]
print("c3 value %.3g dev ms %.3f"%(d["value"], d["device_ms_per_step"]), "pipe frac %.3f"%r["pipeline_frac"], "dom", r["kernel"], "%.3f"%r["frac"], "traffic", r["traffic"])
print("   ", " ".join("%s=%.3f(%.2f)"%(k,v["ms"],v["frac"]) for k,v in r["stages"].items()))
print("    e2e %.3g ms %.2f h2d %.0f MB"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]/1e6))
print("    e2e_bam", json.dumps(d.get("e2e_bam"))[:900])
print("    cpu", d.get("cpu_baseline"))
print(open("gpurun_out/${TAG}_ref_c3.json").read()[:400])
PY
