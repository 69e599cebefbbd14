#!/bin/bash
# round 2, GPU call T: bench lines of c3 and c5 with the single-process clock sampler
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2t}
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_c3.err
timeout 900 python bench.py --preset c5 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --preset c2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --preset c4 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
python - <<PY
import json
for p in ("c3","c5","c2","c4"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%p).read().strip().split("\n")[-1])
        r=d["roofline"]
        print(p, "value %.3g ms/step %.3f dev ms %.3f"%(d["value"], d["ms_per_step"], d["device_ms_per_step"]), "pipe frac %.3f"%r["pipeline_frac"], "dom", r["kernel"], "%.3f"%r["frac"], "clocks", d["clocks"])
        print("   ", " ".join("%s=%.3f(%.2f)"%(k,v["ms"],v["frac"]) for k,v in r["stages"].items()))
        print("    e2e %.3g ms %.2f"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]))
        b=d.get("e2e_bam") or {}
        print("    e2e_bam %.3g  %.2fs all %s parity %s ratio %s"%(b.get("value",0), b.get("seconds",0), b.get("seconds_all"), (b.get("parity") or {}).get("equals_reference_md5"), b.get("ratio_vs_cpu_baseline")))
    except Exception as e:
        print(p, "failed", e)
PY
