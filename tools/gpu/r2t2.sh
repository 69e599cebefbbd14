#!/bin/bash
# round 2, last GPU call: bench lines of c5, c2 and c4 with the final library
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2fin}
for P in c5 c2 c4; do
  timeout 170 python bench.py --preset $P --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$P.json 2> gpurun_out/${TAG}_bench_$P.err; echo "bench $P rc=$?"
done
python - <<PY
import json
for p in ("c5","c2","c4"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%p).read().strip().split("\n")[-1])
        b=d.get("e2e_bam") or {}
        print(p, "value %.3g ms %.3f"%(d["value"], d["ms_per_step"]), "e2e %.3g"%d["e2e"]["value"], "e2e_bam %.3g %.2fs parity %s ratio %s"%(b.get("value",0), b.get("seconds",0), (b.get("parity") or {}).get("equals_reference_md5"), b.get("ratio_vs_cpu_baseline")), "clocks", d["clocks"])
    except Exception as e:
        print(p, "failed", e)
PY
