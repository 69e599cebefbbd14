#!/bin/bash
# round 2, GPU call Y: GPU tests + c3 bench line after the host-side changes (inflate, writers, finalize, background teardown)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2y}
timeout 1500 python -m pytest ${TESTS:-tests} -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.txt
PJ_TRACE=1 timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; grep "pj writer" gpurun_out/${TAG}_bench_c3.err | tail -4
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().split("\n")[-1])
print("c3 value %.3g ms/step %.3f dev %.3f"%(d["value"], d["ms_per_step"], d["device_ms_per_step"]), "e2e %.3g ms %.1f"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]))
b=d["e2e_bam"]; print("e2e_bam %.3g %.2fs all %s parity %s ratio %s"%(b["value"], b["seconds"], b["seconds_all"], b["parity"]["equals_reference_md5"], b.get("ratio_vs_cpu_baseline"))); print(b["breakdown_s_rank0"])
PY
