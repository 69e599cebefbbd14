#!/bin/bash
# round 2, GPU call Q: smoke, the whole GPU test-suite, the bench lines of every preset (kept under profiles/), memcheck of the new kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2q}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${TAG}_gpu_tests.log; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_c3.json 2> gpurun_out/${TAG}_ref_c3.err
for P in c2 c4 c5; do
  timeout 900 python bench.py --preset $P --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$P.json 2> gpurun_out/${TAG}_bench_$P.err; echo "bench $P rc=$?"; tail -c 200 gpurun_out/${TAG}_bench_$P.err
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_lean.py tests/test_gpu_features.py tests/test_gpu_parity.py tests/test_gpu_extra.py -x -q -m gpu -k "not reference_files and not strand_analysis and not synthetic_presets" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/${TAG}_memcheck.log
python - <<PY
import json
for p in ("c3","c2","c4","c5"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%p).read().strip().split("\n")[-1])
        r=d["roofline"]
        print(p, "value %.3g dev ms %.3f"%(d["value"], d["device_ms_per_step"]), "pipe frac %.3f"%r["pipeline_frac"], "dom", r["kernel"], "%.3f"%r["frac"], "traffic", r["traffic"])
        print("   ", " ".join("%s=%.3f(%.2f)"%(k,v["ms"],v["frac"]) for k,v in r["stages"].items()))
        print("    e2e %.3g ms %.2f h2d %.0f MB"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]/1e6))
        b=d.get("e2e_bam") or {}
        print("    e2e_bam %.3g  %.2fs %s parity %s ratio %s"%(b.get("value",0), b.get("seconds",0), b.get("breakdown_s_rank0"), (b.get("parity") or {}).get("equals_reference_md5"), b.get("ratio_vs_cpu_baseline")))
        print("    cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("sample","")[:120])
    except Exception as e:
        print(p, "failed", e)
PY
