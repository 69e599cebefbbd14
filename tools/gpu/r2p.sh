#!/bin/bash
# round 2, GPU call P (N GPUs): the whole c3 job under torchrun, one process per GPU
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-4}; TAG=${2:-r2p}
nvidia-smi -L > gpurun_out/${TAG}_host_n$N.txt; nproc >> gpurun_out/${TAG}_host_n$N.txt; free -g >> gpurun_out/${TAG}_host_n$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 800 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().split("\n")[-1])
r=d["roofline"]
print("N=%d value %.3g ms/step %.3f dev max %.3f by rank %s"%(d["n_gpus"], d["value"], d["ms_per_step"], d["device_ms_per_step"], d["device_ms_per_step_by_rank"]))
print("   stats", d["workload_stats"])
print("   e2e %.3g ms %.2f h2d %.0f MB"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]/1e6))
print("   e2e_bam", json.dumps(d.get("e2e_bam"))[:900])
print("   setup", d["setup"])
PY
