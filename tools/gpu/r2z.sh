#!/bin/bash
# round 2, GPU call Z: where does the BAM-to-files run of c3 spend its host time?  (decode workers against the in-order consumer)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PREP=$(python - <<PY
import bench
prep, meta = bench.make_workload("c3", 1.0, 0, 16)
print(prep)
PY
)
echo "prep $PREP"; nproc; lscpu | grep -i "model name\|^CPU(s)\|Thread\|MHz" | head -6
for TH in 16 24 12; do
  for rep in 1 2; do
    PJ_TRACE=1 portcullis_b200/bin/portcullis junc -t $TH -o /tmp/pj_out_$TH/x $PREP 2>&1 | grep -i "pj pipeline\|wall time\|decode+H2D\|pj writer" | cut -c1-400
  done
done 2>&1 | tee gpurun_out/r2z_host.txt
