#!/bin/bash
# round 2, GPU call X: source-level ncu capture of the three big kernels on c2 (SASS page with per-instruction counters)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=${1:-c2}
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_match|k_scan_emit|k_os_pass" -s 9 -c 9 -f -o /tmp/r2x_$P \
    python bench.py --preset $P --steps 1 --warmup 1 --warmup-seconds 0 --resident-only > gpurun_out/r2x_ncu_$P.log 2>&1
echo "capture rc=$?"
for K in k_match k_scan_emit; do
  ncu -i /tmp/r2x_$P.ncu-rep --page source --csv --print-source sass -k regex:$K -c 1 > gpurun_out/r2x_sass_${K}_$P.csv 2>/dev/null
  ncu -i /tmp/r2x_$P.ncu-rep --page source --csv --print-source cuda,sass -k regex:$K -c 1 > gpurun_out/r2x_cudasass_${K}_$P.csv 2>/dev/null
done
ncu -i /tmp/r2x_$P.ncu-rep --page raw --csv > gpurun_out/r2x_raw_$P.csv 2>/dev/null
ls -la gpurun_out/r2x_*; rm -f /tmp/r2x_$P.ncu-rep
