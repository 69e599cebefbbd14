#!/bin/bash
# round 2, GPU call M: ncu launch lists and --set full captures of the pipeline kernels on every preset (B200_PROFILING.md recipe)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2}
KREG='regex:k_match|k_os_pass|k_scan_emit|k_reduce1|k_reduce2|k_flag_scan|k_entropy_sum|k_finalize|k_os_hist|k_junc_init'
for P in c2 c4 c5 c3; do
  SC=1.0
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$P.csv \
      python bench.py --preset $P --scale $SC --steps 2 --warmup 3 --resident-only > gpurun_out/${TAG}_ncu_launch_$P.log 2>&1
  echo "launch list $P rc=$?"
  timeout 1500 ncu --set full --clock-control none --import-source on -k "$KREG" -s 42 -c 32 -f -o gpurun_out/${TAG}_prof_$P \
      python bench.py --preset $P --scale $SC --steps 1 --warmup 3 --resident-only > gpurun_out/${TAG}_ncu_full_$P.log 2>&1
  echo "full capture $P rc=$?"; ls -la gpurun_out/${TAG}_prof_$P.ncu-rep 2>/dev/null
done
