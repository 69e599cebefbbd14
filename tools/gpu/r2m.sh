#!/bin/bash
# round 2, GPU call M: ncu launch lists and --set full captures of the pipeline kernels on every preset (B200_PROFILING.md recipe).
# The .ncu-rep files are summarised ON THE BOX (raw + source pages as CSV) and deleted: gpurun only brings back 64 MiB.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2}
KREG='regex:k_match|k_os_pass|k_scan_emit|k_reduce1|k_reduce2|k_flag_scan|k_entropy_sum|k_finalize|k_os_hist|k_junc_init'
for P in c2 c4 c5 c3; do
  SC=1.0; [ $P = c3 ] && SC=0.25
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$P.csv \
      python bench.py --preset $P --scale $SC --steps 2 --warmup 3 --resident-only > gpurun_out/${TAG}_ncu_launch_$P.log 2>&1
  echo "launch list $P rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -s 60 -c 16 -f -o /tmp/${TAG}_prof_$P \
      python bench.py --preset $P --scale $SC --steps 1 --warmup 3 --resident-only > gpurun_out/${TAG}_ncu_full_$P.log 2>&1
  echo "full capture $P rc=$?"
  ncu -i /tmp/${TAG}_prof_$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_$P.csv 2>/dev/null
  if [ $P = c2 ] || [ $P = c5 ]; then
    ncu -i /tmp/${TAG}_prof_$P.ncu-rep --page source --csv --print-source cuda -k regex:k_match > gpurun_out/${TAG}_src_match_$P.csv 2>/dev/null
    ncu -i /tmp/${TAG}_prof_$P.ncu-rep --page source --csv --print-source cuda -k regex:k_scan_emit > gpurun_out/${TAG}_src_scan_emit_$P.csv 2>/dev/null
  fi
  rm -f /tmp/${TAG}_prof_$P.ncu-rep
  ls -la gpurun_out/${TAG}_raw_$P.csv
done
du -sh gpurun_out
