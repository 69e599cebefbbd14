#!/bin/bash
# round 2, GPU call D: k_match variant sweep (staging mode x register budget), junction order
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2d}
: > gpurun_out/${TAG}_sweep.txt
for P in c2 c5; do
for V in "0 5" "0 4" "0 3" "1 5" "1 4" "1 3" "2 3" "2 4" "2 2"; do
  set -- $V
  PJ_MATCH_STAGE=$1 PJ_MATCH_CTAS=$2 timeout 300 python bench.py --preset $P --steps 10 --resident-only > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - "$P" "$1" "$2" >> gpurun_out/${TAG}_sweep.txt <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().split("\n")[-1])
    print(sys.argv[1], "stage", sys.argv[2], "ctas", sys.argv[3], "match %.3f"%d["roofline"]["stages"]["match"]["ms"], "dev %.3f"%d["device_ms_per_step"])
except Exception as e:
    print(sys.argv[1:], "failed", e)
PY
done; done
cat gpurun_out/${TAG}_sweep.txt
