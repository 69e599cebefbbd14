#!/bin/bash
# round 2, GPU call J (2 GPUs): one process per GPU under torchrun — bench at N=2 on c3 x0.25 and full c3; CLI trace at N=1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2j}
nvidia-smi -L > gpurun_out/${TAG}_host.txt; nproc >> gpurun_out/${TAG}_host.txt; free -g >> gpurun_out/${TAG}_host.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --preset c3 --scale 0.25 > gpurun_out/${TAG}_bench_n2_q.json 2> gpurun_out/${TAG}_bench_n2_q.err
echo "bench N=2 c3x0.25 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_n2_q.err; cut -c1-3000 gpurun_out/${TAG}_bench_n2_q.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --preset c3 --scale 0.25 > gpurun_out/${TAG}_ref_n2_q.json 2> gpurun_out/${TAG}_ref_n2_q.err
echo "reference arm N=2 rc=$?"; cut -c1-600 gpurun_out/${TAG}_ref_n2_q.json
# CLI on the prep dir the bench left behind: where does the wall time of a fresh process go?
D=/tmp/pj_bench/c3_x0.25_s0
PJ_TRACE=1 portcullis_b200/bin/portcullis junc -t 16 --gpus 1 -o /tmp/pj_cli/p $D > gpurun_out/${TAG}_cli_n1.log 2>&1; tail -25 gpurun_out/${TAG}_cli_n1.log
PJ_TRACE=1 portcullis_b200/bin/portcullis junc -t 16 --gpus 2 -o /tmp/pj_cli2/p $D > gpurun_out/${TAG}_cli_n2.log 2>&1; tail -4 gpurun_out/${TAG}_cli_n2.log
cmp /tmp/pj_cli/p.junctions.tab /tmp/pj_cli2/p.junctions.tab && echo "1-GPU and 2-GPU tabs identical"
