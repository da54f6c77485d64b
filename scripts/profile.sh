#!/bin/bash
# GPU box: launch list + one full capture of each hot kernel, for the same bench command.
# usage: scripts/profile.sh <tag>
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fp_ws_kernel -s 4 -c 2 \
    -o gpurun_out/fp_$TAG -f $CMD > gpurun_out/ncu_fp_$TAG.log 2>&1
echo "fp capture rc=$?"
$CMD > gpurun_out/prof_plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l1_thresh_scan -s 3 -c 1 \
    -o gpurun_out/l1_$TAG -f $CMD > gpurun_out/ncu_l1_$TAG.log 2>&1
echo "l1 capture rc=$?"
ls -la gpurun_out | tail -12
