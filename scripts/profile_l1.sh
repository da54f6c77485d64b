#!/bin/bash
# GPU box: full capture of the L1 scan kernels for the bench command.
TAG=${1:-r1}
PAT=${2:-l1_thresh_scan}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/prof_plain_l1_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 3 -c 1 \
    -o gpurun_out/l1_$TAG -f $CMD > gpurun_out/ncu_l1_$TAG.log 2>&1
echo "l1 capture rc=$?"
