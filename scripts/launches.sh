#!/bin/bash
TAG=${1:-r1}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
