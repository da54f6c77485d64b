#!/bin/bash
# GPU box: round-2 captures.  Every ncu run follows a plain run of the same command that exited 0.
# usage: scripts/profile_r2.sh <tag>      (FP_ONLY=1: only the launch list and the two fingerprint-kernel captures)
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu --no-cfg4 --no-allvsall --no-dctsim"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fp_ws_kernel -s 4 -c 1 \
    -o gpurun_out/fp_$TAG -f $CMD > gpurun_out/ncu_fp_$TAG.log 2>&1
echo "fp capture rc=$?"
python scripts/fp_protein.py 2048 4 > gpurun_out/prof_plain_rider_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fp_ws_kernel -s 2 -c 1 \
    -o gpurun_out/fprider_$TAG -f python scripts/fp_protein.py 2048 4 > gpurun_out/ncu_fprider_$TAG.log 2>&1
echo "rider capture rc=$?"
if [ -z "${FP_ONLY:-}" ]; then
python scripts/l1_phases.py stream,ref13 > gpurun_out/prof_plain_stream_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l1_stream_fused -s 1 -c 1 \
    -o gpurun_out/l1stream_$TAG -f python scripts/l1_phases.py stream > gpurun_out/ncu_l1stream_$TAG.log 2>&1
echo "stream capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stream_$TAG.csv \
    python scripts/l1_phases.py stream,ref13 > gpurun_out/ncu_launches_stream_$TAG.log 2>&1
echo "stream launch list rc=$?"
python scripts/dctsim_run.py 4000 > gpurun_out/prof_plain_dctsim_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l1_protein_kernel -s 1 -c 1 \
    -o gpurun_out/l1prot_$TAG -f python scripts/dctsim_run.py 4000 > gpurun_out/ncu_l1prot_$TAG.log 2>&1
echo "protein capture rc=$?"
fi
# summaries are made here: the reports together exceed what gpurun copies back (64 MiB)
for pair in "fp:fp_ws_kernelILi2ELi1280ELb0:fingerprint" "fprider:fp_ws_kernelILi2ELi1280ELb1:fingerprint" \
            "l1stream:l1_stream_fused_kernel:l1topk" "l1prot:l1_protein_kernel:l1topk"; do
  IFS=: read name pat obj <<< "$pair"
  rep=gpurun_out/${name}_$TAG.ncu-rep
  [ -f $rep ] || continue
  ncu -i $rep --page raw --csv > gpurun_out/${name}_${TAG}_raw.csv 2>/dev/null
  python scripts/ncu_lines.py $rep "$pat" $obj 70 > gpurun_out/${name}_${TAG}_lines.txt 2>&1
  rm -f $rep
done
cat gpurun_out/prof_plain_stream_$TAG.log gpurun_out/prof_plain_dctsim_$TAG.log 2>/dev/null | tail -12
ls -la gpurun_out | tail -20
