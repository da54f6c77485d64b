#!/bin/bash
# GPU box: compute-sanitizer over scripts/sanitize_smoke.py; logs under gpurun_out/sanitizer_<tool>_<tag>.log
TAG=${1:-r2}
mkdir -p gpurun_out
python scripts/sanitize_smoke.py > gpurun_out/sanitize_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain_$TAG.log; exit 1; }
for tool in memcheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$TAG.log | tail -1)"
done
DCTD_SANITIZE_SMALL=1 timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1
echo "racecheck rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck_$TAG.log | tail -1)"
