#!/bin/bash
# compute-sanitizer over scratch/sanitize_case.py; summaries into gpurun_out/<tag>_sanitizer_*.log
TAG=$1
mkdir -p gpurun_out
timeout 120 python scratch/sanitize_case.py all > gpurun_out/${TAG}_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scratch/sanitize_case.py all \
    > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|done" gpurun_out/${TAG}_sanitizer_${tool}.log | head -12
done
