#!/bin/bash
# compute-sanitizer over tools/sanitize.py; summary -> gpurun_out/r2_sanitizer_summary.txt
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r2_san_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_san_$tool.log | tail -1) script_completed=$(grep -c 'sanitize run done' gpurun_out/r2_san_$tool.log)"
done > gpurun_out/r2_sanitizer_summary.txt
cat gpurun_out/r2_sanitizer_summary.txt
