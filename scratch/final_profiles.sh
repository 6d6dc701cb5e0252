#!/bin/bash
# end-of-session evidence: tests, bench line, launch list of the same command, ncu capture of the scan kernel, sanitizers
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r5_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
BENCH_ALLOW_SHORT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ryl|kernel" --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r5_scanx8 python scratch/scan_bench.py 1000000 10000 8 1 1 > /dev/null 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool python scratch/sanitize.py > gpurun_out/r5_san_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r5_san_$tool.log | tail -1) script_completed=$(grep -c 'sanitize run done' gpurun_out/r5_san_$tool.log)"
done > gpurun_out/r5_sanitizer_summary.txt
timeout 300 python bench_rows.py > gpurun_out/r5_rows.json 2> gpurun_out/r5_rows.err
cat gpurun_out/r5_pytest_gpu.txt gpurun_out/r5_sanitizer_summary.txt
