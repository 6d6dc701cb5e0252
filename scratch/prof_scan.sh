#!/bin/bash
# ncu captures of the scan kernels (one launch each) + clocks during a plain timed run
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > gpurun_out/scan_clocks.csv &
SMI=$!
python scratch/scan_bench.py 1000000 10000 8 1 20 > gpurun_out/scan_plain.log 2>&1
kill $SMI
ncu --set full --clock-control none --import-source on -k regex:scan8_kernel -s 2 -c 1 -o gpurun_out/r2_scan8 python scratch/scan_bench.py 1000000 10000 8 1 1 > gpurun_out/ncu_scan8.log 2>&1
RAYUELA_B200_SCANX8=1 ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r2_scanx8 python scratch/scan_bench.py 1000000 10000 8 1 1 > gpurun_out/ncu_scanx8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r2_scanx16 python scratch/scan_bench.py 1000000 10000 16 1 1 > gpurun_out/ncu_scanx16.log 2>&1
ls -la gpurun_out
