#!/bin/bash
# ncu captures of the scan kernel (one launch each: the 4-wave main launch of a 10k-query search)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r3_scanx8 python scratch/scan_bench.py 1000000 10000 8 1 1 > gpurun_out/ncu_scanx8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r3_scanx16 python scratch/scan_bench.py 1000000 10000 16 1 1 > gpurun_out/ncu_scanx16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r3_scanx8_k1000 python scratch/scan_bench.py 1000000 10000 8 1000 1 > gpurun_out/ncu_scanx8k.log 2>&1
ls -la gpurun_out | tail -5
