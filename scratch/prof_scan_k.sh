#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 4 -c 1 -o gpurun_out/r5_scanx8_k1000 python scratch/scan_bench.py 1000000 10000 8 1000 1 > /dev/null 2>&1
