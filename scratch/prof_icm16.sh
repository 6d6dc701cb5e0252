#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r4_icm16 python scratch/icm_bench.py 125000 16 32 1 > gpurun_out/ncu_icm16.log 2>&1
