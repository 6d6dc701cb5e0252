#!/bin/bash
mkdir -p gpurun_out
RAYUELA_B200_ICM_USM=0 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r3_icm_pf python scratch/icm_bench.py 1000000 8 32 1 > gpurun_out/ncu_icm.log 2>&1
ls -la gpurun_out | tail -3
