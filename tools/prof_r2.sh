#!/bin/bash
# round-2 ncu captures at HEAD: K3 m=8 (1M, I=32), K3 m=16 (125k, I=32)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r2_icm8 -f python tools/icm_bench.py 1000000 8 32 1 > gpurun_out/r2_ncu_icm8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r2_icm16 -f python tools/icm_bench.py 125000 16 32 1 > gpurun_out/r2_ncu_icm16.log 2>&1
ls -la gpurun_out | tail -8
