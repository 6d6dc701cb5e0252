#!/bin/bash
# end-of-session evidence: bench line, launch list of the same command, ncu captures of the scan kernels
mkdir -p gpurun_out
python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
python bench.py --path linscan --k 1000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_linscan_k1000.json 2>> gpurun_out/r3_bench.err
BENCH_ALLOW_SHORT=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ryl|kernel" --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r4_scanx8 python scratch/scan_bench.py 1000000 10000 8 1 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r4_scanx16 python scratch/scan_bench.py 1000000 10000 16 1 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
