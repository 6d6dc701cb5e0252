#!/bin/bash
# round-2 evidence at HEAD: tests, bench line, launch list of the same command, ncu --set full of the dominant kernels
# usage: bash tools/final_profiles.sh [quick]   (quick: skip the K3 captures, which take ~4 minutes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
BENCH_ALLOW_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ryl|kernel" --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > /dev/null 2>&1
if [ "$1" != "quick" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r2_icm8 -f python tools/icm_bench.py 1000000 8 32 1 > gpurun_out/r2_ncu_icm8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r2_icm16 -f python tools/icm_bench.py 125000 16 32 1 > gpurun_out/r2_ncu_icm16.log 2>&1
fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r2_scanx8 -f python tools/scan_bench.py 1000000 10000 8 1 1 > /dev/null 2>&1
RAYUELA_B200_SCAN_PREFILTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scanx_kernel -s 2 -c 1 -o gpurun_out/r2_scanx8_fp32 -f python tools/scan_bench.py 1000000 10000 8 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:unary_tc_kernel -s 1 -c 1 -o gpurun_out/r2_unary_tc -f python tools/unary_bench.py 1000000 8 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:unary_kernel -s 1 -c 1 -o gpurun_out/r2_unary_exact -f python tools/unary_bench.py 1000000 8 > /dev/null 2>&1
timeout 300 python tools/configs01.py > gpurun_out/r2_configs01.json 2> gpurun_out/r2_configs01.err
timeout 400 python bench_rows.py > gpurun_out/r2_rows.json 2> gpurun_out/r2_rows.err
timeout 300 python tools/scan_bench.py 1000000 10000 8,16 1,10,100,1000,4000 5 > gpurun_out/r2_scan_bench.txt 2>&1
{ echo '# RAYUELA_B200_SCAN_PREFILTER=0 (fp32 loop for every k)'; RAYUELA_B200_SCAN_PREFILTER=0 timeout 300 python tools/scan_bench.py 1000000 10000 8,16 1,10,100 5; } >> gpurun_out/r2_scan_bench.txt 2>&1
# gpurun copies back at most 64 MiB: summarise the captures here and keep only the K3 report itself
for r in icm8 icm16 scanx8 scanx8_fp32 unary_tc unary_exact; do
  [ -f gpurun_out/r2_$r.ncu-rep ] && python tools/ncu_metrics.py gpurun_out/r2_$r.ncu-rep > gpurun_out/m_$r.txt 2>/dev/null
done
[ -f gpurun_out/r2_icm8.ncu-rep ] && ncu -i gpurun_out/r2_icm8.ncu-rep --page source --csv > gpurun_out/icm8_source.csv 2>/dev/null
rm -f gpurun_out/r2_icm16.ncu-rep gpurun_out/r2_scanx8.ncu-rep gpurun_out/r2_scanx8_fp32.ncu-rep gpurun_out/r2_unary_tc.ncu-rep gpurun_out/r2_unary_exact.ncu-rep
cat gpurun_out/r2_pytest_gpu.txt; ls -la gpurun_out | grep "r2_\|m_" 
