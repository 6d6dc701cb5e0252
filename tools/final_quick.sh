mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:icm_warp_kernel -s 3 -c 1 -o gpurun_out/r2_icm8 -f python tools/icm_bench.py 1000000 8 32 1 > gpurun_out/r2_ncu_icm8.log 2>&1
cat gpurun_out/r2_pytest_gpu.txt; ls -la gpurun_out
