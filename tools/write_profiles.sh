#!/bin/bash
# turns the files tools/final_profiles.sh left in gpurun_out/ into the committed summaries under profiles/
set -e
# summaries made on the GPU box (gpurun_out/m_*.txt) win; else summarise the reports that came back
for r in icm8 icm16 scanx8 scanx8_fp32 unary_tc unary_exact; do
  if [ -f gpurun_out/m_$r.txt ]; then cp gpurun_out/m_$r.txt /tmp/m_$r.txt; else python tools/ncu_metrics.py gpurun_out/r2_$r.ncu-rep > /tmp/m_$r.txt; fi
done
H=$(git rev-parse --short HEAD)
{ echo "# round 2, HEAD ($H) -- ncu --set full --clock-control none, ONE launch of K3 at the bench workload"; echo "# command: tools/icm_bench.py 1000000 8 32 (1M x 128, m=8, ilsiter=32, icmiter=4, npert=4; device arrays, one chunk); kernel name = the shipped <8,1,1> (PF + JSPEC)"; cat /tmp/m_icm8.txt; echo "# instruction mix of the same capture (tools/sass_hot.py over the source page):"; if [ -f gpurun_out/icm8_source.csv ]; then cp gpurun_out/icm8_source.csv /tmp/icm8b.csv; else ncu -i gpurun_out/r2_icm8.ncu-rep --page source --csv 2>/dev/null > /tmp/icm8b.csv; fi; python tools/sass_hot.py /tmp/icm8b.csv samples 12 | head -40; } > profiles/r2_icm_warp_kernel_m8_full_workload.txt
{ echo "# round 2, HEAD ($H) -- ncu --set full, K3 at m = 16 (tools/icm_bench.py 125000 16 32: one GPU's shard of configs[3] on 8 GPUs)"; cat /tmp/m_icm16.txt; } > profiles/r2_icm_warp_kernel_m16.txt
{ echo "# round 2, HEAD ($H) -- ncu --set full, scan kernel main launch (tools/scan_bench.py 1000000 10000 8 <k>: 592 query tiles x 1M codes)"; echo "## k = 1: scanx_kernel<8,1,0,1> = the quantised pre-filter loop (two queries per fp32 word, lane-rotated codebook order)"; cat /tmp/m_scanx8.txt; echo "## k = 1, RAYUELA_B200_SCAN_PREFILTER=0: scanx_kernel<8,1,0,0> = the fp32 loop (also what k > 128 runs)"; cat /tmp/m_scanx8_fp32.txt; } > profiles/r2_scanx8_kernel.txt
{ echo "# round 2, HEAD ($H) -- ncu --set full, the opt-in tcgen05 unary GEMM and the exact fp32 kernel, 1M x 2048 x 128"; cat /tmp/m_unary_tc.txt; cat /tmp/m_unary_exact.txt; echo "# SASS of the shipped library (cuobjdump -sass | grep): tcgen05 / TMEM / bulk-copy instructions"; cuobjdump -sass rayuela.jl_b200/lib/librayuela_b200.so | grep -oE "UTCHMMA[.A-Z0-9_]*|LDTM[.a-zA-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*" | sort | uniq -c; } > profiles/r2_unary_tc_kernel.txt
{ echo "# round 2, HEAD ($H) -- ncu --metrics gpu__time_duration.sum --clock-control none"; echo "# command: bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong (1M x 128, m=8, ilsiter=32; 10k queries k=1; incl. the fast-mode legs); times are serialised/cold: compare SHARES"; python tools/launch_summary.py gpurun_out/r2_launches.csv | grep -v "native::\|cublasLt\|cuda::kernelHist\|larfb\|orgqr\|lacpy\|batch_eye\|copy_info\|randperm"; } > profiles/r2_launches.txt
cp gpurun_out/r2_bench.json profiles/r2_bench.json; cp gpurun_out/r2_configs01.json profiles/r2_baseline_configs_0_1.json; cp gpurun_out/r2_rows.json profiles/r2_secondary_rows.json; cp gpurun_out/r2_pytest_gpu.txt profiles/r2_pytest_gpu.txt; cp gpurun_out/r2_scan_bench.txt profiles/r2_scan_bench.txt
python - <<'PY'
import json,re
def grab(path):
    t=open(path).read(); mul={'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}
    r=re.search(r"dram__bytes_read.sum\s+([\d.]+) (\w+)", t); w=re.search(r"dram__bytes_write.sum\s+([\d.]+) (\w+)", t)
    return float(r.group(1))*mul[r.group(2)]+float(w.group(1))*mul[w.group(2)]
def pct(path, name):
    m = re.search(name + r"\s+([\d.]+)", open(path).read()); return round(float(m.group(1)), 1) if m else None
out={"icm_m8_n1000000_ils32": {"bytes": grab('/tmp/m_icm8.txt'), "lts_throughput_pct": pct('/tmp/m_icm8.txt', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), "issue_active_pct": pct('/tmp/m_icm8.txt', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), "kernel": "icm_warp_kernel<8,1,1>", "capture": "profiles/r2_icm_warp_kernel_m8_full_workload.txt", "what": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full"},
     "scan_m8_n1000000_nq10000_k1": {"bytes": grab('/tmp/m_scanx8.txt'), "kernel": "scanx_kernel<8,1,0,1> (pre-filter loop; main launch, 592 of 625 query tiles)", "capture": "profiles/r2_scanx8_kernel.txt", "what": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full"}}
json.dump(out, open('profiles/ncu_traffic.json','w'), indent=1)
PY
echo written
