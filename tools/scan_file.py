"""Scan timing on a dumped case (BENCH_DUMP_SCAN=... python bench.py --path linscan): python tools/scan_file.py case.npz k1,k2 [reps]"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
z = np.load(sys.argv[1])
ks = [int(x) for x in sys.argv[2].split(',')]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device('cuda')
B, nrm, Q, C = (torch.from_numpy(z[k]).to(dev) for k in ('B', 'nrm', 'Q', 'C'))
ix = core.Index(core.SCAN_LSQ, B, nrm)
for k in ks:
    for _ in range(3): ix.search(Q, C, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = ix.search(Q, C, k)
    e1.record(); torch.cuda.synchronize()
    print(f"k={k}: {e0.elapsed_time(e1) / reps:.3f} ms  ids checksum {int(r[1].long().sum())}", flush=True)
