import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X, Q = bench.make_data(n, 100, 128, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev)
for fast in (False, True):
    for _ in range(2): U = core.get_unaries(X, C, m, fast=fast)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): U = core.get_unaries(X, C, m, fast=fast)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"get_unaries fast={fast} n={n} m={m}: {ms:.3f} ms  ({n*m*256*4/ms/1e6:.0f} GB/s written, {2*n*m*256*128/ms/1e9:.1f} TFLOP/s fp32-equivalent)", flush=True)
    if fast: print("max abs diff vs exact:", float((U - Uex).abs().max()), "max |U|", float(Uex.abs().max()))
    else: Uex = U.clone()
    del U
B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
for fast in (False, True):
    for _ in range(2): core.encode_icm(X, C, B0, 32, 4, 4, True, seed=2024, fast=fast)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): r = core.encode_icm(X, C, B0, 32, 4, 4, True, seed=2024, fast=fast)
    e1.record(); torch.cuda.synchronize()
    print(f"encode_icm fast={fast}: {e0.elapsed_time(e1)/3:.2f} ms  qerr={core.qerror(X, r['B'], C):.6f}", flush=True)
