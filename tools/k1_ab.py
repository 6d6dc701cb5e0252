"""K1 A/B: exact unary kernel v2 (FFMA2, two shared-memory stages) vs the round-1 kernel, bitwise comparison + timing."""
import os, sys, torch, subprocess
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = 8
X, Q = bench.make_data(n, 100, 128, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev)
outs = {}
for v2 in ("1", "0"):
    os.environ["RAYUELA_B200_K1_V2"] = v2
    for _ in range(2): U = core.get_unaries(X, C, m, fast=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): U = core.get_unaries(X, C, m, fast=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"K1 v2={v2}: {ms:.3f} ms  ({2*n*m*256*128/ms/1e9:.1f} TFLOP/s)", flush=True)
    outs[v2] = U[: 200000].clone()
    del U
print("bit-identical:", bool(torch.equal(outs["1"].view(torch.int32), outs["0"].view(torch.int32))))
