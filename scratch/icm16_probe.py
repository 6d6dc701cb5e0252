import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 125000 if m > 8 else 1000000
X, Q = bench.make_data(max(n, 50000), 100, 128, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev); X = X[:n].contiguous()
B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
ref = None
for bps in ("4", "3", "2"):
    os.environ["RAYUELA_B200_ICM_BLOCKS_PER_SM"] = bps
    B = B0.clone()
    def step():
        B.copy_(B0); core.encode_icm(X, C, B, 32, 4, 4, True, seed=2024, inplace=True)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ref = B.clone() if ref is None else ref
    print(f"m={m} n={n} blocks/SM={bps}: {ms:.2f} ms  {n/ms*1e3:,.0f} vectors/s  same codes: {bool(torch.equal(ref, B))}", flush=True)
