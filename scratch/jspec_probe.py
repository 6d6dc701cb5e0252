import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n = 1000000
X, Q = bench.make_data(n, 100, 128, 1000, dev)
for m in (8, 7):
    C = bench.train_codebooks(X[:50000], m, dev)
    B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
    outs = {}
    for v in ("0", "1", "0", "1"):
        os.environ["RAYUELA_B200_ICM_JSPEC"] = v
        B = B0.clone()
        def step():
            B.copy_(B0); core.encode_icm(X, C, B, 32, 4, 4, True, seed=2024, inplace=True)
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): step()
        e1.record(); torch.cuda.synchronize()
        outs[v] = B.clone()
        print(f"m={m} JSPEC={v}: {e0.elapsed_time(e1)/3:.2f} ms", flush=True)
    print("  same codes:", bool(torch.equal(outs["0"], outs["1"])), flush=True)
