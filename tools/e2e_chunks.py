"""Host-array (pinned) encode, end to end, for different chunk counts of the upload / compute pipeline."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n, m = 1000000, 8
X, Q = bench.make_data(n, 100, 128, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev)
Xh = X.cpu().pin_memory(); Ch = C.cpu().pin_memory()
B0 = torch.randint(0, 256, (n, m), dtype=torch.uint8).pin_memory()
for ch in sys.argv[1:] or ["4", "6", "8", "12", "16"]:
    os.environ["RAYUELA_B200_ICM_CHUNKS"] = ch
    ts = []
    for rep in range(4):
        B = B0.clone().pin_memory()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        core.encode_icm(Xh.numpy(), Ch.numpy(), B.numpy(), 32, 4, 4, True, seed=2024, inplace=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("chunks", ch, "e2e ms", ["%.1f" % t for t in ts], flush=True)
