import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ils = int(sys.argv[3]) if len(sys.argv) > 3 else 32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
X, Q = bench.make_data(n, 100, 128, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev)
B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
B = B0.clone()
def step(i):
    B.copy_(B0); return core.encode_icm(X, C, B, i, 4, 4, True, seed=2024, inplace=True, want_stats=(i>0))
for _ in range(2): step(ils)
for i in (0, ils):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"m={m} n={n} ilsiter={i}: {ms:.2f} ms/step  {n/ms*1e3:,.0f} vectors/s  qerr={core.qerror(X, B, C):.4f}", flush=True)
print("steps executed/total:", core.last_icm_steps(), "exact:", core.last_icm_exact_steps())
