import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n, nq, d = 1000000, 10000, 128
g = torch.Generator(device=dev).manual_seed(0)
for m in (8, 16):
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device=dev, generator=g) * 3
    Q = torch.randn(nq, d, device=dev, generator=g)
    C = torch.randn(m * 256, d, device=dev, generator=g)
    ix = core.Index(core.SCAN_LSQ, B, nrm)
    for k in (1, 100, 1000):
        for _ in range(3): ix.search(Q, C, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ix.search(Q, C, k)
        e1.record(); torch.cuda.synchronize()
        print(f"m={m} k={k}: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
