import sys, torch, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
ms_ = [int(x) for x in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['8'])]
ks = [int(x) for x in (sys.argv[4].split(',') if len(sys.argv) > 4 else ['1', '100', '1000'])]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
g = torch.Generator(device=dev).manual_seed(0)
for m in ms_:
    d = 128
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device=dev, generator=g) * 3
    Q = torch.randn(nq, d, device=dev, generator=g)
    C = torch.randn(m * 256, d, device=dev, generator=g)
    for kind, name in ((core.SCAN_LSQ, 'lsq'), (core.SCAN_CQ, 'cq')):
        ix = core.Index(kind, B, nrm if kind == core.SCAN_LSQ else None)
        for k in ks:
            for _ in range(4): ix.search(Q, C, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): ix.search(Q, C, k)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            byts = nq * n * (m + (4 if kind == core.SCAN_LSQ else 0))
            print(f"{name} m={m} n={n} nq={nq} k={k}: {ms:.3f} ms  {nq/ms*1e3:,.0f} q/s  alg {byts/ms/1e6:,.0f} GB/s  lookups/clk/SM {nq*n*m/(ms*1e-3)/148/1.965e9:.2f}", flush=True)
        ix.free()
