"""BASELINE config 5 on ONE GPU: 100M x 8-byte codes + fp32 norms, 10k queries; indices verified against the
reference's compiled C++ (oracle/_ref) on a 64-query subset."""
import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
from oracle import oracle as orc
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nq, m, d = 10000, 8, 128
g = torch.Generator(device=dev).manual_seed(0)
B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
nrm = (torch.randn(n, device=dev, generator=g) * 3).contiguous()
Q = torch.randn(nq, d, device=dev, generator=g)
C = torch.randn(m * 256, d, device=dev, generator=g)
t0 = time.perf_counter(); ix = core.Index(core.SCAN_LSQ, B, nrm); torch.cuda.synchronize(); t_ix = time.perf_counter() - t0
for k in (1, 100):
    ix.search(Q[:64], C, k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dd, ii = ix.search(Q, C, k); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n={n} nq={nq} k={k}: {ms:.1f} ms  {nq/ms*1e3:,.0f} q/s  alg {nq*n*12/ms/1e6:,.0f} GB/s (index build {t_ix*1e3:.0f} ms)", flush=True)
Bh, nh = B.cpu().numpy(), nrm.cpu().numpy()
t0 = time.perf_counter()
d0, i0 = orc.ref_linscan(orc.LSQ, Bh, Q[:64].cpu().numpy(), C.cpu().numpy(), 100, nh)
t_cpu = time.perf_counter() - t0
ok = np.array_equal(ii[:64].cpu().numpy(), i0) and np.array_equal(dd[:64].cpu().numpy().view(np.uint32), d0.view(np.uint32))
print(f"reference C++ on 64 queries: {t_cpu:.1f} s ({64/t_cpu:.1f} q/s); GPU ids+dists identical: {ok}")
