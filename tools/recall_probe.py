import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
n, nq, d, m = 200000, 2000, 128, 8
def gen(alpha, ncl, noise, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    s = (torch.arange(1, d + 1, device=dev).float()) ** (-alpha)
    Rm, _ = torch.linalg.qr(torch.randn(d, d, generator=g, device=dev))
    cen = torch.randn(ncl, d, generator=g, device=dev) * s
    def draw(k):
        z = cen[torch.randint(0, ncl, (k,), generator=g, device=dev)] + noise * torch.randn(k, d, generator=g, device=dev) * s
        return (z @ Rm).contiguous()
    return draw(n), draw(nq)
for alpha, ncl, noise in [(0.0,1024,0.3),(0.5,1024,0.5),(0.75,1024,1.0),(1.0,1024,1.0),(0.75,1,1.0),(1.0,1,1.0),(1.25,1,1.0)]:
    X, Q = gen(alpha, ncl, noise, 1)
    C = bench.train_codebooks(X[:50000], m, dev)
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
    core.encode_icm(X, C, B, 8, 4, 4, True, seed=1, inplace=True)
    q0 = core.qerror(X, B, C); var = float((X*X).sum(1).mean())
    Cm = C.reshape(m, 256, d); rec = sum(Cm[j][B[:, j].long()] for j in range(m))
    ix = core.Index(core.SCAN_LSQ, B, (rec*rec).sum(1).contiguous())
    dd, ii = ix.search(Q, C, 100)
    gt = bench.exact_nn(X, Q)
    hit = (ii.long() - 1 == gt[:, None])
    print(f"alpha={alpha} ncl={ncl} noise={noise}: rel qerr={q0/var:.4f} R@1={hit[:, :1].any(1).float().mean():.3f} R@10={hit[:, :10].any(1).float().mean():.3f} R@100={hit.any(1).float().mean():.3f}", flush=True)
