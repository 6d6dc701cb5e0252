"""Lower bound for the large-k selection cost: rerun the search with the true k-th distance as initial threshold."""
import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n, nq, d, m = 1000000, 10000, 128, 8
g = torch.Generator(device=dev).manual_seed(0)
B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
nrm = torch.randn(n, device=dev, generator=g) * 3
Q = torch.randn(nq, d, device=dev, generator=g)
C = torch.randn(m * 256, d, device=dev, generator=g)
ix = core.Index(core.SCAN_LSQ, B, nrm)
def t(k):
    for _ in range(2): ix.search(Q, C, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): r = ix.search(Q, C, k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 4, r
for k in (100, 1000, 4000):
    os.environ.pop("RAYUELA_B200_SCAN_TAU0", None)
    ms0, (d0, i0) = t(k)
    kth = d0[:, -1]
    for mult, name in ((1.0, "max over queries of the true k-th"),):
        os.environ["RAYUELA_B200_SCAN_TAU0"] = repr(float(kth.max()))
        ms1, (d1, i1) = t(k)
        same = bool((i0 == i1).all())
        print(f"k={k}: normal {ms0:.2f} ms; with tau0 = {name}: {ms1:.2f} ms (ids identical: {same}); k-th dist spread min {float(kth.min()):.2f} max {float(kth.max()):.2f}", flush=True)
