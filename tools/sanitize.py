"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
r = np.random.default_rng(0)
n, d, m = 3000, 32, 8
X = r.standard_normal((n, d)).astype(np.float32); C = (r.standard_normal((m*256, d))/3).astype(np.float32)
B = r.integers(0, 256, (n, m), dtype=np.uint8)
out = core.encode_icm(X, C, B, 2, 2, 4, True, seed=1, want_cost=True, want_stats=True, snap_iters=[1, 2])
core.veccost(X, out['B'], C, want_mean=True)
_, nrm = core.quantize_norms(out['B'], C); core.quantize_norms(out['B'], C, np.sort(r.random(256)).astype(np.float32))
Q = r.standard_normal((37, d)).astype(np.float32)
for kind in (core.SCAN_LSQ, core.SCAN_CQ):
    for k in (1, 700, 3000):
        core.Index(kind, out['B'], nrm).search(Q, C, k)
B16 = r.integers(0, 256, (n, 16), dtype=np.uint8); C16 = r.standard_normal((16*256, d)).astype(np.float32)
for k in (1, 50, 1200):
    core.Index(core.SCAN_CQ, B16).search(Q, C16, k)
B16b = r.integers(0, 256, (6000, 16), dtype=np.uint8)
core.Index(core.SCAN_LSQ, B16b, r.standard_normal(6000).astype(np.float32)).search(Q, C16, 5000)   # k > one pass: lower-bounded passes
# a larger base so that several compactions, piggy-backed ones and the speculative threshold all happen
nb = 24000
Bb = r.integers(0, 256, (nb, m), dtype=np.uint8); nb_nrm = r.standard_normal(nb).astype(np.float32)
for k in (1, 100, 1000):
    core.Index(core.SCAN_LSQ, Bb, nb_nrm).search(Q, C, k)
out16 = core.encode_icm(X[:800], C16, B16[:800], 1, 2, 4, True, seed=3)
Cpq = r.standard_normal((m*256, d//m)).astype(np.float32)
Bp = core.quantize_pq(X, Cpq, m); core.Index(core.SCAN_PQ, Bp).search(Q, Cpq, 20)
core.fast_bin_matmul(X, B); core.quantize_chainq(X[:500], C, m)
dd = np.sort(r.standard_normal((3, 5, 16)).astype(np.float32), axis=2); ii = r.permutation(240).reshape(3, 5, 16).astype(np.int32)
core.topk_merge(dd, ii)
print("sanitize run done")
