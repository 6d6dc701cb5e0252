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
# ---- round 2: chunk pipeline (two streams + copy stream), dynamic schedule, any-h kernels, tcgen05 unaries, merge tree,
# multi-device (two shards on device 0), ragged query tile, get_unaries
import os
import rayuela_b200 as rb
os.environ["RAYUELA_B200_UNARY_BYTES"] = str(8 * 1024 * 1024)             # 1024 vectors per chunk -> 3 chunks
core.encode_icm(X, C, B, 2, 2, 4, True, seed=1, want_cost=True, snap_iters=[2])
del os.environ["RAYUELA_B200_UNARY_BYTES"]
for h in (16, 100):
    Ch = r.standard_normal((4 * h, d)).astype(np.float32); Bh = r.integers(0, h, (700, 4), dtype=np.uint8)
    core.encode_icm(X[:700], Ch, Bh, 2, 2, 3, True, seed=2, h=h, want_cost=True, snap_iters=[1])
    core.veccost(X[:700], Bh, Ch, h=h)
core.get_unaries(X[:300], C, m); core.get_unaries(X[:300], C, m, fast=True)
core.encode_icm(X[:1000], C, B[:1000], 1, 2, 4, True, seed=1, fast=True)
dd = np.sort(r.standard_normal((3, 4, 6000)).astype(np.float32), axis=2); ii = r.permutation(3 * 4 * 6000).reshape(3, 4, 6000).astype(np.int32)
core.topk_merge(dd, ii)                                                      # 18000 keys per query: the rank-merge tree
rb.init([0, 0])
core.encode_icm(X, C, B, 1, 2, 4, True, seed=1, snap_iters=[1])
ix = core.Index(core.SCAN_LSQ, out['B'], nrm); ix.search(Q, C, 10); ix.free()
rb.init(None)
core.Index(core.SCAN_PQ, Bp).search(Q[:17], Cpq, 3)                          # ragged last query tile (dummy queries)
print("sanitize run done")
