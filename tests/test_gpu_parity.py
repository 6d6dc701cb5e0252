"""GPU parity: the CUDA path (through the C ABI) against the oracle, the reference's own compiled C++
(oracle/_ref, when it travelled with the snapshot) and the committed golden vectors.
Bar: codes / ids bit-exact; fp32 distances and costs bit-exact here (the north_star tolerance is 1e-4
relative, stated where a mean is compared)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rayuela_b200
    return rayuela_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _icm_data(n, d, m, seed, kind="gauss"):
    r = np.random.default_rng(seed)
    if kind == "uniform":
        X = (r.random((n, d)) * 10).astype(np.float32)
        C = r.random((m * 256, d)).astype(np.float32)
    else:
        X = r.standard_normal((n, d)).astype(np.float32)
        C = (r.standard_normal((m * 256, d)) / np.sqrt(m)).astype(np.float32)
    return X, C, r.integers(0, 256, (n, m), dtype=np.uint8)


# ---- path (1) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["icm_m8_gauss", "icm_m7_uniform", "icm_m16_gauss"])
def test_icm_golden(rb, name):
    g = gold(name)
    r = rb.core.encode_icm(g["X"], g["C"], g["B"], int(g["ilsiter"]), int(g["icmiter"]), int(g["npert"]),
                           bool(g["randord"]), seed=int(g["seed"]), g0=int(g["g0"]), snap_iters=g["snap_iters"],
                           want_cost=True, want_stats=True)
    assert np.array_equal(r["B"], g["B_out"])
    assert np.array_equal(bits(r["cost"]), bits(g["cost"]))
    assert np.array_equal(r["stats"], g["stats"])
    assert np.array_equal(r["B_snap"], g["B_snap"])
    assert np.allclose(r["objs"], g["objs"], rtol=1e-4)      # qerror: 1e-4 relative (north_star)


@pytest.mark.parametrize("n,d,m,ils,icm,npert,randord,kind", [
    (5000, 128, 8, 4, 4, 4, True, "gauss"),        # BASELINE config-3 shape, reduced n
    (3001, 128, 7, 2, 4, 4, True, "gauss"),        # demo shape: 7 codebooks + norm byte; ragged n
    (1500, 96, 16, 2, 2, 4, False, "uniform"),     # config-4 codebook count, test/common.jl data
    (777, 30, 3, 3, 3, 9, True, "gauss"),          # d % 4 != 0, npert > 4 (three Philox blocks)
    (1, 8, 2, 1, 1, 1, False, "gauss"),            # single vector
    (64, 16, 1, 2, 2, 2, True, "gauss"),           # m = 1: no pairwise terms at all
])
def test_icm_matches_oracle(rb, n, d, m, ils, icm, npert, randord, kind):
    X, C, B = _icm_data(n, d, m, seed=n + m, kind=kind)
    o = orc.encode_icm(X, C, B, ils, icm, npert, randord, seed=42, g0=123456789012)
    r = rb.core.encode_icm(X, C, B, ils, icm, npert, randord, seed=42, g0=123456789012, want_cost=True,
                           want_stats=True)
    assert np.array_equal(r["B"], o["B"])
    assert np.array_equal(bits(r["cost"]), bits(o["cost"]))
    assert np.array_equal(r["stats"], o["stats"])


def test_icm_zero_iterations_and_explicit_orders(rb):
    X, C, B = _icm_data(500, 32, 8, seed=3)
    r = rb.core.encode_icm(X, C, B, 0, 4, 4, True, want_cost=True)
    assert np.array_equal(r["B"], B)
    assert np.array_equal(bits(r["cost"]), bits(orc.veccost(X, B, C)))
    orders = np.array([[7, 6, 5, 4, 3, 2, 1, 0], [0, 2, 4, 6, 1, 3, 5, 7]], dtype=np.int32)
    o = orc.encode_icm(X, C, B, 2, 3, 4, True, seed=1, orders=orders)
    r = rb.core.encode_icm(X, C, B, 2, 3, 4, True, seed=1, orders=orders)
    assert np.array_equal(r["B"], o["B"])


def test_icm_shard_invariance_and_chunking(rb, monkeypatch):
    """Codes depend on the GLOBAL vector index only: encoding halves with g0 offsets == encoding the whole;
    and a tiny unary budget (forces several chunks inside one call) changes nothing."""
    X, C, B = _icm_data(4000, 64, 8, seed=9)
    whole = rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=5)["B"]
    a = rb.core.encode_icm(X[:1700], C, B[:1700], 3, 4, 4, True, seed=5, g0=0)["B"]
    b = rb.core.encode_icm(X[1700:], C, B[1700:], 3, 4, 4, True, seed=5, g0=1700)["B"]
    assert np.array_equal(np.concatenate([a, b]), whole)
    monkeypatch.setenv("RAYUELA_B200_UNARY_BYTES", str(1100 * 8 * 1024))
    assert np.array_equal(rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=5)["B"], whole)


def test_icm_device_pointers(rb):
    import torch
    X, C, B = _icm_data(2000, 128, 8, seed=21)
    o = orc.encode_icm(X, C, B, 2, 4, 4, True, seed=8)
    Xd, Cd, Bd = (torch.from_numpy(a).cuda() for a in (X, C, B))
    r = rb.core.encode_icm(Xd, Cd, Bd, 2, 4, 4, True, seed=8, want_cost=True)
    assert np.array_equal(r["B"].cpu().numpy(), o["B"])
    assert np.array_equal(bits(r["cost"].cpu().numpy()), bits(o["cost"]))
    assert np.array_equal(Bd.cpu().numpy(), B)            # not in place unless asked


def test_icm_improves_and_never_worsens(rb):
    X, C, B = _icm_data(3000, 128, 8, seed=33)
    c0 = rb.core.veccost(X, B, C)
    r = rb.core.encode_icm(X, C, B, 4, 4, 4, True, seed=2, want_cost=True)
    assert np.all(r["cost"] <= c0)
    assert r["cost"].mean() < 0.7 * c0.mean()


@pytest.mark.parametrize("m,d", [(8, 128), (7, 33), (16, 64)])
def test_veccost_qerror(rb, m, d):
    X, C, B = _icm_data(2500, d, m, seed=m)
    cost, mean = rb.core.veccost(X, B, C, want_mean=True)
    assert np.array_equal(bits(cost), bits(orc.veccost(X, B, C)))
    assert abs(mean - orc.qerror(X, B, C)) <= 1e-6 * abs(mean)


@pytest.mark.parametrize("m", [8, 5])
def test_condition_compat_symbol(rb, m):
    """The exact-signature `condition` against the oracle's step (itself pinned to the reference's)."""
    n, d = 2000, 32
    X, C, B = _icm_data(n, d, m, seed=50 + m)
    U = orc.get_unaries(X, C, m)
    bins, bins_t, cbi = orc.get_binaries(C, m)
    p2i = np.zeros((m, m), dtype=np.int32)
    for i, (a, b) in enumerate(cbi):
        p2i[a, b] = p2i[b, a] = i
    B1, B2 = B.copy(), B.copy()
    for j in (2, 0, m - 1):
        tc = np.array([k for k in range(m) if k != j], dtype=np.int32)
        u1, u2 = U[j].copy(), U[j].copy()
        orc.condition(B1, u1, bins, bins_t, p2i, tc, j, use_ref=orc.have_ref())
        rb.core.c_condition(B2, u2, bins, bins_t, p2i, tc, j)
        assert np.array_equal(B1, B2)
        assert np.array_equal(bits(u1), bits(u2))


# ---- path (2) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["scan_lsq_m8", "scan_lsq_m7_ties", "scan_cq_m8", "scan_pq_m8", "scan_pq_m16_ties"])
def test_scan_golden(rb, name):
    g = gold(name)
    kind = int(g["kind"])
    ix = rb.core.Index(kind, g["B"], g.get("nrm"))
    d, i = ix.search(g["Xq"], g["cb"], int(g["k"]))
    assert np.array_equal(bits(d), bits(g["dists"]))
    assert np.array_equal(i, g["idx"])


def _scan_case(kind, n, nq, m, d, seed, ties=False):
    r = np.random.default_rng(seed)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d // m if kind == orc.PQ else d)).astype(np.float32)
    if ties:
        Xq, cb = np.round(Xq * 2), np.round(cb * 2)
        B = r.integers(0, 4, (n, m), dtype=np.uint8)
    nrm = None
    if kind == orc.LSQ:
        nrm = (np.round(r.standard_normal(n) * 3) if ties else r.standard_normal(n) * 3).astype(np.float32)
    return B, Xq, cb, nrm


@pytest.mark.parametrize("kind", [orc.LSQ, orc.CQ, orc.PQ])
@pytest.mark.parametrize("n,nq,m,k,ties", [
    (100000, 40, 8, 1, False),
    (100000, 33, 8, 1000, False),      # demos' knn (demos_train_query_base.jl:16); several compactions
    (50000, 7, 7, 100, True),          # 7 codebooks (+ norm byte in the demo); massive ties
    (30000, 16, 16, 10, False),
    (300000, 5, 8, 50, True),          # few queries -> several DB slices + merge
    (1000, 3, 4, 1000, False),         # k == n
    (1500, 1, 1, 7, False),            # m = 1, single query
    (100000, 33, 16, 1000, False),     # 16-codebook period of the skewed scan, several compactions
    (50000, 7, 15, 100, True),         # 15 codebooks (+ norm byte, demos' m = 16); massive ties
    (300000, 5, 12, 50, True),         # m = 12, few queries -> DB slices + merge
    (2000, 3, 9, 2000, False),         # m = 9, k == n
])
def test_scan_matches_reference(rb, kind, n, nq, m, k, ties):
    d = 16 * m if kind == orc.PQ else 64
    B, Xq, cb, nrm = _scan_case(kind, n, nq, m, d, seed=n + k + kind, ties=ties)
    if orc.have_ref():
        d0, i0 = orc.ref_linscan(kind, B, Xq, cb, k, nrm)      # the reference's own compiled C++
    else:
        d0, i0 = orc.linscan(kind, B, Xq, cb, k, nrm)
    ix = rb.core.Index(kind, B, nrm)
    d1, i1 = ix.search(Xq, cb, k)
    assert np.array_equal(i1, i0)
    assert np.array_equal(bits(d1), bits(d0))


@pytest.mark.parametrize("kind,m,k", [(orc.LSQ, 8, 10000), (orc.PQ, 16, 5000), (orc.CQ, 7, 4097)])
def test_scan_large_k_multi_pass(rb, kind, m, k):
    """k beyond one pass (the reference's default is k = 10000, src/Linscan.jl:10): several passes, each bounded
    below by the last key of the previous one; tie-heavy data so pass boundaries fall inside runs of equal
    distances."""
    n, nq = 30000, 6
    d = 16 * m if kind == orc.PQ else 32
    B, Xq, cb, nrm = _scan_case(kind, n, nq, m, d, seed=k, ties=True)
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(kind, B, Xq, cb, k, nrm)
    d1, i1 = rb.core.Index(kind, B, nrm).search(Xq, cb, k)
    assert np.array_equal(i1, i0) and np.array_equal(bits(d1), bits(d0))


@pytest.mark.parametrize("m,k", [(8, 300), (5, 1), (16, 300), (11, 2500)])
def test_scan_compaction_schedule_does_not_change_results(rb, monkeypatch, m, k):
    """The soft limit only decides WHEN a candidate buffer is compacted (event-driven, any warp may raise the
    flag); the result must be the reference's bits for every schedule, on tie-heavy data."""
    B, Xq, cb, nrm = _scan_case(orc.LSQ, 70000, 19, m, 64, seed=m, ties=True)
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(orc.LSQ, B, Xq, cb, k, nrm)
    for soft in ("1", "700", "100000"):
        monkeypatch.setenv("RAYUELA_B200_SCAN_SOFT", soft)
        d1, i1 = rb.core.Index(orc.LSQ, B, nrm).search(Xq, cb, k)
        assert np.array_equal(i1, i0) and np.array_equal(bits(d1), bits(d0)), soft


@pytest.mark.parametrize("m,k", [(8, 1000), (8, 64), (16, 300)])
def test_scan_speculative_threshold_failure_is_redone_exactly(rb, monkeypatch, m, k):
    """The speculative filter threshold assumes nothing it does not verify: on a base ORDERED by distance to the
    query (best codes first -- the sample the threshold is estimated from is as unrepresentative as it gets) the
    verification must fail and the redo launch must return the reference's bits; same with the knob off."""
    n, nq = 120000, 5
    B, Xq, cb, nrm = _scan_case(orc.LSQ, n, nq, m, 64, seed=100 + m)
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    _, order = fn(orc.LSQ, B, Xq[:1], cb, n, nrm)              # full ranking for query 0 (ids are 1-based)
    perm = order[0].astype(np.int64) - 1
    B, nrm = np.ascontiguousarray(B[perm]), np.ascontiguousarray(nrm[perm])
    d0, i0 = fn(orc.LSQ, B, Xq, cb, k, nrm)
    assert np.array_equal(i0[0], np.arange(1, k + 1))          # query 0's neighbours are now the first k codes
    for val in ("1", "0"):
        monkeypatch.setenv("RAYUELA_B200_SCAN_SPEC", val)
        d1, i1 = rb.core.Index(orc.LSQ, B, nrm).search(Xq, cb, k)
        assert np.array_equal(i1, i0) and np.array_equal(bits(d1), bits(d0)), val
    # and the mirror image: best codes last
    B, nrm = np.ascontiguousarray(B[::-1]), np.ascontiguousarray(nrm[::-1])
    d0, i0 = fn(orc.LSQ, B, Xq, cb, k, nrm)
    monkeypatch.setenv("RAYUELA_B200_SCAN_SPEC", "1")
    d1, i1 = rb.core.Index(orc.LSQ, B, nrm).search(Xq, cb, k)
    assert np.array_equal(i1, i0) and np.array_equal(bits(d1), bits(d0))


@pytest.mark.parametrize("case", ["outlier_norms", "offset_tables", "flat_tables", "outlier_query", "near_ties",
                                  "negative_zero", "tiny_h"])
@pytest.mark.parametrize("m,k", [(8, 1), (8, 200), (16, 50), (5, 1000)])
def test_scan_prefilter_degenerate_inputs(rb, monkeypatch, case, m, k):
    """The scan's quantised pre-filter (two queries per fp32 word, exact re-evaluation of its survivors) must return
    the reference's bits whatever the quantisation does to the data: one huge norm stretching the tile's scale, tables
    with a large common offset (fp32 rounding slack of the exact chain dominates the window), all-equal tables (scale
    0 -> 1), one query with a 1e4 x larger range sharing the tile scale with the others, distances that differ in the
    last bits only, and the -0.0 / h < 256 corners.  Same bits with the pre-filter off."""
    r = np.random.default_rng(1000 + m + k)
    n, nq, d, h = 60000, 21, 32, 256
    kind = orc.LSQ
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d)).astype(np.float32)
    nrm = (r.standard_normal(n) * 3).astype(np.float32)
    if case == "outlier_norms":
        nrm[r.integers(0, n, 3)] = np.float32(1e9)
        nrm[r.integers(0, n, 3)] = np.float32(-1e9)
    elif case == "offset_tables":
        kind = orc.CQ                                  # LUT = ||q - c||^2 with a huge common part
        cb += np.float32(300.0)
        nrm = None
    elif case == "flat_tables":
        cb[:] = np.float32(0.25)
        nrm[:] = np.float32(1.5)                       # every distance identical: pure id order
    elif case == "outlier_query":
        Xq[3] *= np.float32(1e4)
    elif case == "near_ties":
        cb = np.round(cb * 4) / np.float32(4)          # few distinct table values, many exact and near ties
        Xq = np.round(Xq * 2) / np.float32(2)
        B = r.integers(0, 6, (n, m), dtype=np.uint8)
        nrm = (np.round(r.standard_normal(n) * 8) / 8).astype(np.float32)
    elif case == "negative_zero":
        Xq[5] = 0.0                                    # LUT entries -0.0 / +0.0
        nrm[::7] = np.float32(-0.0)
    elif case == "tiny_h":
        h = 16
        B = r.integers(0, h, (n, m), dtype=np.uint8)
        cb = r.standard_normal((m * h, d)).astype(np.float32)
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    d0, i0 = fn(kind, B, Xq, cb, k, nrm) if h == 256 else fn(kind, B, Xq, cb, k, nrm, h=h)
    for val in ("1", "0"):
        monkeypatch.setenv("RAYUELA_B200_SCAN_PREFILTER", val)
        ix = rb.core.Index(kind, B, nrm, h=h) if h != 256 else rb.core.Index(kind, B, nrm)
        d1, i1 = ix.search(Xq, cb, k)
        assert np.array_equal(i1, i0), (case, val)
        assert np.array_equal(bits(d1), bits(d0)), (case, val)


@pytest.mark.parametrize("with_inf", [True, False])
def test_scan_prefilter_non_finite_norms(rb, with_inf):
    """An infinite norm cannot be quantised: the index then scans with the plain fp32 loop.  NaN norms are skipped when
    the norm range is taken, so the pre-filter stays on; either way a NaN distance is never returned and +inf ones
    sort last, as ever."""
    r = np.random.default_rng(77)
    n, nq, d, m, k = 20000, 9, 32, 8, 20
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d)).astype(np.float32)
    nrm = (r.standard_normal(n) * 3).astype(np.float32)
    nrm[100] = np.inf if with_inf else np.nan
    nrm[200] = np.nan
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    good = np.ones(n, bool)
    good[[100, 200]] = False
    d0, i0 = fn(orc.LSQ, B[good], Xq, cb, k, nrm[good])
    ids = np.flatnonzero(good) + 1
    d1, i1 = rb.core.Index(orc.LSQ, B, nrm).search(Xq, cb, k)
    assert np.array_equal(i1, ids[i0 - 1]) and np.array_equal(bits(d1), bits(d0))


def test_icm_prefilter_on_and_off_agree(rb, monkeypatch):
    r = np.random.default_rng(3)
    n, d, m = 5000, 48, 8
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 3).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    want = orc.encode_icm(X, C, B, 3, 4, 4, True, seed=9)
    for val in ("0", "1"):
        monkeypatch.setenv("RAYUELA_B200_ICM_PF", val)
        got = rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=9, want_cost=True, want_stats=True)
        assert np.array_equal(got["B"], want["B"]), val
        assert np.array_equal(bits(got["cost"]), bits(want["cost"])), val
        ex, tot = rb.core.last_icm_steps()
        exact = rb.core.last_icm_exact_steps()
        assert 0 < ex <= tot and (exact == ex if val == "0" else exact < ex // 4), (val, ex, tot, exact)


@pytest.mark.parametrize("case", ["ties", "zero_codebook", "huge_unaries", "nan_input"])
def test_icm_prefilter_degenerate_inputs(rb, case):
    """Inputs on which the quantised pre-filter cannot decide anything must fall back to the exact rows and still
    give the oracle's codes: exact ties everywhere, an all-zero codebook (tmax = 0 for its partner tables), unaries
    far above the table range (integer sums would overflow), NaN in the data."""
    r = np.random.default_rng(11)
    n, d, m = 3000, 16, 8
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 3).astype(np.float32)
    if case == "ties":
        X, C = np.round(X), np.round(C * 2)                      # small integers: masses of exactly equal sums
        C[256:512] = C[:256]                                     # two identical codebooks
    elif case == "zero_codebook":
        C[3 * 256:4 * 256] = 0
    elif case == "huge_unaries":
        X *= 1e9
        C[:256] *= 1e-3
    elif case == "nan_input":
        X[7, 3] = np.nan
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    want = orc.encode_icm(X, C, B, 2, 3, 4, True, seed=5)
    got = rb.core.encode_icm(X, C, B, 2, 3, 4, True, seed=5, want_cost=True)
    ok = ~np.isnan(want["cost"])
    assert np.array_equal(got["B"][ok], want["B"][ok])
    assert np.array_equal(bits(got["cost"][ok]), bits(want["cost"][ok]))


@pytest.mark.parametrize("m", [4, 5, 6, 7, 8, 9, 12, 13, 16])
@pytest.mark.parametrize("kind", ["ties", "wide_unaries", "gauss"])
def test_icm_window_and_saturation_paths(rb, m, kind):
    """The three ways a step of the pre-filter kernels can end, for the m <= 8 (per-j row loops) and the m > 8 (uniform
    loop) kernels: a unique window member (gauss), a near-tie decided on the window's candidates by their exact
    chains -- lowest index on exactly equal sums (ties: small integers), more than 4 / 8 of them falling back to the
    whole rows -- and codebooks whose unaries do not fit the 16-bit fields of the shared-memory copy (wide_unaries:
    flagged per vector and codebook, every such step takes the exact rows)."""
    r = np.random.default_rng(100 + m)
    n, d = 2500, 32
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 3).astype(np.float32)
    if kind == "ties":
        X, C = np.round(X * 2), np.round(C * 3)
    elif kind == "wide_unaries":
        X *= 40.0
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    want = orc.encode_icm(X, C, B, 2, 3, 4, True, seed=21)
    got = rb.core.encode_icm(X, C, B, 2, 3, 4, True, seed=21, want_cost=True, want_stats=True)
    assert np.array_equal(got["B"], want["B"])
    assert np.array_equal(bits(got["cost"]), bits(want["cost"]))
    ex, _ = rb.core.last_icm_steps()
    exact = rb.core.last_icm_exact_steps()
    if kind == "gauss":
        assert exact < ex // 20, (ex, exact)             # whole-row fallbacks are rare on ordinary data
    if kind == "wide_unaries":
        assert exact > ex // 4, (ex, exact)              # the saturation flag sends these steps to the exact rows


@pytest.mark.parametrize("knob", ["RAYUELA_B200_ICM_JSPEC", "RAYUELA_B200_K1_V2"])
def test_icm_alternative_kernels_agree(rb, monkeypatch, knob):
    """RAYUELA_B200_ICM_JSPEC=0: the uniform row loop (zero diagonal blocks) for m <= 8; RAYUELA_B200_K1_V2=0: the
    round-1 unary kernel.  Same bits as the default kernels and the oracle."""
    r = np.random.default_rng(8)
    n, d, m = 4000, 64, 8
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 3).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    want = orc.encode_icm(X, C, B, 2, 4, 4, True, seed=4)
    for val in ("0", "1"):
        monkeypatch.setenv(knob, val)
        got = rb.core.encode_icm(X, C, B, 2, 4, 4, True, seed=4, want_cost=True)
        assert np.array_equal(got["B"], want["B"]), (knob, val)
        assert np.array_equal(bits(got["cost"]), bits(want["cost"])), (knob, val)
        U = np.asarray(rb.core.get_unaries(X[:300], C, m)).reshape(300, m, 256).transpose(1, 0, 2)   # -> (m, n, 256)
        assert np.array_equal(bits(np.ascontiguousarray(U)), bits(orc.get_unaries(X[:300], C, m))), (knob, val)


def test_scan_compat_symbols(rb):
    for kind, fn in ((orc.PQ, "pq"), (orc.LSQ, "lsq"), (orc.CQ, "cq")):
        B, Xq, cb, nrm = _scan_case(kind, 20000, 9, 8, 128, seed=kind)
        d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(kind, B, Xq, cb, 25, nrm)
        if kind == orc.PQ:
            d1, i1 = rb.core.c_linscan_aqd_query(B, Xq, cb, 25)
        elif kind == orc.LSQ:
            d1, i1 = rb.core.c_linscan_aqd_query_extra_byte(B, Xq, cb, nrm, 25)
        else:
            d1, i1 = rb.core.c_linscan_aqd_cq_query_extra_byte(B, Xq, cb, 25)
        assert np.array_equal(i1.astype(np.int64), i0.astype(np.int64))
        assert np.array_equal(bits(d1), bits(d0))


def test_scan_device_pointers_and_id_offset(rb):
    import torch
    B, Xq, cb, nrm = _scan_case(orc.LSQ, 60000, 20, 8, 128, seed=77)
    d0, i0 = orc.linscan(orc.LSQ, B, Xq, cb, 10, nrm, id_offset=5_000_000)
    ix = rb.core.Index(orc.LSQ, torch.from_numpy(B).cuda(), torch.from_numpy(nrm).cuda(), id_offset=5_000_000)
    d1, i1 = ix.search(torch.from_numpy(Xq).cuda(), torch.from_numpy(cb).cuda(), 10)
    assert np.array_equal(i1.cpu().numpy(), i0)
    assert np.array_equal(bits(d1.cpu().numpy()), bits(d0))


def test_base_sharded_scan_equals_single(rb):
    """Config-5 structure on one GPU: shard the base, search each shard with its id offset, merge."""
    B, Xq, cb, nrm = _scan_case(orc.LSQ, 90000, 25, 8, 64, seed=5, ties=True)
    k = 100
    d0, i0 = orc.linscan(orc.LSQ, B, Xq, cb, k, nrm)
    parts = [(0, 30000), (30000, 52000), (52000, 90000)]
    ds, is_ = [], []
    for a, b in parts:
        ix = rb.core.Index(orc.LSQ, B[a:b], nrm[a:b], id_offset=a)
        d, i = ix.search(Xq, cb, k)
        ds.append(d)
        is_.append(i)
    d1, i1 = rb.core.topk_merge(np.stack(ds), np.stack(is_))
    assert np.array_equal(i1, i0)
    assert np.array_equal(bits(d1), bits(d0))


def test_topk_merge_against_lexsort(rb):
    r = np.random.default_rng(0)
    S, nq, k = 5, 37, 64
    d = np.round(r.standard_normal((S, nq, k)) * 3).astype(np.float32)    # ties + negatives
    i = r.permutation(S * nq * k).reshape(S, nq, k).astype(np.int32)
    for s in range(S):                                                    # each list sorted by (dist, id)
        for q in range(nq):
            o = np.lexsort((i[s, q], d[s, q]))
            d[s, q], i[s, q] = d[s, q][o], i[s, q][o]
    dm, im = rb.core.topk_merge(d, i)
    for q in range(nq):
        dd, ii = d[:, q].reshape(-1), i[:, q].reshape(-1)
        o = np.lexsort((ii, dd))[:k]
        assert np.array_equal(im[q], ii[o]) and np.array_equal(bits(dm[q]), bits(dd[o]))


# ---- PQ / OPQ encode -----------------------------------------------------------------------------------------
def test_pq_encode_golden(rb):
    g = gold("pq_encode_m8")
    assert np.array_equal(rb.core.quantize_pq(g["X"], g["Cpq"], int(g["m"])), g["B_out"])


@pytest.mark.parametrize("n,m,sub", [(20000, 8, 16), (3333, 4, 8), (1000, 16, 2), (500, 2, 24)])
def test_pq_encode_matches_oracle(rb, n, m, sub):
    r = np.random.default_rng(n)
    X = r.standard_normal((n, m * sub)).astype(np.float32)
    Cpq = r.standard_normal((m * 256, sub)).astype(np.float32)
    assert np.array_equal(rb.core.quantize_pq(X, Cpq, m), orc.quantize_pq(X, Cpq, m))


# ---- the Julia-shaped API ---------------------------------------------------------------------------------
def test_julia_api_round_trip(rb):
    """Same call sequence as experiment_lsq / demos: encode -> qerror -> norms -> linscan -> recall."""
    n, d, m, nq, k = 4000, 32, 7, 50, 20
    r = np.random.default_rng(1)
    X = np.asfortranarray(r.standard_normal((d, n)).astype(np.float32))
    Xq = np.asfortranarray((X[:, :nq] + 0.01 * r.standard_normal((d, nq))).astype(np.float32))
    C = [np.asfortranarray((r.standard_normal((d, 256)) / 3).astype(np.float32)) for _ in range(m)]
    B0 = r.integers(1, 257, (m, n)).astype(np.int16)
    rb.seed_b200(99)
    oldB = B0.copy()
    B = rb.encoding_icm(X, oldB, C, 3, 4, True, 4, True, False)
    assert B.shape == (m, n) and B.dtype == np.int16 and B.min() >= 1 and B.max() <= 256
    assert np.array_equal(oldB, B)                                  # oldB mutated (src/LSQ.jl:248)
    Cimg = np.concatenate([c.T for c in C])
    o = orc.encode_icm(X.T, Cimg, (B0.T - 1).astype(np.uint8), 3, 4, 4, True, seed=99)
    assert np.array_equal(B.T - 1, o["B"])
    q = rb.qerror(X, B, C)
    assert abs(q - orc.qerror(X.T, o["B"], Cimg)) <= 1e-4 * q       # 1e-4 relative (north_star)
    rec = sum(c[:, B[i] - 1] for i, c in enumerate(C))
    dbnorms = (rec ** 2).sum(0).astype(np.float32)
    dists, res = rb.linscan_lsq(B, Xq, C, dbnorms, np.eye(d, dtype=np.float32), k)
    assert dists.shape == (k, nq) and res.shape == (k, nq) and res.min() >= 1
    d0, i0 = orc.linscan(orc.LSQ, o["B"], Xq.T, Cimg, k, dbnorms)
    assert np.array_equal(res.T.astype(np.int64), i0) and np.array_equal(bits(dists.T), bits(d0))
    recall = rb.eval_recall(np.arange(1, nq + 1), res, k, V=False)
    assert recall[-1] > 0.9 and np.allclose(recall, orc.eval_recall(np.arange(1, nq + 1), i0, k))
    Bs, objs = rb.encode_icm_cuda(X, B0.copy(), C, [1, 3], 4, 4, True, 2, False)
    assert len(Bs) == 2 and objs[1] < objs[0]


def test_errors_are_reported_not_swallowed(rb):
    X, C, B = _icm_data(10, 8, 2, seed=0)
    with pytest.raises(rb.RayuelaError):
        rb.core.encode_icm(X, C[:, :4].copy(), B, 1, 1, 1, False)       # wrong codebook shape
    with pytest.raises(rb.RayuelaError):
        rb.core.Index(orc.LSQ, B, None)                                 # LSQ scan needs norms
    ix = rb.core.Index(orc.CQ, B)
    with pytest.raises(rb.RayuelaError):
        ix.search(X, C, 11)                                             # k > n


# ---- "next" row 1: codebook update --------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,m", [(10000, 32, 4), (30011, 128, 8), (700, 20, 16), (5, 3, 1)])
def test_fast_bin_matmul_bit_exact(rb, n, d, m):
    """A (exact counts + rho) and b (Float64, ascending-l accumulation) are bit-identical to the oracle."""
    r = np.random.default_rng(n)
    X = (r.random((n, d)) * 10).astype(np.float32)               # test/common.jl:2-8
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    A0, b0 = orc.fast_bin_matmul(X, B)
    A1, b1 = rb.core.fast_bin_matmul(X, B)
    assert np.array_equal(A1, A0)
    assert np.array_equal(b1.view(np.uint64), b0.view(np.uint64))
    import torch
    A2, b2 = rb.core.fast_bin_matmul(torch.from_numpy(X).cuda(), torch.from_numpy(B).cuda())
    assert np.array_equal(A2.cpu().numpy(), A0) and np.array_equal(b2.cpu().numpy().view(np.uint64), b0.view(np.uint64))


def test_update_codebooks_and_reference_own_test(rb):
    """test/chainq.jl:2-11: update_codebooks_fast_bin ~ update_codebooks_fast_bin2 (d=32, n=10k, m=4, h=256)."""
    d, n, m, h = 32, 10000, 4, 256
    r = np.random.default_rng(1)
    X = np.asfortranarray((r.random((d, n)) * 10).astype(np.float32))
    B = r.integers(1, h + 1, (m, n)).astype(np.int16)
    C1 = rb.update_codebooks_fast_bin(X, B, h, False, 1e-4)
    A, b = rb.core.fast_bin_matmul(np.ascontiguousarray(X.T), (B.T - 1).astype(np.uint8), 1e-4)
    C2 = (np.linalg.inv(A) @ b.T).astype(np.float32)               # update_codebooks_fast_bin2, :209-229
    C1img = np.concatenate([c.T for c in C1])
    assert np.allclose(C1img, C2, rtol=1e-4, atol=1e-5)            # Julia's isapprox default is rtol = sqrt(eps)
    C0 = orc.update_codebooks_fast_bin(X.T, (B.T - 1).astype(np.uint8))
    assert np.allclose(C1img, C0, rtol=1e-6, atol=1e-7)            # same A, b, same LAPACK routine
    assert len(C1) == m and C1[0].shape == (d, h)


def test_train_lsq_and_sr_run_end_to_end(rb):
    """The alternation of src/LSQ.jl:323-372 / src/SR.jl:88-175 through the GPU encoder and codebook update."""
    d, n, m, h = 16, 6000, 4, 256
    r = np.random.default_rng(2)
    basis = r.standard_normal((d, d)) * (np.arange(1, d + 1) ** -0.7)[:, None]
    X = np.asfortranarray((basis.T @ r.standard_normal((d, n))).astype(np.float32))
    B0 = r.integers(1, h + 1, (m, n)).astype(np.int16)
    C0 = [np.zeros((d, h), dtype=np.float32) for _ in range(m)]
    R = np.eye(d, dtype=np.float32)
    rb.seed_b200(7)
    C, B, obj = rb.train_lsq(X, m, h, R, B0, C0, 4, 2, 2, True, 2, True, False)
    assert B.shape == (m, n) and len(C) == m and obj.shape == (4,)
    assert obj[-1] < obj[0] and obj[0] < 0.6 * float((X ** 2).sum(0).mean())
    C2, B2, obj2 = rb.train_sr_cuda(X, m, h, R, B0, C0, 4, 2, 2, True, 2, "SR_D", 1, 0.5, 1, False)
    assert obj2.shape == (5,) and obj2[-1] < obj2[0]
    C3, B3, obj3 = rb.train_sr_cuda(X, m, h, R, B0, C0, 3, 2, 2, True, 2, "SR_C", 1, 0.5, 1, False)
    assert obj3[-1] < obj3[0]


# ---- "next" row 2: norm quantization, and the demo pipeline around it ---------------------------------------
@pytest.mark.parametrize("n,d,m", [(20000, 128, 7), (3001, 30, 16), (1, 8, 1)])
def test_quantize_norms_bit_exact(rb, n, d, m):
    r = np.random.default_rng(n + m)
    C = r.standard_normal((m * 256, d)).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    cb = np.sort(r.random(256) * 4 * d).astype(np.float32)
    cb[10] = cb[11]                                         # duplicate centroid: first-min must pick the lower one
    c0, n0 = orc.quantize_norms(B, C, cb)
    c1, n1 = rb.core.quantize_norms(B, C, cb)
    assert np.array_equal(bits(n1), bits(n0)) and np.array_equal(c1, c0)
    _, n2 = rb.core.quantize_norms(B, C)
    assert np.array_equal(bits(n2), bits(n0))


def test_lsq_demo_pipeline(rb):
    """experiment_lsq_cuda's tail (src/LSQ_GPU.jl:348-366): norms codebook -> encode base -> quantize_norms ->
    linscan_lsq with the QUANTISED norms -> eval_recall."""
    d, n, m, h, nq, k = 32, 8000, 7, 256, 64, 50
    r = np.random.default_rng(4)
    basis = r.standard_normal((d, d)) * (np.arange(1, d + 1) ** -0.8)[:, None]
    X = np.asfortranarray((basis.T @ r.standard_normal((d, n))).astype(np.float32))
    Xq = np.asfortranarray((X[:, :nq] + 0.02 * r.standard_normal((d, nq))).astype(np.float32))
    rb.seed_b200(11)
    C0 = [np.zeros((d, h), dtype=np.float32) for _ in range(m)]
    C, B, _ = rb.train_lsq(X, m, h, np.eye(d, dtype=np.float32), r.integers(1, h + 1, (m, n)).astype(np.int16), C0,
                           3, 2, 3, True, 2, True, False)
    norms_B, norms_C = rb.get_norms_codebook(B, C)
    assert norms_C.shape == (h,) and norms_B.min() >= 1 and norms_B.max() <= h
    B_base = rb.encode_icm_cuda(X, r.integers(1, h + 1, (m, n)).astype(np.int16), C, [8], 3, 2, True, 1, False)[0][-1]
    base_norms_B, db_norms_X = rb.quantize_norms(B_base, C, norms_C)
    Cimg = np.concatenate([c.T for c in C])
    c0, n0 = orc.quantize_norms((B_base.T - 1).astype(np.uint8), Cimg, norms_C)
    assert np.array_equal(base_norms_B - 1, c0) and np.array_equal(bits(db_norms_X), bits(n0))
    db_norms = norms_C[base_norms_B - 1]
    dists, idx = rb.linscan_lsq(B_base, Xq, C, db_norms, np.eye(d, dtype=np.float32), k)
    d0, i0 = orc.linscan(orc.LSQ, (B_base.T - 1).astype(np.uint8), Xq.T, Cimg, k, db_norms)
    assert np.array_equal(idx.T.astype(np.int64), i0) and np.array_equal(bits(dists.T), bits(d0))
    recall = rb.eval_recall(np.arange(1, nq + 1, dtype=np.uint32), idx, k, V=False)
    assert recall[-1] > 0.8


# ---- "next" row 3: ChainQ Viterbi encode -----------------------------------------------------------------------
@pytest.mark.parametrize("n,d,m,kind", [(1000, 32, 4, "uniform"),      # test/chainq.jl:27-39
                                        (5003, 128, 8, "gauss"), (300, 24, 16, "gauss"), (9, 8, 1, "gauss"),
                                        (77, 16, 2, "uniform")])
def test_chainq_viterbi_exact(rb, n, d, m, kind):
    """The reference asserts Julia == CUDA == C++ codes exactly (test/chainq.jl:27-39); same bar here against the
    oracle, which is itself pinned to the reference's compiled viterbi_encoding."""
    X, C, _ = _icm_data(n, d, m, seed=900 + n, kind=kind)
    # m = 1 is degenerate and the reference's C++ reads an uninitialised OpenMP-private `mincost` there
    # (deps/src/encode_icm.cpp:81,123-125), so only the restatement is a valid checker for it
    want = orc.quantize_chainq(X, C, m, use_ref=orc.have_ref() and m > 1)
    got = rb.core.quantize_chainq(X, C, m)
    assert np.array_equal(got, want)


def test_viterbi_compat_symbol_and_julia_api(rb):
    n, d, m = 400, 32, 4
    X, C, _ = _icm_data(n, d, m, seed=5, kind="uniform")
    U = orc.get_unaries(X, C, m)                                            # [m][n][h]
    U2 = np.ascontiguousarray(U.transpose(1, 0, 2).reshape(n, m * 256))     # vcat(unaries...) image
    bins, _, cbi = orc.get_binaries(C, m)
    chain = np.stack([bins[[tuple(p) for p in cbi.tolist()].index((i, i + 1))] for i in range(m - 1)])
    want = orc.viterbi_encoding(U2, chain, m, use_ref=orc.have_ref())
    assert np.array_equal(rb.core.c_viterbi_encoding(U2, chain, m), want)
    Cs = [np.asfortranarray(C[i * 256:(i + 1) * 256].T) for i in range(m)]
    B, secs = rb.quantize_chainq(np.asfortranarray(X.T), Cs, True, False)
    assert B.shape == (m, n) and B.dtype == np.int16 and np.array_equal(B.T - 1, want) and secs >= 0
