"""The Julia-level API (rayuela_b200.julia_api: same names, positional arguments, shapes and index bases as the
Julia package) against the oracle: the PQ / OPQ / CQ entry points, an oracle-driven train_lsq alternation that the GPU
trainer must reproduce, the experiment_* drivers and the demos_train_query_base.jl pipeline."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

H = 256
GOLDEN = 0x9E3779B97F4A7C15


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rayuela_b200
    return rayuela_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _julia(a):
    """memory image (n, d) -> Julia-shaped d-by-n (same bytes)."""
    return np.asfortranarray(np.asarray(a).T)


def _pq_setup(n=6000, nq=40, d=64, m=8, seed=0):
    r = np.random.default_rng(seed)
    X = r.standard_normal((n, d)).astype(np.float32)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    Cpq = r.standard_normal((m * H, d // m)).astype(np.float32)          # image of cat(C..., dims=3)
    C = [_julia(Cpq[j * H:(j + 1) * H]) for j in range(m)]               # Julia: m matrices (d/m)-by-h
    R = np.linalg.qr(r.standard_normal((d, d)))[0].astype(np.float32)
    return X, Xq, Cpq, C, R


def _full(Cpq, m, d):
    """PQ blocks embedded as full-dimensional additive codebooks, image (m*h, d)."""
    sub = d // m
    out = np.zeros((m * H, d), dtype=np.float32)
    for j in range(m):
        out[j * H:(j + 1) * H, j * sub:(j + 1) * sub] = Cpq[j * H:(j + 1) * H]
    return out


def test_quantize_pq_opq_and_qerrors(rb):
    """quantize_pq (src/PQ.jl:18-48), quantize_opq (src/OPQ.jl:19-27), qerror_pq / qerror_opq (src/qerrors.jl:77-100)."""
    X, Xq, Cpq, C, R = _pq_setup()
    n, d = X.shape
    m = len(C)
    B = rb.quantize_pq(_julia(X), C)
    assert B.shape == (m, n) and B.dtype == np.int16
    assert np.array_equal(B.T - 1, orc.quantize_pq(X, Cpq, m))
    RX = (R.T @ _julia(X))                                               # what quantize_opq rotates (Julia shape)
    Bo = rb.quantize_opq(_julia(X), R, C)
    assert np.array_equal(Bo.T - 1, orc.quantize_pq(np.ascontiguousarray(RX.T), Cpq, m))
    want = orc.qerror(X, (B.T - 1).astype(np.uint8), _full(Cpq, m, d))
    assert abs(float(rb.qerror_pq(_julia(X), B, C)) - want) <= 1e-4 * want           # north_star: 1e-4 relative
    want_o = orc.qerror(np.ascontiguousarray(RX.T), (Bo.T - 1).astype(np.uint8), _full(Cpq, m, d))
    assert abs(float(rb.qerror_opq(_julia(X), Bo, C, R)) - want_o) <= 1e-4 * want_o
    assert want_o > 0 and want > 0


@pytest.mark.parametrize("k", [1, 100])
def test_linscan_pq_opq_cq_julia_api(rb, k):
    """linscan_pq / linscan_opq / linscan_cq (src/Linscan.jl:5-37,93-115,160-193): k-by-nq results, ONE-based ids,
    both the UInt8 (zero-based) and the Integer (one-based) code methods."""
    X, Xq, Cpq, C, R = _pq_setup(seed=k)
    n, d = X.shape
    m = len(C)
    B0 = orc.quantize_pq(X, Cpq, m)                                      # (n, m) uint8 zero-based
    B16 = _julia(B0).astype(np.int16) + 1                                # Julia m-by-n Int16 one-based
    ref = orc.ref_linscan if orc.have_ref() else orc.linscan
    d0, i0 = ref(orc.PQ, B0, Xq, Cpq, k)                                 # zero-based ids
    for codes in (B16, _julia(B0)):
        dists, res = rb.linscan_pq(codes, _julia(Xq), C, 8 * m, k)
        assert dists.shape == (k, Xq.shape[0]) and res.dtype == np.uint32
        assert np.array_equal(res.T, i0 + 1)                             # res .+= 1, src/Linscan.jl:25
        assert np.array_equal(bits(dists.T), bits(d0))
    with pytest.raises(rb.RayuelaError):
        rb.linscan_pq(B16, _julia(Xq), C, 8 * m + 1, k)
    # OPQ: rotate the queries, then PQ
    RXq = np.ascontiguousarray((R.T @ _julia(Xq)).T)
    d1, i1 = ref(orc.PQ, B0, RXq, Cpq, k)
    dists, res = rb.linscan_opq(B16, _julia(Xq), C, 8 * m, R, k)
    assert np.array_equal(res.T, i1 + 1) and np.array_equal(bits(dists.T), bits(d1))
    # CQ: full-dimensional codebooks, no norms
    r = np.random.default_rng(7)
    Ccq = r.standard_normal((m * H, d)).astype(np.float32)
    Bc = r.integers(0, H, (n, m), dtype=np.uint8)
    d2, i2 = ref(orc.CQ, Bc, Xq, Ccq, k)
    dists, res = rb.linscan_cq(_julia(Bc).astype(np.int16) + 1, _julia(Xq),
                               [_julia(Ccq[j * H:(j + 1) * H]) for j in range(m)], k)
    assert np.array_equal(res.T, i2) and np.array_equal(bits(dists.T), bits(d2))


def _oracle_train_lsq(X, B0, niter, ilsiter, icmiter, randord, npert, seed0):
    """train_lsq (src/LSQ.jl:323-372) restated on the oracle's pieces (R = I): update_codebooks_fast_bin, encoding_icm,
    then niter x {obj, update_codebooks, encoding_icm}.  Encode call number c draws the seed seed0 + c * GOLDEN, the
    stream julia_api.seed_b200 defines."""
    seeds = [(seed0 + c * GOLDEN) & 0xFFFFFFFFFFFFFFFF for c in range(niter + 1)]
    C = orc.update_codebooks_fast_bin(X, B0)                                          # :343-344
    B = orc.encode_icm(X, C, B0, ilsiter, icmiter, npert, randord, seed=seeds[0],
                       use_ref_step=orc.have_ref())["B"]                              # :351
    obj = []
    for it in range(niter):
        obj.append(orc.qerror(X, B, C))                                               # :357
        C = orc.update_codebooks_fast_bin(X, B)                                       # :361
        B = orc.encode_icm(X, C, B, ilsiter, icmiter, npert, randord, seed=seeds[it + 1],
                           use_ref_step=orc.have_ref())["B"]                          # :364
    return C, B, np.array(obj)


def test_train_lsq_matches_oracle_alternation(rb):
    """The GPU trainer against the same alternation driven entirely by the oracle: codes bit-exact after every one of
    the niter + 1 encodes (so also at the end), codebooks within 1e-6, objective within 1e-4 relative."""
    r = np.random.default_rng(11)
    n, d, m, niter = 4000, 32, 4, 3
    centres = r.standard_normal((64, d)).astype(np.float32) * 2
    X = (centres[r.integers(0, 64, n)] + r.standard_normal((n, d))).astype(np.float32)
    B0 = r.integers(0, H, (n, m), dtype=np.uint8)
    Co, Bo, objo = _oracle_train_lsq(X, B0, niter, 2, 3, True, 4, seed0=1234)
    rb.seed_b200(1234)
    C, B, obj = rb.train_lsq(_julia(X), m, H, np.eye(d, dtype=np.float32), _julia(B0).astype(np.int16) + 1,
                             None, niter, 2, 3, True, 4, True, False)
    assert np.array_equal(B.T - 1, Bo)
    Cimg = np.concatenate([c.T for c in C], axis=0)
    assert np.allclose(Cimg, Co, rtol=1e-6, atol=1e-6 * np.abs(Co).max())
    assert np.allclose(obj, objo, rtol=1e-4)
    assert obj[-1] < obj[0]


def test_experiment_sr_cuda_pipeline(rb):
    """experiment_sr_cuda (src/SR.jl:247-306): returns C, B, R, train_error, B_base, recall; every GPU product of the
    pipeline is re-derived from the oracle given the trained codebooks (base codes are not re-derivable -- the random
    initial codes come from the seeded host stream -- so the search half is checked on the returned B_base)."""
    Xt, Xb, Xq, gt = rb.demos.synthetic_sift(3000, 8000, 60, d=32, seed=3)
    m, niter, knn = 4, 2, 50
    r = np.random.default_rng(0)
    B = np.asfortranarray(r.integers(1, H + 1, (m, Xt.shape[1])).astype(np.int16))
    rb.seed_b200(99)
    R = np.eye(32, dtype=np.float32)
    C, Btr, R2, train_error, B_base, recall = rb.experiment_sr_cuda(
        Xt, B, None, R, Xb, Xq, gt, m, H, niter, 2, 2, True, 4, knn, 1, 1, "SR_D", 1, 0.5, False)
    assert len(C) == m and C[0].shape == (32, H) and Btr.shape == B.shape
    assert B_base.shape == (m, Xb.shape[1]) and B_base.min() >= 1 and B_base.max() <= H
    assert train_error.shape == (niter + 1,) and train_error[-1] < train_error[0]
    assert recall.shape == (knn,) and np.all(np.diff(recall) >= 0) and recall[-1] > 0.5
    # the search half, from the returned codes: norms -> linscan -> recall, against the oracle / reference C++
    Cimg = np.concatenate([np.asarray(c, dtype=np.float32).T for c in C], axis=0)
    codes0 = np.ascontiguousarray(B_base.T - 1).astype(np.uint8)
    _, norms = orc.quantize_norms(codes0, Cimg)
    nb, nc = rb.get_norms_codebook(Btr, C)
    ncodes, _ = orc.quantize_norms(codes0, Cimg, nc)
    db_norms = nc[ncodes.astype(np.int64)]
    ref = orc.ref_linscan if orc.have_ref() else orc.linscan
    d0, i0 = ref(orc.LSQ, codes0, np.ascontiguousarray(Xq.T), Cimg, knn, db_norms)
    dists, idx = rb.linscan_lsq(B_base, Xq, C, db_norms, R, knn)
    assert np.array_equal(idx.T.astype(np.int32), i0) and np.array_equal(bits(dists.T), bits(d0))
    assert norms.shape == (Xb.shape[1],)


def test_run_demos_acceptance(rb):
    """demos/demos_train_query_base.jl:9-105 on the synthetic SIFT-like set, reduced sizes: PQ -> OPQ -> OPQ(m-1) ->
    ChainQ -> LSQ -> LSQ++ SR-D / SR-C.  Acceptance = it runs end to end with the reference's call sequence and the
    methods rank the way the reference's README reports (additive codes beat the orthogonal ones on qerror)."""
    out = rb.demos.run_demos("synthetic", ntrain=4000, m=8, h=H, niter=3, nquery=100, nbase=10000, knn=100,
                             verbose=False, seed=5, ilsiter=2, icmiter=2)
    for name in ("pq", "opq", "lsq", "sr_d", "sr_c"):
        rec = out[name]["recall"]
        assert rec.shape == (100,) and np.all(np.diff(rec) >= 0)
        assert rec[-1] >= 0.6, (name, rec[-1])          # R@100 over a 10k base; measured 0.78 (LSQ, 7+1 bytes) .. 0.95
    errs = {k: float(out[k]["train_error"]) for k in ("pq", "opq", "chainq", "lsq", "sr_d", "sr_c")}
    print("train errors:", errs)
    assert errs["opq"] <= errs["pq"] * 1.02, errs
    assert errs["lsq"] < errs["chainq"], errs            # LSQ training continues to improve its ChainQ initialisation
    assert errs["sr_d"] < errs["chainq"] and errs["sr_c"] < errs["chainq"], errs
    assert out["sr_d"]["B_base"].shape == (7, 10000)


@pytest.mark.parametrize("h,m,d", [(16, 4, 24), (64, 8, 32), (100, 3, 17), (255, 2, 8)])
def test_encoding_icm_any_h(rb, h, m, d):
    """h != 256: the reference's cpp=false path iterated_conditional_modes! (src/LSQ.jl:83-149) -- codes, costs, stats,
    snapshots bit-exact against its oracle restatement; through core and through the Julia-level encoding_icm."""
    r = np.random.default_rng(h)
    n = 1500
    X = r.standard_normal((n, d)).astype(np.float32)
    C = r.standard_normal((m * h, d)).astype(np.float32)
    B = r.integers(0, h, (n, m), dtype=np.uint8)
    want = orc.encode_icm(X, C, B, 3, 2, 3, True, seed=21, g0=5, h=h, snap_iters=[2])
    got = rb.core.encode_icm(X, C, B, 3, 2, 3, True, seed=21, g0=5, h=h, snap_iters=[2], want_cost=True,
                             want_stats=True)
    assert np.array_equal(got["B"], want["B"]) and got["B"].max() < h
    assert np.array_equal(bits(got["cost"]), bits(want["cost"]))
    assert np.array_equal(got["stats"], want["stats"])
    assert np.array_equal(got["B_snap"], want["B_snap"])
    assert np.allclose(got["objs"], want["objs"], rtol=1e-4)
    assert np.array_equal(bits(rb.core.veccost(X, B, C, h=h)), bits(orc.veccost(X, B, C, h)))
    rb.seed_b200(21)
    Cj = [_julia(C[j * h:(j + 1) * h]) for j in range(m)]
    oldB = _julia(B).astype(np.int16) + 1
    Bj = rb.encoding_icm(_julia(X), oldB, Cj, 3, 2, True, 3, False, False)
    want0 = orc.encode_icm(X, C, B, 3, 2, 3, True, seed=21, g0=0, h=h)
    assert np.array_equal(Bj.T - 1, want0["B"]) and np.array_equal(oldB, Bj)          # oldB mutated, src/LSQ.jl:248
    assert abs(float(rb.qerror(_julia(X), Bj, Cj)) - orc.qerror(X, want0["B"], C, h)) < 1e-4 * orc.qerror(X, B, C, h)


@pytest.mark.parametrize("kind,h,m,k", [(orc.LSQ, 64, 8, 10), (orc.CQ, 100, 4, 3), (orc.LSQ, 16, 16, 100), (orc.CQ, 255, 8, 1)])
def test_linscan_any_h(rb, kind, h, m, k):
    """The reference's extra_byte scans take h as an argument (deps/src/linscan_aqd_pairwise_byte.cpp:14-24,97-106):
    codebooks with fewer than 256 entries scan bit-identically too (index, compat symbol and Julia-level API)."""
    r = np.random.default_rng(h + m)
    n, nq, d = 30_000, 21, 48
    B = r.integers(0, h, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * h, d)).astype(np.float32)
    nrm = r.standard_normal(n).astype(np.float32) if kind == orc.LSQ else None
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(kind, B, Xq, cb, k, nrm, h=h)
    ix = rb.core.Index(rb.core.SCAN_LSQ if kind == orc.LSQ else rb.core.SCAN_CQ, B, nrm, h=h)
    dg, ig = ix.search(Xq, cb, k)
    assert np.array_equal(ig, i0) and np.array_equal(bits(dg), bits(d0))
    Cj = [_julia(cb[j * h:(j + 1) * h]) for j in range(m)]
    if kind == orc.LSQ:
        dj, ij = rb.linscan_lsq(_julia(B), _julia(Xq), Cj, nrm, np.eye(d, dtype=np.float32), k)
        d2, i2 = rb.core.c_linscan_aqd_query_extra_byte(B, Xq, cb, nrm, k, h=h)
        assert np.array_equal(i2, i0) and np.array_equal(bits(d2), bits(d0))
    else:
        dj, ij = rb.linscan_cq(_julia(B), _julia(Xq), Cj, k)
    assert np.array_equal(ij.T.astype(np.int32), i0) and np.array_equal(bits(dj.T), bits(d0))


def test_high_recall_experiments_snapshots(rb):
    """demos/demos_train_query_base.jl:107-158: one encode with `ilsiters` snapshots; the error of the snapshots can only
    go down with more ILS iterations (strict-< accept, src/LSQ.jl:242), and each snapshot equals a separate encode with
    that many iterations (same seed => same perturbations)."""
    Xt, Xb, Xq, gt = rb.demos.synthetic_sift(2000, 6000, 40, d=32, seed=9)
    m = 4
    r = np.random.default_rng(1)
    B = np.asfortranarray(r.integers(1, H + 1, (m, Xt.shape[1])).astype(np.int16))
    rb.seed_b200(5)
    C, B, _ = rb.train_lsq(Xt, m, H, np.eye(32, dtype=np.float32), B, None, 2, 2, 2, True, 4, True, False)
    rb.seed_b200(77)
    out = rb.demos.high_recall_experiments(C, B, Xb, Xq, gt, m, H, ilsiters=(1, 2, 4, 8), knn=20, V=False)
    errs = [out[i]["base_error"] for i in (1, 2, 4, 8)]
    assert all(a >= b for a, b in zip(errs, errs[1:])) and errs[-1] < errs[0]
    assert all(out[i]["recall"].shape == (20,) for i in out)
