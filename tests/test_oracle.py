"""CPU tests: the oracle restatement against the reference's own compiled C++ (oracle/_ref) and
against independent numpy arithmetic.  No GPU."""
import numpy as np
import pytest

from oracle import oracle as orc

needs_ref = pytest.mark.skipif(not orc.have_ref() and not __import__("os").path.isdir("/root/reference"),
                               reason="oracle/_ref not built and /root/reference absent")


def _data(n, d, m, h=256, seed=0, kind="gauss"):
    r = np.random.default_rng(seed)
    if kind == "gauss":
        X = r.standard_normal((n, d)).astype(np.float32)
        C = (r.standard_normal((m * h, d)) / np.sqrt(m)).astype(np.float32)
    else:  # test/common.jl:2-8 fixture semantics: X = rand*10, C = rand
        X = (r.random((n, d)) * 10).astype(np.float32)
        C = r.random((m * h, d)).astype(np.float32)
    B = r.integers(0, h, (n, m), dtype=np.uint8)
    return X, C, B


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10 of zero counter/key, and of all-ones.
    out = orc.philox([0, 0, 0, 0], 0, 0)
    assert [hex(int(x)) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = orc.philox([0xFFFFFFFF] * 4, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [hex(int(x)) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_randperm_is_permutation_and_varies():
    seen = set()
    for it in range(50):
        p = orc.randperm(1234, it, 8)
        assert sorted(p.tolist()) == list(range(8))
        seen.add(tuple(p.tolist()))
    assert len(seen) > 40


def test_perturb_with_replacement_semantics():
    n, m = 20000, 8
    B = np.zeros((n, m), dtype=np.uint8)
    P = orc.perturb_codes(B + 7, 256, 4, seed=5, it=3)
    changed = (P != 7).sum(1)
    # with replacement: between 0 and 4 positions differ; some vectors must show a repeated position
    assert changed.max() <= 4 and (changed < 4).any() and (changed == 4).any()
    # positions uniform over m, values uniform over h
    pos_hist = (P != 7).sum(0) / n
    assert np.all(np.abs(pos_hist - (1 - (1 - 1 / m) ** 4)) < 0.02)
    # pure function of the global index: shard invariance
    P2 = orc.perturb_codes((B + 7)[5000:], 256, 4, seed=5, it=3, g0=5000)
    assert np.array_equal(P[5000:], P2)


@needs_ref
@pytest.mark.parametrize("m,kind", [(8, "gauss"), (7, "uniform"), (16, "gauss")])
def test_condition_matches_reference_bitwise(m, kind):
    orc.build()
    n, d, h = 3000, 32, 256
    X, C, B = _data(n, d, m, seed=m, kind=kind)
    U = orc.get_unaries(X, C, m)
    bins, bins_t, cbi = orc.get_binaries(C, m)
    pair2idx = np.zeros((m, m), dtype=np.int32)
    for i, (a, b) in enumerate(cbi):
        pair2idx[a, b] = pair2idx[b, a] = i
    B1, B2 = B.copy(), B.copy()
    for j in [3, 0, m - 1, 1]:
        tc = np.array([k for k in range(m) if k != j], dtype=np.int32)
        ub1, ub2 = U[j].copy(), U[j].copy()
        orc.condition(B1, ub1, bins, bins_t, pair2idx, tc, j, use_ref=False)
        orc.condition(B2, ub2, bins, bins_t, pair2idx, tc, j, use_ref=True)
        assert np.array_equal(B1, B2)
        assert np.array_equal(ub1.view(np.uint32), ub2.view(np.uint32))
    assert not np.array_equal(B1, B)


@needs_ref
def test_encode_icm_same_with_reference_step():
    orc.build()
    n, d, m = 2000, 32, 8
    X, C, B = _data(n, d, m, seed=11)
    a = orc.encode_icm(X, C, B, 3, 4, 4, True, seed=9)
    b = orc.encode_icm(X, C, B, 3, 4, 4, True, seed=9, use_ref_step=True)
    assert np.array_equal(a["B"], b["B"])
    assert np.array_equal(a["cost"].view(np.uint32), b["cost"].view(np.uint32))
    # ILS never makes a vector worse (strict-< accept, src/LSQ.jl:242)
    assert np.all(a["cost"] <= orc.veccost(X, B, C))
    assert orc.qerror(X, a["B"], C) < 0.8 * orc.qerror(X, B, C)


def test_unaries_binaries_against_numpy_blas():
    n, d, m, h = 500, 128, 4, 256
    X, C, _ = _data(n, d, m, seed=2)
    U = orc.get_unaries(X, C, m)
    Cm = C.reshape(m, h, d).astype(np.float64)
    Uref = -2 * np.einsum("mhd,nd->mnh", Cm, X.astype(np.float64)) + (Cm ** 2).sum(-1)[:, None, :]
    assert np.allclose(U, Uref, rtol=0, atol=2e-4)
    bins, bins_t, cbi = orc.get_binaries(C, m)
    for i, (a, b) in enumerate(cbi):
        ref = 2 * Cm[a] @ Cm[b].T          # [a_entry, b_entry]
        assert np.allclose(bins[i], ref.T, atol=1e-4)      # bin[idx][b*h + a]
        assert np.array_equal(bins_t[i], bins[i].T)
    assert cbi.tolist() == [[i, j] for i in range(m) for j in range(i + 1, m)]


def test_veccost_against_numpy():
    n, d, m, h = 1000, 64, 8, 256
    X, C, B = _data(n, d, m, seed=3)
    rec = sum(C.reshape(m, h, d)[k][B[:, k]] for k in range(m)).astype(np.float64)
    ref = ((rec - X) ** 2).sum(1)
    assert np.allclose(orc.veccost(X, B, C), ref, rtol=1e-5)


@needs_ref
@pytest.mark.parametrize("kind", [orc.LSQ, orc.CQ, orc.PQ])
@pytest.mark.parametrize("m,k", [(8, 1), (8, 100), (7, 33), (16, 10)])
def test_linscan_matches_reference_bitwise(kind, m, k):
    orc.build()
    n, nq, h = 20000, 12, 256
    d = 128 if kind != orc.PQ else 16 * m
    r = np.random.default_rng(100 * kind + m)
    B = r.integers(0, h, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    if kind == orc.PQ:
        cb = r.standard_normal((m * h, d // m)).astype(np.float32)
    else:
        cb = r.standard_normal((m * h, d)).astype(np.float32)
    # integer-valued data in half the cases -> many exactly tied distances (tie-break on idx)
    if m == 7:
        Xq, cb = np.round(Xq * 2), np.round(cb * 2)
        B[:, :] = r.integers(0, 4, (n, m), dtype=np.uint8)
    nrm = np.round(r.standard_normal(n) * 4).astype(np.float32) if kind == orc.LSQ else None
    d1, i1 = orc.linscan(kind, B, Xq, cb, k, nrm)
    d2, i2 = orc.ref_linscan(kind, B, Xq, cb, k, nrm)
    assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    assert np.array_equal(i1, i2)
    assert i1.min() >= (0 if kind == orc.PQ else 1)


def test_quantize_pq_against_numpy():
    n, m, h, sub = 2000, 8, 256, 16
    r = np.random.default_rng(4)
    X = r.standard_normal((n, m * sub)).astype(np.float32)
    Cpq = r.standard_normal((m * h, sub)).astype(np.float32)
    B = orc.quantize_pq(X, Cpq, m)
    Xs = X.reshape(n, m, sub).astype(np.float64)
    Cs = Cpq.reshape(m, h, sub).astype(np.float64)
    dd = ((Xs[:, :, None, :] - Cs[None]) ** 2).sum(-1)
    ref = dd.argmin(-1)
    agree = (ref == B).mean()
    assert agree > 0.999  # fp32 near-ties may flip; the chosen centroid must still be (near-)optimal
    chosen = np.take_along_axis(dd, B[..., None].astype(np.int64), -1)[..., 0]
    assert np.all(chosen - dd.min(-1) < 1e-3)


def test_eval_recall_counts_exactly_once():
    gt = np.array([5, 7, 9])
    idx = np.array([[5, 1, 2], [7, 7, 3], [1, 2, 9]])
    rec = orc.eval_recall(gt, idx, 3)
    assert rec.tolist() == [1 / 3, 1 / 3, 2 / 3]  # query 1 lists the NN twice -> miss (src/Linscan.jl:208-214)


def test_fast_bin_matmul_against_numpy():
    """fast_bin_matmul restatement vs an explicit one-hot matrix (test/common.jl data, test/chainq.jl:2-11 shape)."""
    n, d, m = 4000, 32, 4
    r = np.random.default_rng(6)
    X = (r.random((n, d)) * 10).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    A, b = orc.fast_bin_matmul(X, B)
    oh = np.zeros((n, m * 256))
    for i in range(m):
        oh[np.arange(n), i * 256 + B[:, i].astype(np.int64)] = 1
    assert np.array_equal(A, oh.T @ oh + 1e-4 * np.eye(m * 256))
    assert np.allclose(b.T, oh.T @ X.astype(np.float64), rtol=1e-12)
    # the reference's own test: update_codebooks_fast_bin ~ update_codebooks_fast_bin2 (inv(A)*b), chainq.jl:2-11
    C1 = orc.update_codebooks_fast_bin(X, B)
    C2 = (np.linalg.inv(A) @ b.T).astype(np.float32)
    assert np.allclose(C1, C2, rtol=1e-4, atol=1e-5)
    # and it is the least-squares minimiser: no worse than the mean-of-assigned-vectors initialisation
    assert np.linalg.norm(oh @ C1 - X) <= np.linalg.norm(oh @ C2 - X) * (1 + 1e-6)


@needs_ref
@pytest.mark.parametrize("m", [4, 8, 2])
def test_viterbi_matches_reference_bitwise(m):
    """test/chainq.jl:27-39 shape (d=32, n=1000, uniform data): restated Viterbi == the reference's own C++."""
    orc.build()
    n, d = 1000, 32
    X, C, _ = _data(n, d, m, seed=40 + m, kind="uniform")
    a = orc.quantize_chainq(X, C, m)
    b = orc.quantize_chainq(X, C, m, use_ref=True)
    assert np.array_equal(a, b)
    # Viterbi is exact: no single-code change can lower the chain objective
    U = orc.get_unaries(X[:50], C, m)                      # [m][n][h]
    Cm = C.reshape(m, 256, d).astype(np.float64)

    def energy(codes, l):
        e = sum(U[i, l, codes[i]] for i in range(m))
        return e + sum(2 * Cm[i, codes[i]] @ Cm[i + 1, codes[i + 1]] for i in range(m - 1))
    for l in range(0, 50, 7):
        base = energy(a[l], l)
        for i in range(m):
            for c in range(0, 256, 17):
                alt = a[l].copy()
                alt[i] = c
                assert energy(alt, l) >= base - 1e-2


# ---- the unpinned half: how far can the reference's (unknowable) BLAS summation order move the result? ----------
# Julia is not installed, so the association order of OpenBLAS sgemm inside get_unaries / get_binaries
# (src/utils.jl:135-136,164) cannot be reproduced; the oracle pins ONE order (sequential fmaf chains).  SURVEY 7:
# the same encode driven by BLAS-order unaries and tables (numpy float32 sgemm -- a different, blocked order) must
# give a code-mismatch rate far below 1e-3 and the same qerror within 1e-6 relative.
def test_blas_order_vs_fixed_order_statistics():
    orc.build()
    n, d, m = 100_000, 128, 8
    r = np.random.default_rng(2024)
    centres = r.standard_normal((512, d)).astype(np.float32)
    X = (centres[r.integers(0, 512, n)] + 0.5 * r.standard_normal((n, d))).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / np.sqrt(m)).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    U0, U1 = orc.get_unaries(X[:4096], C, m), orc.blas_unaries(X[:4096], C, m)
    assert not np.array_equal(U0, U1), "the two orders coincide: the test would prove nothing"
    assert np.abs(U0 - U1).max() < 1e-3 * np.abs(U0).max()
    fixed = orc.encode_icm(X, C, B, 4, 4, 4, True, seed=5, use_ref_step=orc.have_ref())
    blas = orc.encode_icm(X, C, B, 4, 4, 4, True, seed=5, use_ref_step=orc.have_ref(), blas=True)
    mismatch = float((fixed["B"] != blas["B"]).any(axis=1).mean())
    qf, qb = orc.qerror(X, fixed["B"], C), orc.qerror(X, blas["B"], C)
    print("BLAS-order vs fixed-order: vector mismatch rate %.2e, qerror %.8g vs %.8g" % (mismatch, qf, qb))
    assert mismatch < 2e-4                     # measured: 0 .. 3e-5 (a flipped near-tie cascades through the ILS)
    assert abs(qf - qb) <= 1e-6 * qf


def test_quantize_pq_vs_float64_argmin_bound():
    """quantize_pq is parity-unpinned (Distances.jl / Clustering.jl are not vendored).  Bound on what any fp32 order
    can change: against an exact float64 argmin the restated fp32 formula differs on < 2e-3 of the codes, and every
    differing code is a near-tie (relative distance gap < 1e-5)."""
    orc.build()
    n, m, sub = 50_000, 8, 16
    r = np.random.default_rng(3)
    X = r.standard_normal((n, m * sub)).astype(np.float32)
    Cpq = r.standard_normal((m * 256, sub)).astype(np.float32)
    B = orc.quantize_pq(X, Cpq, m)
    bad = 0
    for j in range(m):
        xs = X[:, j * sub:(j + 1) * sub].astype(np.float64)
        cs = Cpq[j * 256:(j + 1) * 256].astype(np.float64)
        d2 = (xs ** 2).sum(1)[:, None] + (cs ** 2).sum(1)[None, :] - 2 * xs @ cs.T
        ref = d2.argmin(1)
        diff = np.nonzero(ref != B[:, j])[0]
        bad += diff.size
        gap = d2[diff, B[diff, j]] - d2[diff, ref[diff]]
        assert np.all(gap <= 1e-5 * np.maximum(d2[diff, ref[diff]], 1e-30))
    assert bad < 2e-3 * n * m
