"""The OPT-IN tensor-core mode (RAYUELA_FAST_UNARIES: tcgen05 bf16x3 GEMM for the unaries).  It is validated by
tolerance and by encode-quality statistics, never by bit identity; the default (exact) kernel is checked bit-for-bit
against the oracle in the same test so the two are never confused."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rayuela_b200
    return rayuela_b200


def _data(n, d, m, seed):
    r = np.random.default_rng(seed)
    centres = r.standard_normal((64, d)).astype(np.float32) * 3
    X = (centres[r.integers(0, 64, n)] + r.standard_normal((n, d))).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) * 1.5).astype(np.float32)
    return X, C, r.integers(0, 256, (n, m), dtype=np.uint8)


@pytest.mark.parametrize("n,d,m", [(1000, 128, 8), (333, 64, 7), (129, 100, 16), (128, 8, 1), (5000, 128, 2)])
def test_unaries_exact_and_fast(rb, n, d, m):
    X, C, _ = _data(n, d, m, seed=n + d)
    want = orc.get_unaries(X, C, m)                                     # (m, n, 256)
    want = np.ascontiguousarray(want.transpose(1, 0, 2)).reshape(n, m * 256)
    exact = rb.core.get_unaries(X, C, m)
    assert np.array_equal(exact.view(np.uint32), want.view(np.uint32))  # exact kernel: bit-identical to the oracle
    fast = rb.core.get_unaries(X, C, m, fast=True)
    # bf16x3: every product within 2^-16 relative (the dropped lo*lo term and the rounding of lo), fp32 accumulation
    bound = 2.0 ** -15 * (np.abs(X) @ np.abs(C).T) * 2 + 1e-6 * np.abs(want)
    err = np.abs(fast.astype(np.float64) - want.astype(np.float64))
    assert np.all(err <= bound), float((err / np.maximum(bound, 1e-30)).max())
    assert not np.array_equal(fast, want) or d <= 8                     # it really is a different arithmetic


def test_fast_mode_encode_quality(rb):
    n, d, m = 20000, 128, 8
    X, C, B = _data(n, d, m, seed=7)
    want = orc.encode_icm(X, C, B, 4, 4, 4, True, seed=3, use_ref_step=orc.have_ref())
    exact = rb.core.encode_icm(X, C, B, 4, 4, 4, True, seed=3, want_cost=True)
    fast = rb.core.encode_icm(X, C, B, 4, 4, 4, True, seed=3, want_cost=True, fast=True)
    assert np.array_equal(exact["B"], want["B"])                        # the default is untouched by the fast mode
    mismatch = float((fast["B"] != want["B"]).any(axis=1).mean())
    qe, qf = float(exact["cost"].mean()), float(fast["cost"].mean())
    print("fast unaries: vector mismatch rate %.3e, qerror %.6f vs %.6f" % (mismatch, qf, qe))
    assert mismatch < 2e-2
    assert abs(qf - qe) <= 1e-4 * qe                                    # north_star tolerance on qerror
    # costs are still exact veccosts of the codes it returns
    assert np.array_equal(rb.core.veccost(X, fast["B"], C).view(np.uint32), fast["cost"].view(np.uint32))


@pytest.mark.parametrize("m,k", [(8, 1), (8, 100), (16, 10)])
def test_fast_lut_scan(rb, m, k):
    """RAYUELA_FAST_LUT: the LSQ lookup tables by the tcgen05 GEMM; scan / norm add / top-k unchanged.  Distances within
    1e-4 relative (north_star), neighbour lists agree except on near-ties, Recall@1 against the exact search >= 0.995;
    the default search stays bit-identical to the reference."""
    r = np.random.default_rng(m * 100 + k)
    n, nq, d = 60_000, 200, 128
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    C = (r.standard_normal((m * 256, d)) / np.sqrt(m)).astype(np.float32)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    rec = sum(C.reshape(m, 256, d)[j][B[:, j]] for j in range(m))
    nrm = (rec * rec).sum(1).astype(np.float32)
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(orc.LSQ, B, Xq, C, k, nrm)
    ix = rb.core.Index(rb.core.SCAN_LSQ, B, nrm)
    de, ie = ix.search(Xq, C, k)
    assert np.array_equal(ie, i0) and np.array_equal(de.view(np.uint32), d0.view(np.uint32))
    df, jf = ix.search(Xq, C, k, fast=True)
    scale = np.abs(d0).max()
    assert np.abs(np.sort(df, 1) - np.sort(d0, 1)).max() <= 1e-4 * scale
    assert (jf[:, 0] == i0[:, 0]).mean() >= 0.995
    agree = np.mean([len(set(a) & set(b)) / k for a, b in zip(jf, i0)])
    print("fast LUT: top-%d list overlap %.5f, top-1 agreement %.4f" % (k, agree, (jf[:, 0] == i0[:, 0]).mean()))
    assert agree >= 0.99


def test_fast_mode_falls_back_to_exact_when_unsupported(rb):
    """d > 128 is outside the tensor-core kernel's tile: the flag is ignored and the exact kernels run (same bits)."""
    X, C, B = _data(700, 160, 4, seed=1)
    a = rb.core.get_unaries(X, C, 4)
    b = rb.core.get_unaries(X, C, 4, fast=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ea = rb.core.encode_icm(X, C, B, 2, 2, 4, True, seed=5)["B"]
    eb = rb.core.encode_icm(X, C, B, 2, 2, 4, True, seed=5, fast=True)["B"]
    assert np.array_equal(ea, eb)


def test_phase_timers_through_the_abi(rb):
    """rayuela_encode_icm_timings: CUDA-event phase times of the last call that asked for stats."""
    X, C, B = _data(20000, 64, 8, seed=2)
    rb.core.encode_icm(X, C, B, 4, 4, 4, True, seed=1, want_stats=True)
    t = rb.core.last_icm_timings()
    assert t["total"] > 0 and t["icm"] > 0 and t["unaries"] > 0 and t["setup"] >= 0
    assert t["icm"] <= t["total"] * 1.01 and t["unaries"] + t["setup"] <= t["total"] * 1.01      # one chunk: no overlap
