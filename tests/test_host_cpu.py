"""CPU tests of host-side pieces that need no GPU: xvecs IO (the reference's own test/xvecs.jl), the 1-D
k-means used by get_norms_codebook, SR perturbation schedules, eval_recall, splitarray."""
import numpy as np
import pytest

import rayuela_b200 as rb
from oracle import oracle as orc


def test_xvecs_round_trip(tmp_path):
    """test/xvecs.jl:2-18 -- fvecs / ivecs write-read round trip, d=32, n=1000."""
    d, n = 32, 1000
    r = np.random.default_rng(0)
    X = (r.random((d, n)) * 10).astype(np.float32)
    fn = str(tmp_path / "x.fvecs")
    rb.fvecs_write(X, fn)
    X2 = rb.fvecs_read(n, fn)
    assert X2.shape == (d, n) and np.array_equal(X, X2)
    assert np.array_equal(rb.fvecs_read((11, 20), fn), X[:, 10:20])       # [a b] bounds, 1-based inclusive
    Xi = (np.floor(X - 0.5) * 1000).astype(np.int32)
    fi = str(tmp_path / "x.ivecs")
    rb.ivecs_write(Xi, fi)
    assert np.array_equal(rb.ivecs_read(n, fi), Xi)
    # bvecs: written by hand (the reference has no bvecs_write)
    Xb = r.integers(0, 256, (d, 7), dtype=np.uint8)
    fb = str(tmp_path / "x.bvecs")
    with open(fb, "wb") as f:
        for j in range(7):
            f.write(np.int32(d).tobytes() + Xb[:, j].tobytes())
    assert np.array_equal(rb.bvecs_read(None, fb), Xb)
    with pytest.raises(AssertionError):
        rb.fvecs_read(n + 1, fn)


def test_kmeans_1d_is_a_lloyd_fixed_point():
    from rayuela_b200.julia_api import kmeans_1d
    r = np.random.default_rng(1)
    x = np.concatenate([r.normal(mu, 0.05, 500) for mu in range(20)])
    a, c = kmeans_1d(x, 20, np.random.default_rng(3))
    assert a.min() == 0 and a.max() == 19 and np.all(np.diff(c) > 0)
    for i in range(20):                                   # centres are the means of their cells
        assert abs(x[a == i].mean() - c[i]) < 1e-4
    assert np.all(np.abs(x - c[a]) <= np.abs(x[:, None] - c[None]).min(1) + 1e-6)   # nearest-centre assignment


def test_apply_schedule_matches_reference_formulas():
    s = np.array([2.0, 4.0])
    assert np.allclose(rb.apply_schedule(s, 5, 25, 1, 0.5), s * (1 - 5 / 25) ** 0.5)    # SR_perturbations.jl:13-14
    assert np.allclose(rb.apply_schedule(s, 5, 25, 2, 0.5), s / (1 + 5) ** 0.5)         # :15-16
    assert np.allclose(rb.apply_schedule(s, 5, 25, 3, 0.5), s * 0.5 ** (5 / 2))         # :17-18
    with pytest.raises(rb.RayuelaError):
        rb.apply_schedule(s, 1, 2, 4)


def test_sr_perturbations_scale():
    r = np.random.default_rng(0)
    C = [r.standard_normal((8, 256)).astype(np.float32) for _ in range(4)]
    C2 = rb.SR_D_perturb(C, 1, 25, 1, 0.5, rng=np.random.default_rng(1))
    allc = np.concatenate(C, axis=1)
    want = allc.std(1, ddof=1) / 4 * (1 - 1 / 25) ** 0.5                  # SR_perturbations.jl:38-39
    got = np.concatenate([c2 - c for c, c2 in zip(C, C2)], axis=1).std(1)
    assert np.allclose(got, want, rtol=0.15)
    X = r.standard_normal((8, 5000)).astype(np.float32)
    Y = rb.SR_C_perturb(X, 25, 25, 1, 0.5, rng=np.random.default_rng(2))
    assert np.array_equal(X, Y)                                           # schedule 1 reaches zero noise at iter = niter


def test_quantize_norms_oracle_against_numpy():
    r = np.random.default_rng(5)
    n, d, m = 2000, 24, 5
    C = r.standard_normal((m * 256, d)).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    rec = sum(C.reshape(m, 256, d)[k][B[:, k]] for k in range(m)).astype(np.float64)
    cb = np.sort(r.random(256) * 60).astype(np.float32)
    codes, norms = orc.quantize_norms(B, C, cb)
    assert np.allclose(norms, (rec ** 2).sum(1), rtol=1e-5)
    assert (codes == np.abs(norms[:, None] - cb[None]).argmin(1)).mean() > 0.995


def test_sift1m_loader_round_trip(tmp_path):
    """read_dataset / load_experiment_data (src/read_datasets.jl:63-85, demos/experiment_utils.jl:63-86) on a tiny
    directory laid out like ./data/sift/: shapes, the +1 on the zero-based ground truth, range reads."""
    import numpy as np
    from rayuela_b200 import demos, xvecs
    r = np.random.default_rng(0)
    d, nt, nb, nq = 16, 50, 200, 7
    (tmp_path / "sift").mkdir()
    Xt = np.asfortranarray(r.random((d, nt)).astype(np.float32))
    Xb = np.asfortranarray(r.random((d, nb)).astype(np.float32))
    Xq = np.asfortranarray(r.random((d, nq)).astype(np.float32))
    gt0 = ((Xb.T[None, :, :] - Xq.T[:, None, :]) ** 2).sum(-1).argsort(1)[:, :5].T.astype(np.int32)   # 5-by-nq, 0-based
    xvecs.fvecs_write(Xt, str(tmp_path / "sift" / "sift_learn.fvecs"))
    xvecs.fvecs_write(Xb, str(tmp_path / "sift" / "sift_base.fvecs"))
    xvecs.fvecs_write(Xq, str(tmp_path / "sift" / "sift_query.fvecs"))
    xvecs.ivecs_write(gt0, str(tmp_path / "sift" / "sift_groundtruth.ivecs"))
    assert demos.have_sift1m(str(tmp_path)) and not demos.have_sift1m(str(tmp_path / "nope"))
    a, b, q, gt = demos.load_experiment_data("SIFT1M", 40, nb, nq, False, str(tmp_path))
    assert a.shape == (d, 40) and b.shape == (d, nb) and q.shape == (d, nq)
    assert np.array_equal(a, Xt[:, :40]) and np.array_equal(b, Xb)
    assert np.array_equal(gt, gt0[0] + 1)                       # a truncated base recomputes it: same answer here
    part = demos.read_dataset("SIFT1M_base", (11, 20), False, str(tmp_path))
    assert np.array_equal(part, Xb[:, 10:20])
