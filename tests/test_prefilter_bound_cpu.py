"""CPU check of the inequality the ICM pre-filter (csrc/icm.cu, icm_warp_kernel<M, PF>) relies on, with the kernel's own
arithmetic restated in numpy: q = rint(T / scale_j) (16 bit), S(c) = rint(u(c) / scale_j) + sum_k q_k(c), and

    |scale_j * S(c) - exact(c)| <= delta = scale_j * ((m-1) * 0.51 + 0.75) + 2^-20 * (umax + (m-1) * tmax_j)

where exact(c) is the reference's fp32 chain ((u + r_1) + r_2) + ... (encode_icm.cpp:28-45).  It follows that the exact
first-minimum lies inside the window min(S) + 2*delta/scale_j -- the only thing the kernel concludes from S."""
import numpy as np
import pytest


def _case(m, seed, scale_u=1.0, scale_t=1.0, integers=False):
    r = np.random.default_rng(seed)
    nv = 400
    T = (r.standard_normal((m - 1, nv, 256)) * scale_t).astype(np.float32)     # the (m-1) gathered rows of one step
    u = (r.standard_normal((nv, 256)) * scale_u).astype(np.float32)
    if integers:
        T, u = np.round(T * 3), np.round(u * 3)
    return T, u


@pytest.mark.parametrize("m", [2, 7, 8, 16])
@pytest.mark.parametrize("kind", ["gauss", "big_unaries", "small_tables", "integers"])
def test_prefilter_window_contains_the_exact_argmin(m, kind):
    T, u = _case(m, seed=m * 10 + len(kind),
                 scale_u=40.0 if kind == "big_unaries" else 1.0,
                 scale_t=0.05 if kind == "small_tables" else 1.0,
                 integers=kind == "integers")
    f32 = np.float32
    tmax = f32(np.abs(T).max())
    scale = f32(tmax / f32(32767.0))                                  # __fdiv_rn(tmax, 32767)
    inv = f32(f32(32767.0) / tmax)                                    # __fdiv_rn(32767, tmax)
    q = np.rint((T / scale).astype(np.float32)).astype(np.int64)      # quant_tables_kernel
    assert np.abs(q).max() <= 32767
    # exact fp32 chain, ascending k
    exact = u.copy()
    for k in range(m - 1):
        exact = (exact + T[k]).astype(np.float32)
    # integer sums: unary rounded by fma(u, inv, 1.5 * 2^23) (exact product, one rounding to an integer)
    uq = np.rint(u.astype(np.float64) * np.float64(inv)).astype(np.int64)
    umax = np.abs(u).max(1, keepdims=True).astype(np.float64)
    on = (umax[:, 0] * np.float64(inv) * (2.002 * 2.0 ** -20) < 4.0)  # the kernel's enabling condition: umax/scale < ~2^21
    assert on.mean() > 0.5
    T, u, q, uq, exact, umax = T[:, on], u[on], q[:, on], uq[on], exact[on], umax[on]
    assert np.abs(uq).max() < 2 ** 21 + 2 ** 12
    S = uq + q.sum(0)
    delta = np.float64(scale) * ((m - 1) * 0.51 + 0.75) + 2.0 ** -20 * (umax + (m - 1) * np.float64(tmax))
    err = np.abs(np.float64(scale) * S - exact.astype(np.float64))
    assert (err <= delta).all(), float((err / delta).max())
    # window in units of scale_j as the kernel computes it: W = trunc(2.002 * delta_units) + 1
    slack = umax * (2.002 * 2.0 ** -20)
    w0 = 2.002 * ((m - 1) * 0.51 + 0.75 + 2.0 ** -20 * (m - 1) * 32767.0)
    W = np.floor(slack * np.float64(inv) + w0).astype(np.int64) + 1
    first_min = exact.argmin(1)                                       # first minimum, like encode_icm.cpp:47-58
    inside = S[np.arange(S.shape[0]), first_min] <= S.min(1) + W[:, 0]
    assert inside.all()
    # and when exactly one candidate is inside the window, it is that argmin
    cnt = (S <= (S.min(1) + W[:, 0])[:, None]).sum(1)
    unique = cnt == 1
    assert (S.argmin(1)[unique] == first_min[unique]).all()
    if kind == "gauss":
        assert unique.mean() > 0.9                                    # the filter decides almost every step
