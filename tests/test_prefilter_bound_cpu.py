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
    Q = 8191.0                                                        # pf_q(m): 14-bit fields
    T, u = _case(m, seed=m * 10 + len(kind),
                 scale_u=40.0 if kind == "big_unaries" else 1.0,
                 scale_t=0.05 if kind == "small_tables" else 1.0,
                 integers=kind == "integers")
    f32 = np.float32
    tmax = f32(np.abs(T).max())
    scale = f32(tmax / f32(Q))                                        # __fdiv_rn(tmax, Q)
    inv = f32(f32(Q) / tmax)                                          # __fdiv_rn(Q, tmax)
    q = np.rint((T / scale).astype(np.float32)).astype(np.int64)      # quant_tables_kernel
    assert np.abs(q).max() <= Q
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
    w0 = 2.002 * ((m - 1) * 0.51 + 0.75 + 2.0 ** -20 * (m - 1) * Q)
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


@pytest.mark.parametrize("m", [2, 5, 7, 8, 12, 16])
def test_packed_14bit_sums_equal_the_plain_integer_sums(m):
    """icm_warp_kernel / pf_rows14: the rows are offset fields q + 8192 in 1..16383, two per 32-bit word; the rows of
    every group of four codebooks (k >> 2) are added as whole words (at most four rows each, no carry between the
    fields), the unary is a 16-bit field pair relative to the codebook minimum.  The kernel's split
    h = (xu >> 16) + sum_g (G_g >> 16), lo = (sum_g G_g + xu) - (h << 16) (mod 2^32) must reproduce the plain sums, and
    keys S * 8 + slot / the second-smallest test must name the unique window member."""
    r = np.random.default_rng(m)
    nv, off = 2000, 8192
    for j in range(m):
        q = r.integers(-8191, 8192, (m, nv, 2)).astype(np.int64)      # [k][vector][lo / hi candidate]
        q[:, :5] = 8191                                               # extreme fields: 4 * 16383 per group
        q[j] = -off                                                   # the diagonal row is stored as zero words
        uq = r.integers(0, 65536, (nv, 2)).astype(np.int64)
        uq[:5] = 65535
        words = ((q + off) | ((q + off)[..., 1:2] << 16))[..., 0].astype(np.uint64)     # lo | hi << 16
        xu = (uq[:, 0] | (uq[:, 1] << 16)).astype(np.uint64)
        h, t = xu >> 16, xu.copy()
        for g in range(0, m, 4):
            G = words[g:g + 4].sum(0) & 0xFFFFFFFF
            h, t = h + (G >> 16), t + G
        lo = (t - (h << 16)) & 0xFFFFFFFF
        want_lo = uq[:, 0] + (q[:, :, 0] + off).sum(0)
        want_hi = uq[:, 1] + (q[:, :, 1] + off).sum(0)
        assert (lo.astype(np.int64) == want_lo).all() and (h.astype(np.int64) == want_hi).all()
        assert want_hi.max() < 2 ** 19                                # keys S * 8 + slot, << 5 | lane stay below 2^27


def test_unary_fields_saturation_is_flagged_per_codebook():
    """K3 keeps rint(u / scale_j) - min as 16-bit fields; a codebook whose unaries span more than 65535 units for the
    vector cannot use them: the warp's window constant becomes +inf (every step of that codebook takes the exact path).
    Otherwise the fields are the exact integers up to the constant, so the argmin / window of the integer sums is
    unchanged."""
    r = np.random.default_rng(3)
    inv = np.float32(8191.0 / 3.7)
    for spread, want_sat in ((1.0, False), (40.0, True)):
        u = (r.standard_normal(256) * spread).astype(np.float32)
        qf = (u.astype(np.float64) * np.float64(inv) + 12582912.0).astype(np.float32)   # one fma: a single rounding
        q = qf.view(np.int32).astype(np.int64) - 0x4B400000
        assert (q == np.rint(u.astype(np.float64) * np.float64(inv))).all()
        sat = (q.max() - q.min()) > 65535
        assert sat == want_sat
        if not sat:
            fields = np.minimum(q - q.min(), 65535)
            assert (fields - fields.min() == q - q.min()).all()
