"""CPU check of the inequality the scan's quantised pre-filter (csrc/linscan.cu: lut_quant_kernel, scanx_kernel<.., QPF>)
relies on, with the kernels' own arithmetic restated in numpy (fp32 where the kernels compute in fp32, double where they
use double):

    v[k][c]  = rint((LUT[k][c] - lo_k) / s)  (+ 1 for k = 0)                    lut_quant_kernel
    nq(i)    = min(rint(norm_i * (1/s) - K), cap),  K = floor(min norm * (1/s))   the hot loop's magic-number conversion
    A(i)     = sum_k v[k][code_k(i)] + nq(i)                                    exact integer sums (two per fp32 word)
    T(tau)   = floor((tau - off) / s) + mu,  off = sum_k lo_k + K / (1/s),
               mu = ceil(0.5 (m + 1) + 0.1) + 1 + ceil((m + 1) 2^-23 (sum_k max|LUT_k| + max|norm|) / s)

    claim: every code whose exact fp32 distance E(i) (ascending-k chain from 0, norm last: pairwise_byte.cpp:70-74)
           satisfies E(i) <= tau has A(i) <= T(tau); and A stays below 2^11 (no carry between the packed digits).

The kernel concludes nothing else from A: survivors are re-evaluated with the exact chain."""
import numpy as np
import pytest

f32 = np.float32


def _quantise(LUT, nrm, m, h):
    """One query's tables (m, h) + the base's norms (or None) -> v, nq(), off, mu, s exactly as the kernels form them."""
    lo = LUT.min(1)
    hi = LUT.max(1)
    rng = float((hi.astype(np.float64) - lo.astype(np.float64)).sum())
    has_norms = nrm is not None
    if has_norms:
        fin = nrm[np.isfinite(nrm)].astype(np.float64)
        nmin, nmax = f32(fin.min()), f32(fin.max())
        cap = fin.mean() + 4.0 * fin.std()
        ncap = f32(cap) if (nmin < cap < nmax) else nmax
        nrange = float(ncap) - float(nmin)
    else:
        nmin = nmax = ncap = f32(0)
        nrange = 0.0
    s = f32((rng + nrange) / 2000.0)
    if not (s > 0) or not np.isfinite(s):
        s = f32(1.0)
    inv = f32(f32(1.0) / s)
    K = np.floor(float(nmin) * float(inv)) if has_norms else 0.0
    cap_units = f32(np.ceil(float(ncap) * float(inv) - K) + 1.0) if has_norms else f32(0)
    v = np.rint(((LUT - lo[:, None]).astype(np.float32) / s).astype(np.float32))
    v = np.minimum(np.maximum(v, 0), 2010).astype(np.int64)
    v[0] += 1
    bmax = float(np.maximum(np.abs(lo), np.abs(hi)).astype(np.float64).sum())
    nb = max(abs(float(nmin)), abs(float(nmax))) if has_norms else 0.0
    fp = np.ceil((m + 1) * 2.0 ** -23 * (bmax + nb) / float(s))
    mu = int(np.ceil(0.5 * (m + (1 if has_norms else 0)) + 0.1)) + 1 + int(min(fp, 40000.0))
    off = float(lo.astype(np.float64).sum()) + (K / float(inv) if has_norms else 0.0)

    def nq(x):
        # fma(norm, 1/s, 1.5 * 2^23 - K): exact product and sum, ONE rounding to an integer; then - 1.5 * 2^23, min(cap)
        t = np.rint(x.astype(np.float64) * float(inv) - K)
        return np.minimum(t, float(cap_units)).astype(np.int64)

    return v, nq, off, mu, s


def _exact(LUT, codes, nrm):
    d = np.zeros(codes.shape[0], dtype=np.float32)
    for k in range(LUT.shape[0]):
        d = (d + LUT[k][codes[:, k]]).astype(np.float32)
    if nrm is not None:
        d = (d + nrm).astype(np.float32)
    return d


@pytest.mark.parametrize("m", [1, 5, 8, 16])
@pytest.mark.parametrize("kind", ["gauss", "outlier_norms", "offset", "flat", "integers", "no_norms", "tiny_h"])
def test_every_true_candidate_survives_the_prefilter(m, kind):
    r = np.random.default_rng(m * 100 + len(kind))
    n, h = 40000, 16 if kind == "tiny_h" else 256
    LUT = (r.standard_normal((m, h)) * 20).astype(np.float32)
    nrm = (r.standard_normal(n) * 3 + 10).astype(np.float32)
    codes = r.integers(0, h, (n, m))
    if kind == "outlier_norms":
        nrm[:5] = f32(1e6)
        nrm[5:9] = f32(-1e6)
    elif kind == "offset":
        LUT += f32(1e5)                       # fp32 rounding slack of the exact chain dominates the margin
        nrm += f32(3e5)
    elif kind == "flat":
        LUT[:] = f32(0.25)
        nrm[:] = f32(1.5)
    elif kind == "integers":
        LUT, nrm = np.round(LUT / 8), np.round(nrm)
    elif kind == "no_norms":
        nrm = None
    v, nq, off, mu, s = _quantise(LUT, nrm, m, h)
    A = sum(v[k][codes[:, k]] for k in range(m))
    if nrm is not None:
        q = nq(nrm)
        assert q.min() >= 0
        A = A + q
    assert A.min() >= 1 and A.max() < 2048                    # digits never carry; 0 is free for "nothing passes"
    E = _exact(LUT, codes, nrm)
    order = np.sort(E)
    for kth in (0, 9, 99, 999, n // 2, n - 1):
        tau = order[kth]
        T = np.floor((float(tau) - off) / float(s)) + mu
        T = 0 if not T >= 1 else min(T, 2046.0)
        assert (A[E <= tau] <= T).all(), (kind, m, kth)
    # and the filter is a filter: at the k = 1 threshold it lets through a small fraction only (not for degenerate data)
    if kind in ("gauss", "no_norms") and m >= 5:
        T = min(np.floor((float(order[0]) - off) / float(s)) + mu, 2046.0)
        assert (A <= T).mean() < 0.01
