"""CPU model of scanx_kernel's verified speculative threshold (csrc/linscan.cu): keys stream by in periods, keys <= tau
are appended, a compaction keeps the r smallest -- r = x + 5.5 sqrt(x) + 15, x = k*f, or k without speculation -- and
makes the r-th the new threshold (ONE selection per compaction).
Invariant under test: thresholds only decrease and every key <= the current threshold that has been seen is in the
buffer, so  k-th smallest of the final buffer <= final tau  ==>  the buffer's k smallest ARE the true top-k;
and when the stream is ordered best-first the check must fail (the kernel then redoes the block exactly)."""
import numpy as np
import pytest


def stream_topk(keys, k, soft, period=256, spec=True):
    n = len(keys)
    buf, tau, seen = [], np.inf, 0
    for s in range(0, n, period):
        chunk = keys[s:s + period]
        buf.extend(chunk[chunk <= tau].tolist())
        seen += len(chunk)
        softq = soft
        if len(buf) > softq or (len(buf) >= k and tau == np.inf):
            buf.sort()
            if len(buf) > k:
                x = k * min(1.0, seen / n)
                r = int(x + 5.5 * np.sqrt(x) + 15.125) + 1
                r = r if (spec and r * 4 < k * 3) else k
                tau = min(tau, buf[r - 1])
                buf = buf[:r]
            elif len(buf) == k:
                tau = min(tau, buf[-1])
    buf.sort()
    ok = len(buf) >= k and buf[k - 1] <= tau if np.isfinite(tau) else True
    return np.array(buf[:k]), ok


@pytest.mark.parametrize("k", [16, 100, 1000])
@pytest.mark.parametrize("order", ["random", "best_last", "best_first", "clustered"])
def test_verified_speculation_is_exact_or_says_so(k, order):
    r = np.random.default_rng(k)
    n = 60000
    keys = r.permutation(n * 4)[:n].astype(np.float64)              # distinct keys, like (dist, id)
    if order == "best_last":
        keys = np.sort(keys)[::-1].copy()
    elif order == "best_first":
        keys = np.sort(keys)
    elif order == "clustered":                                       # long runs of similar keys
        keys = np.concatenate([np.sort(c) for c in np.array_split(keys, 37)])
    want = np.sort(keys)[:k]
    got, ok = stream_topk(keys, k, soft=max(384, 4 * k))
    if ok:
        assert np.array_equal(got, want)
    if order == "random":
        assert ok                                                    # exchangeable data: the threshold holds
    if order == "best_first" and k >= 100:                           # (at k = 16 the safety margin leaves no room to speculate)
        assert not ok                                                # the sample is as unrepresentative as it gets
    # without speculation the stream is always exact and always verifies
    got0, ok0 = stream_topk(keys, k, soft=max(384, 4 * k), spec=False)
    assert ok0 and np.array_equal(got0, want)
