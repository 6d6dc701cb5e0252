"""CPU model of the certified early rejection in K3's cost evaluation (csrc/icm.cu, warp_cost): the ILS accept test
(src/LSQ.jl:242, strict <) needs the reference's SEQUENTIAL fp32 sum of the d squares only when it can be <= the current
cost.  The kernel first adds the same squares as a tree (4 per lane, the lanes' chunks in order, then a 5-step butterfly)
and skips the sequential chain when  tree > curcost * (1 + 1.3e-7 * (d + 32)).  For that to be safe the sequential sum
must then be strictly above curcost, i.e.  seq >= tree / (1 + 1.3e-7 * (d + 32))  must hold for every input."""
import numpy as np
import pytest

f32 = np.float32


def seq_sum(sq):
    acc = f32(0)
    for v in sq:
        acc = f32(acc + v)
    return acc


def tree_sum(sq, d):
    """warp_cost's tree: lane L owns t = 4L .. 4L+3 of every 128-chunk (d % 4 == 0) or t = L, L+32, .. (otherwise)."""
    part = np.zeros(32, f32)
    if d % 4 == 0:
        for t4 in range(0, d, 4):
            lane = (t4 % 128) // 4
            o = sq[t4:t4 + 4]
            part[lane] = f32(part[lane] + f32(f32(o[0] + o[1]) + f32(o[2] + o[3])))
    else:
        for t in range(d):
            part[t % 32] = f32(part[t % 32] + sq[t])
    off = 16
    while off:
        part = (part + part[np.arange(32) ^ off]).astype(f32)
        off >>= 1
    assert (part == part[0]).all()                      # the butterfly leaves the same bits in every lane
    return part[0]


@pytest.mark.parametrize("d", [4, 30, 64, 128, 132, 960, 2048])
@pytest.mark.parametrize("kind", ["uniform", "heavy_tail", "one_big", "ascending", "descending", "tiny"])
def test_sequential_sum_is_never_below_the_tree_sum_by_more_than_the_margin(d, kind):
    r = np.random.default_rng(d * 7 + len(kind))
    worst = 0.0
    for rep in range(60):
        if kind == "uniform":
            sq = r.random(d)
        elif kind == "heavy_tail":
            sq = np.exp(r.standard_normal(d) * 4)
        elif kind == "one_big":
            sq = r.random(d) * 1e-3
            sq[r.integers(d)] = 1e4
        elif kind == "ascending":
            sq = np.sort(r.random(d) ** 8)
        elif kind == "descending":
            sq = np.sort(r.random(d) ** 8)[::-1]
        else:
            sq = r.random(d) * 1e-40                    # subnormal squares: additions are still exact up to rounding
        sq = (sq.astype(f32) ** 1).astype(f32)
        s, t = seq_sum(sq), tree_sum(sq, d)
        margin = f32(f32(1.0) + f32(1.3e-7) * f32(d + 32))
        # the kernel's test with curcost = the largest value it would still reject against
        assert float(s) * float(margin) >= float(t), (kind, d, float(s), float(t))
        if t > 0:
            worst = max(worst, (float(t) - float(s)) / float(t))
    assert worst <= 0.5 * 1.3e-7 * (d + 32)             # the margin is at least twice what the data ever needs


def test_rejection_never_changes_a_decision():
    """Whenever the kernel's test fires, the sequential sum is strictly above curcost (neither `<` nor `==`)."""
    r = np.random.default_rng(5)
    d = 128
    fired = 0
    for rep in range(400):
        sq = (r.random(d) ** 2).astype(f32)
        s, t = seq_sum(sq), tree_sum(sq, d)
        for cur in (f32(float(s) * (1 - 3e-5)), f32(float(s) * (1 - 1e-6)), s, f32(float(s) * (1 + 1e-6))):
            thr = f32(cur * f32(f32(1.0) + f32(1.3e-7) * f32(d + 32)))
            if t > thr:
                fired += 1
                assert s > cur
    assert fired > 100                                   # clear rejections do skip the chain
    # NaN / inf on either side fail the test (the exact chain runs)
    assert not (f32(np.nan) > f32(1.0)) and not (f32(1.0) > f32(np.inf) * f32(1.00002))
