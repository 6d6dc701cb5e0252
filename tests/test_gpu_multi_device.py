"""Multi-GPU inside ONE process (rayuela_init / RAYUELA_B200_DEVICES), the general top-k merge, and the scan's
edge cases fixed in round 2.  The device set [0, 0] puts two shards on one GPU, so every sharded path runs on a
1-GPU box; with two GPUs visible the same tests also run on [0, 1].
Bar: codes / ids / distances bit-identical to the single-device call and to the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rayuela_b200
    yield rayuela_b200
    rayuela_b200.init(None)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def device_sets():
    import torch
    sets = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append([0, 1])
    return sets


def _icm_data(n, d, m, seed):
    r = np.random.default_rng(seed)
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / np.sqrt(m)).astype(np.float32)
    return X, C, r.integers(0, 256, (n, m), dtype=np.uint8)


@pytest.mark.parametrize("n,m", [(6001, 8), (4100, 16)])
def test_multi_device_encode_equals_single_and_oracle(rb, n, m):
    X, C, B = _icm_data(n, 64, m, seed=n)
    snaps = [1, 3]
    want = orc.encode_icm(X, C, B, 3, 4, 4, True, seed=9, g0=77, snap_iters=snaps)
    rb.init(None)
    single = rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=9, g0=77, snap_iters=snaps, want_cost=True,
                                want_stats=True)
    assert rb.device_count() == 1
    for devs in device_sets():
        rb.init(devs)
        assert rb.device_count() == len(devs)
        got = rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=9, g0=77, snap_iters=snaps, want_cost=True,
                                 want_stats=True)
        for r in (single, got):
            assert np.array_equal(r["B"], want["B"])
            assert np.array_equal(bits(r["cost"]), bits(want["cost"]))
            assert np.array_equal(r["stats"], want["stats"])
            assert np.array_equal(r["B_snap"], want["B_snap"])
            assert np.allclose(r["objs"], want["objs"], rtol=1e-6)
        assert rb.core.last_icm_steps() == (rb.core.last_icm_steps()[0], n * 3 * 4 * m)
    rb.init(None)


@pytest.mark.parametrize("kind,m,k", [(orc.LSQ, 8, 10), (orc.PQ, 8, 100), (orc.CQ, 16, 1), (orc.LSQ, 7, 1000)])
def test_multi_device_index_equals_reference(rb, kind, m, k):
    r = np.random.default_rng(5 + m + k)
    n, nq, d = 40_000 + 37, 50, 64
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d // m if kind == orc.PQ else d)).astype(np.float32)
    if kind == orc.PQ:
        Xq = np.round(Xq * 2)                        # integer data: exact ties across shards
        cb = np.round(cb * 2)
    nrm = r.standard_normal(n).astype(np.float32) if kind == orc.LSQ else None
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(kind, B, Xq, cb, k, nrm)
    ckind = {orc.LSQ: rb.core.SCAN_LSQ, orc.CQ: rb.core.SCAN_CQ, orc.PQ: rb.core.SCAN_PQ}[kind]
    for devs in device_sets():
        rb.init(devs)
        ix = rb.core.Index(ckind, B, nrm)
        dg, ig = ix.search(Xq, cb, k)
        ix.free()
        assert np.array_equal(ig, i0), devs
        assert np.array_equal(bits(dg), bits(d0)), devs
    rb.init(None)


def test_devices_from_environment_variable():
    """RAYUELA_B200_DEVICES is read when rayuela_init was never called (fresh process)."""
    code = ("import sys; sys.path[:0] = %r\n"
            "import numpy as np, rayuela_b200 as rb\n"
            "from oracle import oracle as orc\n"
            "assert rb.device_count() == 2\n"
            "r = np.random.default_rng(1); n, m = 5000, 8\n"
            "X = r.standard_normal((n, 32)).astype(np.float32)\n"
            "C = r.standard_normal((m * 256, 32)).astype(np.float32)\n"
            "B = r.integers(0, 256, (n, m), dtype=np.uint8)\n"
            "got = rb.core.encode_icm(X, C, B, 2, 2, 4, True, seed=3)['B']\n"
            "assert np.array_equal(got, orc.encode_icm(X, C, B, 2, 2, 4, True, seed=3)['B'])\n"
            "print('ok')\n") % (sys.path[:3],)
    env = dict(os.environ, RAYUELA_B200_DEVICES="0,0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("S,nq,k", [(2, 7, 10000), (8, 5, 2500), (3, 4, 6000), (5, 3, 3277), (2, 3, 8192),
                                    (8, 6, 1000), (3, 5, 700), (7, 4, 333), (2, 9, 4000), (16, 3, 512)])
def test_topk_merge_beyond_one_block(rb, S, nq, k):
    """Merges of sorted lists beyond the small bitonic case: the rank-merge tree in shared memory (more than 1024 keys per
    query, e.g. 8 GPUs at k = 1000) and in global memory (more than ~16k keys per query, ADVICE r1: 2 shards at the
    reference's default k = 10000); heavy ties in the distances."""
    r = np.random.default_rng(S * k)
    d = np.round(r.standard_normal((S, nq, k)) * 20).astype(np.float32) + np.float32(0)   # no -0.0: keys canonicalise it
    i = r.permutation(S * nq * k).astype(np.int32).reshape(S, nq, k)        # distinct ids
    order = np.lexsort((i, d), axis=-1)                                     # each list sorted by (dist, id)
    d = np.take_along_axis(d, order, -1)
    i = np.take_along_axis(i, order, -1)
    dg, ig = rb.core.topk_merge(d, i)
    for q in range(nq):
        dd, ii = d[:, q].reshape(-1), i[:, q].reshape(-1)
        o = np.lexsort((ii, dd))[:k]
        assert np.array_equal(ig[q], ii[o])
        assert np.array_equal(bits(dg[q]), bits(dd[o]))


@pytest.mark.parametrize("kind,m,nq", [(orc.PQ, 8, 17), (orc.CQ, 8, 1), (orc.CQ, 16, 9), (orc.PQ, 4, 33)])
def test_scan_ragged_last_query_tile(rb, kind, m, nq):
    """nq not a multiple of the 16 / 8 queries of a block: the padded dummy queries (all-zero LUT, every distance 0)
    must not append anything -- results unchanged, and the scan does not crawl."""
    r = np.random.default_rng(nq)
    n, d, k = 50_000, 32, 5
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d // m if kind == orc.PQ else d)).astype(np.float32)
    d0, i0 = (orc.ref_linscan if orc.have_ref() else orc.linscan)(kind, B, Xq, cb, k)
    ckind = {orc.CQ: rb.core.SCAN_CQ, orc.PQ: rb.core.SCAN_PQ}[kind]
    ix = rb.core.Index(ckind, B)
    dg, ig = ix.search(Xq, cb, k)
    assert np.array_equal(ig, i0) and np.array_equal(bits(dg), bits(d0))


def test_scan_rejects_non_finite_lookup_tables(rb):
    """One inf / NaN codebook entry: the reference confines it to the codes that use the entry; the GPU scan's
    accumulator restart (x * 0) cannot, so the search fails loudly instead of dropping neighbours."""
    import torch
    r = np.random.default_rng(0)
    n, m, d, nq = 20_000, 8, 32, 4
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq, d)).astype(np.float32)
    cb = r.standard_normal((m * 256, d)).astype(np.float32)
    nrm = r.standard_normal(n).astype(np.float32)
    ix = rb.core.Index(rb.core.SCAN_LSQ, B, nrm)
    for bad in (np.inf, np.nan, 3e38):
        cb2 = cb.copy()
        cb2[300, 5] = bad
        with pytest.raises(rb.RayuelaError, match="lookup-table"):
            ix.search(Xq, cb2, 3)
        dd, ii = ix.search(torch.from_numpy(Xq).cuda(), torch.from_numpy(cb2).cuda(), 3)
        assert bool(torch.isnan(dd).all()) and bool((ii == -1).all())
    dg, ig = ix.search(Xq, cb, 3)                                   # the handle is still usable
    d0, i0 = orc.linscan(orc.LSQ, B, Xq, cb, 3, nrm)
    assert np.array_equal(ig, i0)


def test_encode_chunk_pipeline_matches_single_chunk(rb, monkeypatch):
    """Host arrays are encoded in >= 2 chunks on alternating streams with the upload of chunk c+1 overlapping the
    kernels of chunk c; forced here with a small unary budget.  Same bits as one chunk, host and device arrays."""
    import torch
    X, C, B = _icm_data(9000, 32, 8, seed=4)
    want = orc.encode_icm(X, C, B, 2, 4, 4, True, seed=1, snap_iters=[2])
    monkeypatch.setenv("RAYUELA_B200_UNARY_BYTES", str(8 * 1024 * 2048))     # 2048 vectors per chunk -> 5 chunks
    for conv in (lambda a: a, lambda a: torch.from_numpy(a).cuda()):
        r = rb.core.encode_icm(conv(X), conv(C), conv(B), 2, 4, 4, True, seed=1, snap_iters=[2], want_cost=True,
                               want_stats=True)
        Bg = r["B"].cpu().numpy() if hasattr(r["B"], "cpu") else r["B"]
        cg = r["cost"].cpu().numpy() if hasattr(r["cost"], "cpu") else r["cost"]
        sg = r["B_snap"].cpu().numpy() if hasattr(r["B_snap"], "cpu") else r["B_snap"]
        assert np.array_equal(Bg, want["B"])
        assert np.array_equal(bits(cg), bits(want["cost"]))
        assert np.array_equal(sg, want["B_snap"])
        assert np.array_equal(r["stats"], want["stats"])
