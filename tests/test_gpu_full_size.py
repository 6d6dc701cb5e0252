"""GPU parity at BASELINE.json's full sizes (1M x 128 base, 10k queries, 100M-code index) through properties that do
not need a full CPU run: contiguous sub-ranges against the oracle / the reference's compiled C++ (vectors and
queries are independent, the RNG is keyed on the global index), shard invariance, sortedness, never-worsening
costs, pre-filter on/off equality, k = 1 being the head of k = 100."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rayuela_b200
    return rayuela_b200, torch


@pytest.fixture(scope="module")
def base(env):
    rb, torch = env
    g = torch.Generator(device="cuda").manual_seed(7)
    n, d, m = 1_000_000, 128, 8
    X = torch.randn(n, d, device="cuda", generator=g)
    C = torch.randn(m * 256, d, device="cuda", generator=g) / m ** 0.5
    B0 = torch.randint(0, 256, (n, m), device="cuda", dtype=torch.uint8, generator=g)
    return X, C, B0


def test_icm_1M_subranges_match_oracle_and_shards_agree(env, base, monkeypatch):
    rb, torch = env
    X, C, B0 = base
    n, m = B0.shape
    ils, icm, npert = 8, 4, 4                                  # configs[2] with the train-time ilsiter (LSQ_GPU.jl:333)
    full = rb.core.encode_icm(X, C, B0, ils, icm, npert, True, seed=77, want_cost=True, want_stats=True)
    ex, tot = rb.core.last_icm_steps()
    assert tot == n * ils * icm * m and 0 < ex < tot
    Bf, cf = full["B"], full["cost"]
    # (1) contiguous sub-ranges against the oracle (CPU restatement + the reference's `condition` when built)
    Xh, Ch, B0h = X.cpu().numpy(), C.cpu().numpy(), B0.cpu().numpy()
    for s in (0, 499_321, n - 1500):
        want = orc.encode_icm(Xh[s:s + 1500], Ch, B0h[s:s + 1500], ils, icm, npert, True, seed=77, g0=s,
                              use_ref_step=orc.have_ref())
        assert np.array_equal(Bf[s:s + 1500].cpu().numpy(), want["B"]), s
        assert np.array_equal(bits(cf[s:s + 1500].cpu().numpy()), bits(want["cost"])), s
    # (2) shard invariance: two ranks' slices with their global offsets give the same codes as one call
    h = n // 2 + 12345
    a = rb.core.encode_icm(X[:h], C, B0[:h], ils, icm, npert, True, seed=77, g0=0)["B"]
    b = rb.core.encode_icm(X[h:], C, B0[h:], ils, icm, npert, True, seed=77, g0=h)["B"]
    assert torch.equal(torch.cat([a, b]), Bf)
    # (3) the ILS accept rule never worsens a vector, and cost_out is veccost of the returned codes
    c0 = rb.core.veccost(X, B0, C)
    c1 = rb.core.veccost(X, Bf, C)
    assert bool((cf <= c0).all()) and torch.equal(c1.view(torch.int32), cf.view(torch.int32))
    # (4) the quantised pre-filter does not change a single code at full size
    monkeypatch.setenv("RAYUELA_B200_ICM_PF", "0")
    plain = rb.core.encode_icm(X, C, B0, ils, icm, npert, True, seed=77)["B"]
    assert torch.equal(plain, Bf)
    # (5) zero ILS iterations return the input codes
    assert torch.equal(rb.core.encode_icm(X, C, B0, 0, icm, npert, True, seed=77)["B"], B0)


def test_icm_1M_timed_configuration_ilsiter32(env, base):
    """The TIMED configuration (configs[2]: 1M x 128, m = 8, ilsiter = 32 as src/LSQ_GPU.jl:352 uses for the base set):
    sub-ranges against the oracle, and the three ways of running it -- device arrays (one chunk), host arrays (chunk
    pipeline on alternating streams with overlapped uploads), two shards inside one process -- give the same bits."""
    rb, torch = env
    X, C, B0 = base
    n, m = B0.shape
    ils, icm, npert = 32, 4, 4
    dev_run = rb.core.encode_icm(X, C, B0, ils, icm, npert, True, seed=2024)["B"].cpu().numpy()
    Xh, Ch, B0h = X.cpu().numpy(), C.cpu().numpy(), B0.cpu().numpy()
    for s in (123_456, n - 1024):
        want = orc.encode_icm(Xh[s:s + 1024], Ch, B0h[s:s + 1024], ils, icm, npert, True, seed=2024, g0=s,
                              use_ref_step=orc.have_ref())
        assert np.array_equal(dev_run[s:s + 1024], want["B"]), s
    host_run = rb.core.encode_icm(Xh, Ch, B0h, ils, icm, npert, True, seed=2024)["B"]
    assert np.array_equal(host_run, dev_run)
    rb.init([0, 0])
    try:
        two = rb.core.encode_icm(Xh, Ch, B0h, ils, icm, npert, True, seed=2024)["B"]
    finally:
        rb.init(None)
    assert np.array_equal(two, dev_run)


@pytest.mark.parametrize("m", [8, 16])
def test_scan_10k_by_1M_properties(env, m):
    rb, torch = env
    g = torch.Generator(device="cuda").manual_seed(m)
    n, nq, d, k = 1_000_000, 10_000, 128, 100
    B = torch.randint(0, 256, (n, m), device="cuda", dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device="cuda", generator=g) * 3
    Q = torch.randn(nq, d, device="cuda", generator=g)
    C = torch.randn(m * 256, d, device="cuda", generator=g)
    ix = rb.core.Index(rb.core.SCAN_LSQ, B, nrm)
    dd, ii = ix.search(Q, C, k)
    # sorted by (dist, id); ids unique per query and in range (1-based, pairwise_byte.cpp:76)
    assert bool((dd[:, 1:] >= dd[:, :-1]).all())
    tie = dd[:, 1:] == dd[:, :-1]
    assert bool((ii[:, 1:][tie] > ii[:, :-1][tie]).all())
    assert int(ii.min()) >= 1 and int(ii.max()) <= n
    assert bool((torch.sort(ii, dim=1).values[:, 1:] != torch.sort(ii, dim=1).values[:, :-1]).all())
    # k = 1 is the head of k = 100
    d1, i1 = ix.search(Q, C, 1)
    assert torch.equal(i1[:, 0], ii[:, 0]) and torch.equal(d1[:, 0].view(torch.int32), dd[:, 0].view(torch.int32))
    # a query subset against the reference's own C++ over the full base
    sel = torch.tensor([0, 17, 4095, 4096, 9999, 5000, 1234, 8191], device="cuda")
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    d0, i0 = fn(orc.LSQ, B.cpu().numpy(), Q[sel].cpu().numpy(), C.cpu().numpy(), k, nrm.cpu().numpy())
    assert np.array_equal(ii[sel].cpu().numpy(), i0)
    assert np.array_equal(bits(dd[sel].cpu().numpy()), bits(d0))
    # base-sharded search + merge (config-5 structure) equals the single index
    cut = 612_345
    parts = [rb.core.Index(rb.core.SCAN_LSQ, B[:cut], nrm[:cut], id_offset=0).search(Q, C, k),
             rb.core.Index(rb.core.SCAN_LSQ, B[cut:], nrm[cut:], id_offset=cut).search(Q, C, k)]
    dm, im = rb.core.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(im, ii) and torch.equal(dm.view(torch.int32), dd.view(torch.int32))


def test_scan_config5_100M_subset_vs_reference(env):
    """BASELINE configs[4] on one GPU: 100M x 8-byte codes + fp32 norms; 16 of the queries against the reference's C++."""
    rb, torch = env
    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs ~40 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    n, nq, d, m, k = 100_000_000, 512, 128, 8, 50
    B = torch.randint(0, 256, (n, m), device="cuda", dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device="cuda", generator=g) * 3
    Q = torch.randn(nq, d, device="cuda", generator=g)
    C = torch.randn(m * 256, d, device="cuda", generator=g)
    ix = rb.core.Index(rb.core.SCAN_LSQ, B, nrm)
    dd, ii = ix.search(Q, C, k)
    assert bool((dd[:, 1:] >= dd[:, :-1]).all()) and int(ii.max()) <= n
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    d0, i0 = fn(orc.LSQ, B.cpu().numpy(), Q[:16].cpu().numpy(), C.cpu().numpy(), k, nrm.cpu().numpy())
    assert np.array_equal(ii[:16].cpu().numpy(), i0)
    assert np.array_equal(bits(dd[:16].cpu().numpy()), bits(d0))
