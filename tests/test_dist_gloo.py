"""CPU, world_size = 2 over gloo: the partition / exchange logic of rayuela_b200.dist.
The compute primitives are replaced by an oracle-backed stand-in (TEST ONLY -- the product backend is CUDA and
has no CPU path); what is under test is splitarray, global-index RNG keying, id offsets, the single all_gather
and the merge order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import oracle as orc


class OracleBackend:
    def encode(self, X, C, B, ilsiter, icmiter, npert, randord, seed, g0):
        return orc.encode_icm(np.asarray(X), np.asarray(C), np.asarray(B), ilsiter, icmiter, npert, randord,
                              seed=seed, g0=g0)["B"]

    def make_index(self, kind, codes, norms, id_offset):
        class _Ix:
            def search(_s, q, cb, k):
                return orc.linscan(kind, codes, np.asarray(q), np.asarray(cb), k, norms, id_offset=id_offset)
        return _Ix()

    def merge(self, dists, idx):
        d, i = dists.numpy(), idx.numpy()
        S, nq, k = d.shape
        do, io = np.empty((nq, k), np.float32), np.empty((nq, k), np.int32)
        for q in range(nq):
            dd, ii = d[:, q].reshape(-1), i[:, q].reshape(-1)
            o = np.lexsort((ii, dd))[:k]
            do[q], io[q] = dd[o], ii[o]
        return torch.from_numpy(do), torch.from_numpy(io)

    def to_tensor(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a


def _data():
    r = np.random.default_rng(0)
    n, d, m, nq = 1001, 16, 4, 9
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 2).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = np.round(r.standard_normal((nq, d)) * 2).astype(np.float32)
    return X, C, B, Xq


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rayuela_b200 import dist
        X, C, B, Xq = _data()
        be = OracleBackend()
        full, span = dist.sharded_encode_icm(X, C, B, 2, 2, 4, True, seed=3, gather=True, backend=be)
        a, b = dist.splitarray(X.shape[0], world)[rank]
        loc, span2 = dist.sharded_encode_icm(X[a:b], C, B[a:b], 2, 2, 4, True, seed=3, backend=be,
                                             local_slice=(a, b, X.shape[0]))
        codes = full.numpy()
        nrm = np.round(np.random.default_rng(1).standard_normal(codes.shape[0]) * 2).astype(np.float32)
        six = dist.ShardedIndex(orc.LSQ, codes[a:b], nrm[a:b], a, backend=be)
        d_b, i_b = six.search(Xq, np.round(C * 2), 20)
        class _Whole:
            def search(_s, q, cb, k):
                return orc.linscan(orc.LSQ, codes, q, cb, k, nrm)
        d_q, i_q = dist.query_sharded_search(_Whole(), Xq, np.round(C * 2), 20, backend=be)
        out[rank] = dict(full=codes, span=span, loc=np.asarray(loc), span2=span2, d_b=d_b.numpy(), i_b=i_b.numpy(),
                         d_q=d_q.numpy(), i_q=i_q.numpy(), nrm=nrm)
    finally:
        tdist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_splitarray_rule():
    from rayuela_b200 import dist
    assert dist.splitarray(10, 3) == [(0, 4), (4, 7), (7, 10)]      # src/utils.jl:179-203
    assert dist.splitarray(8, 8) == [(i, i + 1) for i in range(8)]
    assert dist.splitarray(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]


@pytest.mark.timeout(300)
def test_world2_sharding_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    X, C, B, Xq = _data()
    single = orc.encode_icm(X, C, B, 2, 2, 4, True, seed=3)["B"]
    from rayuela_b200 import dist
    for rank in range(world):
        r = out[rank]
        a, b = dist.splitarray(X.shape[0], world)[rank]
        assert np.array_equal(r["full"], single)                    # sharded encode == single-process encode
        assert np.array_equal(r["loc"], single[a:b]) and tuple(r["span2"]) == (a, b)
        d0, i0 = orc.linscan(orc.LSQ, single, Xq, np.round(C * 2), 20, r["nrm"])
        assert np.array_equal(r["i_b"], i0) and np.array_equal(r["d_b"], d0)   # base-sharded scan + merge
        assert np.array_equal(r["i_q"], i0) and np.array_equal(r["d_q"], d0)   # query-sharded scan
