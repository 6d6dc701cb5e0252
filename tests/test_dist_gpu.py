"""GPU, world_size = 2 over NCCL: rayuela_b200.dist with the real CUDA backend (skipped with fewer than 2 GPUs).
Sharded encode == single-GPU encode bit for bit; base-sharded scan + all-gather + merge == single-GPU scan."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    r = np.random.default_rng(0)
    n, d, m, nq = 20001, 32, 8, 33
    X = r.standard_normal((n, d)).astype(np.float32)
    C = (r.standard_normal((m * 256, d)) / 3).astype(np.float32)
    B = r.integers(0, 256, (n, m), dtype=np.uint8)
    Xq = np.round(r.standard_normal((nq, d)) * 2).astype(np.float32)
    return X, C, B, Xq


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as tdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from rayuela_b200 import core, dist
        X, C, B, Xq = _data()
        Xd, Cd, Bd = (torch.from_numpy(a).cuda() for a in (X, C, B))
        full, _ = dist.sharded_encode_icm(Xd, Cd, Bd, 3, 4, 4, True, seed=5, gather=True)
        codes = full.cpu().numpy()
        nrm = np.round(np.random.default_rng(1).standard_normal(len(codes)) * 2).astype(np.float32)
        a, b = dist.splitarray(len(codes), world)[rank]
        six = dist.ShardedIndex(core.SCAN_LSQ, full[a:b].contiguous(), torch.from_numpy(nrm[a:b]).cuda(), a)
        Cq = torch.round(Cd * 2)
        d_b, i_b = six.search(torch.from_numpy(Xq).cuda(), Cq, 50)
        whole = core.Index(core.SCAN_LSQ, full, torch.from_numpy(nrm).cuda())
        d_q, i_q = dist.query_sharded_search(whole, torch.from_numpy(Xq).cuda(), Cq, 50)
        out[rank] = dict(codes=codes, d_b=d_b.cpu().numpy(), i_b=i_b.cpu().numpy(), d_q=d_q.cpu().numpy(),
                         i_q=i_q.cpu().numpy(), nrm=nrm)
    finally:
        tdist.destroy_process_group()


def test_world2_nccl_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    import rayuela_b200 as rb
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    X, C, B, Xq = _data()
    single = rb.core.encode_icm(X, C, B, 3, 4, 4, True, seed=5)["B"]
    assert np.array_equal(single, orc.encode_icm(X, C, B, 3, 4, 4, True, seed=5)["B"])
    for rank in range(2):
        r = out[rank]
        assert np.array_equal(r["codes"], single)
        d0, i0 = orc.linscan(orc.LSQ, single, Xq, np.round(C * 2), 50, r["nrm"])
        assert np.array_equal(r["i_b"], i0) and np.array_equal(r["d_b"].view(np.uint32), d0.view(np.uint32))
        assert np.array_equal(r["i_q"], i0) and np.array_equal(r["d_q"].view(np.uint32), d0.view(np.uint32))
