"""Generates tests/golden/*.npz -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference for oracle/_ref):   python -m oracle.gen_golden
Every scan fixture's expected output comes from the REFERENCE'S OWN compiled C++ (oracle/_ref, built
unmodified from deps/src/linscan_aqd*.cpp); every ICM fixture's expected output comes from the oracle
restatement driven through the reference's own compiled `condition` (deps/src/encode_icm.cpp).
Inputs follow test/common.jl:2-8 (X = rand*10, C = rand, B = rand(1:h)) or N(0,1) data.
"""
import os

import numpy as np

from oracle import oracle as orc

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def dataset(kind, n, d, m, seed, h=256):
    r = np.random.default_rng(seed)
    if kind == "uniform":   # test/common.jl:2-8
        X = (r.random((n, d)) * 10).astype(np.float32)
        C = r.random((m * h, d)).astype(np.float32)
    else:
        X = r.standard_normal((n, d)).astype(np.float32)
        C = (r.standard_normal((m * h, d)) / np.sqrt(m)).astype(np.float32)
    B = r.integers(0, h, (n, m), dtype=np.uint8)
    return X, C, B


def main():
    os.makedirs(OUT, exist_ok=True)
    orc.build()
    assert orc.have_ref(), "oracle/_ref missing: golden vectors must come from the reference's own code"

    for name, kind, n, d, m, ils, icm, npert, randord, seed in [
        ("icm_m8_gauss", "gauss", 384, 32, 8, 3, 4, 4, True, 7),
        ("icm_m7_uniform", "uniform", 256, 16, 7, 2, 2, 4, False, 11),
        ("icm_m16_gauss", "gauss", 192, 24, 16, 2, 3, 5, True, 13),
    ]:
        X, C, B = dataset(kind, n, d, m, seed)
        r = orc.encode_icm(X, C, B, ils, icm, npert, randord, seed=seed, g0=1000, snap_iters=[1, ils],
                           use_ref_step=True)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, C=C, B=B, ilsiter=ils, icmiter=icm, npert=npert,
                            randord=randord, seed=seed, g0=1000, B_out=r["B"], cost=r["cost"], stats=r["stats"],
                            B_snap=r["B_snap"], objs=r["objs"], snap_iters=np.array([1, ils]))

    for name, kind, m, d, n, nq, k, ties in [
        ("scan_lsq_m8", orc.LSQ, 8, 32, 6000, 6, 40, False),
        ("scan_lsq_m7_ties", orc.LSQ, 7, 16, 5000, 5, 64, True),
        ("scan_cq_m8", orc.CQ, 8, 32, 4000, 4, 16, False),
        ("scan_pq_m8", orc.PQ, 8, 32, 6000, 6, 40, False),
        ("scan_pq_m16_ties", orc.PQ, 16, 32, 3000, 4, 100, True),
    ]:
        r = np.random.default_rng(k + m)
        B = r.integers(0, 256, (n, m), dtype=np.uint8)
        Xq = r.standard_normal((nq, d)).astype(np.float32)
        cols = d // m if kind == orc.PQ else d
        cb = r.standard_normal((m * 256, cols)).astype(np.float32)
        if ties:   # small-integer data: many exactly equal distances, order decided by the id
            Xq, cb = np.round(Xq * 2), np.round(cb * 2)
            B = r.integers(0, 3, (n, m), dtype=np.uint8)
        nrm = (np.round(r.standard_normal(n) * 3) if ties else r.standard_normal(n) * 3).astype(np.float32) \
            if kind == orc.LSQ else None
        dists, idx = orc.ref_linscan(kind, B, Xq, cb, k, nrm)
        kw = dict(B=B, Xq=Xq, cb=cb, k=k, kind=kind, dists=dists, idx=idx)
        if nrm is not None:
            kw["nrm"] = nrm
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)

    n, m, sub = 512, 8, 4
    r = np.random.default_rng(5)
    X = r.standard_normal((n, m * sub)).astype(np.float32)
    Cpq = r.standard_normal((m * 256, sub)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "pq_encode_m8.npz"), X=X, Cpq=Cpq, m=m, B_out=orc.quantize_pq(X, Cpq, m))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
