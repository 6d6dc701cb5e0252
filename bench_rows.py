#!/usr/bin/env python
"""bench_rows.py -- secondary measurements: the non-headline rows of SURVEY section 8 (PQ/OPQ encode, veccost,
codebook update, norm quantization, ChainQ Viterbi), each on the GPU with device-resident inputs (CUDA events)
and on the host through the oracle / reference C++ on a bounded sample.  Prints one JSON object.
    python bench_rows.py [--n 1000000] [--out profiles/xxx.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "rayuela.jl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def gpu_time(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def cpu_time(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import bench
    from oracle import oracle as orc
    from rayuela_b200 import core
    orc.build()
    dev = torch.device("cuda")
    n, d, m = args.n, 128, 8
    cores, _ = bench.force_host_threads()          # OpenMP / BLAS pools at the host core count, verified
    X, _ = bench.make_data(n, 16, d, 1000, dev)
    C = bench.train_codebooks(X[:50000], m, dev)
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
    core.encode_icm(X, C, B, 2, 4, 4, True, seed=1, inplace=True)
    Xh, Ch, Bh = X.cpu().numpy(), C.cpu().numpy(), B.cpu().numpy()
    rows = {}

    def row(name, unit_count, gpu_fn, cpu_fn, cpu_count, note):
        tg = gpu_time(gpu_fn)
        tc = cpu_time(cpu_fn)
        rows[name] = {"gpu_s": tg, "gpu_per_s": unit_count / tg, "cpu_per_s": cpu_count / tc, "cpu_sample": cpu_count,
                      "cpu_s": tc, "speedup": (unit_count / tg) / (cpu_count / tc), "note": note}

    ns = min(n, 100_000)
    row("veccost (vectors/s)", n, lambda: core.veccost(X, B, C), lambda: orc.veccost(Xh[:ns], Bh[:ns], Ch), ns,
        "src/qerrors.jl:36-66")
    row("fast_bin_matmul (vectors/s)", n, lambda: core.fast_bin_matmul(X, B),
        lambda: orc.fast_bin_matmul(Xh[:ns], Bh[:ns]), ns, "src/codebook_update.jl:96-171; oracle is single-threaded "
        "like the reference loop")
    row("update_codebooks_fast_bin incl. LU solve (vectors/s)", n, lambda: core.update_codebooks_fast_bin(X, B),
        lambda: orc.update_codebooks_fast_bin(Xh[:ns], Bh[:ns]), ns, "solve: torch.linalg (GPU) vs scipy LAPACK (host)")
    cb = torch.sort(torch.rand(256, device=dev) * 60)[0].contiguous()
    row("quantize_norms (vectors/s)", n, lambda: core.quantize_norms(B, C, cb),
        lambda: orc.quantize_norms(Bh[:ns], Ch, cb.cpu().numpy()), ns, "src/utils.jl:29-59")
    nv = min(n, 200_000)
    Xv = X[:nv].contiguous()
    row("quantize_chainq Viterbi (vectors/s)", nv, lambda: core.quantize_chainq(Xv, C, m),
        lambda: orc.quantize_chainq(Xh[:5000], Ch, m, use_ref=orc.have_ref()), 5000,
        "src/ChainQ.jl:287-348; CPU = reference viterbi_encoding C++ + restated unaries")
    Cpq = torch.randn(m * 256, d // m, device=dev)
    row("quantize_pq (vectors/s)", n, lambda: core.quantize_pq(X, Cpq, m),
        lambda: orc.quantize_pq(Xh[:ns], Cpq.cpu().numpy(), m), ns, "src/PQ.jl:18-48")
    out = {"n": n, "d": d, "m": m, "cores": cores, "gpu": torch.cuda.get_device_name(0), "rows": rows}
    txt = json.dumps(out, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    main()
