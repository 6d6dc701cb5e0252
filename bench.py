#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path icm|linscan]

Default (--path icm): BASELINE.json configs[2] -- LSQ++ (SR-D) m=8 h=256 icmiter=4 npert=4 randord, 1M x 128
synthetic base encoded from random initial codes with ilsiter=32 (the reference's base-set setting,
src/LSQ_GPU.jl:351-352).  One step = one rayuela_encode_icm call over the whole base shard.
`value` = encoded vectors/s with inputs resident in HBM; `e2e` = the same through the C ABI with pinned HOST
buffers (H2D of X/C/B and D2H of the codes inside the timed region).  The JSON line also carries a `linscan`
object: queries/s and Recall@1 of linscan_lsq (10k queries over the just-encoded base) with its own roofline.
--impl reference times the reference's CPU algorithm (oracle restatement of the Julia host logic driving the
reference's own compiled `condition` / linscan symbols from oracle/_ref; unaries and pairwise tables by BLAS sgemm
as src/utils.jl:135-136,164 compute them) on a bounded sample per step, with the OpenMP / BLAS pools forced to
the host core count (torchrun exports OMP_NUM_THREADS=1) and the effective thread count printed.
Under torchrun (N > 1) every rank works on its own shard (weak scaling, no data-path collective for the
encode; one all-gather + merge for the base-sharded scan).  The line also carries two STRONG-scaling objects,
`config4` (BASELINE.json configs[3]: m=16, 1M vectors in total split over the N ranks) and `config5`
(configs[4]: 100M codes in total, base-sharded, k in {1, 1000}); --no-strong skips them.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# The CPU arm (--impl reference, and the cpu_baseline leg of a 1-GPU run) must use every host core.  torchrun exports
# OMP_NUM_THREADS=1 to its workers; libgomp and OpenBLAS read the variable when they are LOADED, so it is corrected
# here, before numpy / the oracle are imported (force_host_threads() then verifies what the runtimes report).
if "reference" in sys.argv[1:] or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(host_threads())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "rayuela.jl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

H = 256


def ncu_traffic(kind):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the `ncu --set full`
    capture of the SAME workload committed under profiles/ (bench.py cannot run ncu): profiles/ncu_traffic.json
    maps kind -> {"bytes": ..., "kernel": ..., "capture": file, "commit": ...}; null when there is no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(kind)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------
# synthetic workload (harness code: torch on the GPU, never timed)
# ---------------------------------------------------------------------------------------------------------
def make_data(n, nq, d, seed, device, rank=0):
    """SIFT-like synthetic base: 1024 Gaussian clusters with a power-law spectrum (sigma_i ~ i^-0.75) plus
    unit within-cluster noise on the same spectrum, then a fixed random rotation; queries come from the same
    distribution.  Isotropic noise would make Recall@1 degenerate (~1%); this gives R@1 ~ 0.3 at 64-bit codes,
    the regime of the reference's SIFT1M demo (probe: tools/recall_probe.py)."""
    import torch
    g = torch.Generator(device=device).manual_seed(1234)
    s = torch.arange(1, d + 1, device=device, dtype=torch.float32) ** -0.75
    rot, _ = torch.linalg.qr(torch.randn(d, d, generator=g, device=device))
    centres = torch.randn(1024, d, generator=g, device=device) * s

    def draw(k, gen):
        z = centres[torch.randint(0, 1024, (k,), generator=gen, device=device)] + \
            torch.randn(k, d, generator=gen, device=device) * s
        return (z @ rot).contiguous()
    X = draw(n, torch.Generator(device=device).manual_seed(seed + 7919 * rank))
    Q = draw(nq, torch.Generator(device=device).manual_seed(4321))
    return X, Q


def kmeans(x, k, iters, gen):
    import torch
    c = x[torch.randperm(x.shape[0], generator=gen, device=x.device)[:k]].clone()
    for _ in range(iters):
        d2 = (c * c).sum(1)[None, :] - 2 * x @ c.T
        a = d2.argmin(1)
        # segment sums as a GEMM (one-hot^T @ x): run-to-run deterministic, unlike index_add_'s atomics -- the strong-
        # scaling objects compare code checksums across N, which needs bit-identical codebooks on every rank and run
        s = torch.nn.functional.one_hot(a, k).to(x.dtype).T @ x
        cnt = torch.bincount(a, minlength=k).clamp(min=1)[:, None]
        c = s / cnt
    d2 = (c * c).sum(1)[None, :] - 2 * x @ c.T
    return c, d2.argmin(1)


def train_codebooks(Xt, m, device):
    """Harness stand-in for the (out-of-scope) trainer: residual k-means init, then SR-D noise as
    SR_D_perturb does at iter 1 of 25, schedule 1, p = 0.5 (src/SR_perturbations.jl:27-49)."""
    import torch
    gen = torch.Generator(device=device).manual_seed(99)
    res = Xt.clone()
    Cs = []
    for _ in range(m):
        c, a = kmeans(res, H, 8, gen)
        Cs.append(c)
        res -= c[a]
    C = torch.stack(Cs)                                     # [m][256][d]
    std = C.reshape(m * H, -1).std(0, unbiased=True) / m
    std = std * (1 - 1 / 25) ** 0.5
    C = C + torch.randn(C.shape, generator=gen, device=device) * std
    return C.reshape(m * H, -1).contiguous()


def exact_nn(X, Q):
    import torch
    best = torch.full((Q.shape[0],), float("inf"), device=X.device)
    arg = torch.zeros(Q.shape[0], dtype=torch.long, device=X.device)
    qn = (Q * Q).sum(1)
    for s in range(0, X.shape[0], 1 << 18):
        xb = X[s:s + (1 << 18)]
        d2 = qn[:, None] + (xb * xb).sum(1)[None, :] - 2 * Q @ xb.T
        v, i = d2.min(1)
        upd = v < best
        best = torch.where(upd, v, best)
        arg = torch.where(upd, i + s, arg)
    return arg


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_sized(fn, first, target_s=12.0, cap=None):
    """Run fn(size) once at `first`; if that took well under the target, run again at a size scaled to
    ~target_s (bounded by cap) and return the larger run."""
    t, kind, extra = fn(first)
    size = first
    if t < target_s / 3:
        size = int(min(cap or 10 ** 9, max(first, first * target_s / max(t, 1e-3))))
        if size > first:
            t, kind, extra = fn(size)
    return size, t, kind, extra


def force_host_threads():
    """All host cores for the CPU arm, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1, and an
    environment variable changed after libgomp / OpenBLAS are loaded is ignored): runtime calls, then read back."""
    from oracle import oracle as orc
    orc.build()
    cores = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    got = orc.set_num_threads(cores)
    assert got == cores, "OpenMP gives %d threads, wanted %d" % (got, cores)
    return cores, got


def cpu_icm_sample(cfg, n_s, seed=0):
    """The reference's CPU encode on n_s vectors: oracle.encode_icm (encode_icm_fully! restated,
    src/LSQ.jl:152-252) driving the reference's own compiled `condition` when oracle/_ref is present, with
    unaries / tables from BLAS sgemm like src/utils.jl:135-136,164 (blas=True: the oracle's pinned fixed-order
    chains are for parity, not for timing).  Returns (seconds, kind, phases)."""
    from oracle import oracle as orc
    r = np.random.default_rng(seed)
    X = r.standard_normal((n_s, cfg["d"])).astype(np.float32)
    C = (r.standard_normal((cfg["m"] * H, cfg["d"])) / np.sqrt(cfg["m"])).astype(np.float32)
    B = r.integers(0, H, (n_s, cfg["m"]), dtype=np.uint8)
    t0 = time.perf_counter()
    o = orc.encode_icm(X, C, B, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=1,
                       use_ref_step=orc.have_ref(), blas=True, want_phases=True)
    return time.perf_counter() - t0, ("reference" if orc.have_ref() else "port"), o["phases"]


def cpu_scan_sample(cfg, nq_s, n, seed=0):
    from oracle import oracle as orc
    r = np.random.default_rng(seed)
    m = cfg["m"]
    B = r.integers(0, H, (n, m), dtype=np.uint8)
    Xq = r.standard_normal((nq_s, cfg["d"])).astype(np.float32)
    cb = r.standard_normal((m * H, cfg["d"])).astype(np.float32)
    nrm = r.standard_normal(n).astype(np.float32)
    fn = orc.ref_linscan if orc.have_ref() else orc.linscan
    t0 = time.perf_counter()
    fn(orc.LSQ, B, Xq, cb, cfg["k"], nrm)
    return time.perf_counter() - t0, ("reference" if orc.have_ref() else "port"), None


CPU_NOTE = ("reference C++ (oracle/_ref: `condition`, linscan_aqd*) + restated Julia host logic (Julia is not "
            "installed); unaries and pairwise tables by OpenBLAS sgemm as the reference computes them")


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, omp = force_host_threads()
    times, phases = [], None
    if args.path == "icm":
        n_s = args.ref_sample or 32768
        for i in range(args.warmup + args.steps):
            t, kind, ph = cpu_icm_sample(cfg, n_s, seed=i)
            if i >= args.warmup:
                times.append(t)
                phases = ph
        per = float(np.mean(times))
        value, unit, metric = n_s / per, "vectors/s", "lsq_icm_encode_vectors_per_sec"
        sample = "%d of %d vectors per step, same m/ilsiter/icmiter/npert; vectors/s is size-independent" % (
            n_s, cfg["n"])
    else:
        nq_s = args.ref_sample or 256 * cores
        for i in range(args.warmup + args.steps):
            t, kind, _ = cpu_scan_sample(cfg, nq_s, cfg["n"], seed=i)
            if i >= args.warmup:
                times.append(t)
        per = float(np.mean(times))
        value, unit, metric = nq_s / per, "queries/s", "linscan_lsq_queries_per_sec"
        sample = "%d of %d queries per step over the full %d-code base" % (nq_s, cfg["nq"], cfg["n"])
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "omp_max_threads": omp, "kind": kind,
                         "sample": sample, "phases_s": phases, "note": CPU_NOTE},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfg, args):
    return {"workload": "BASELINE.json configs[2]: LSQ++ (SR-D) m=%d h=256 icmiter=%d npert=%d randord ilsiter=%d, "
                        "%dx%d synthetic base per GPU, random initial codes; linscan_lsq %d queries k=%d over the "
                        "encoded base" % (cfg["m"], cfg["icmiter"], cfg["npert"], cfg["ilsiter"], cfg["n"], cfg["d"],
                                          cfg["nq"], cfg["k"]),
            "path": args.path, "n_per_gpu": cfg["n"], "d": cfg["d"], "m": cfg["m"], "h": H,
            "ilsiter": cfg["ilsiter"], "icmiter": cfg["icmiter"], "npert": cfg["npert"], "nq": cfg["nq"],
            "k": cfg["k"], "l2": "inputs larger than L2 (X 512 MB + unaries 8 GB per step; codes+norms 12 MB are "
                                 "re-read per query tile by design, LUTs 82 MB)"}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, warmup, dist, device):
    """warmup, then exactly `steps` calls between barrier+synchronize; device time by CUDA events on the
    launching (current) stream; max over ranks."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import rayuela_b200 as _rb
    n0 = _rb.launch_count()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    timed_steps.last_launches = _rb.launch_count() - n0       # this library's kernels inside the timed region only
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def wall_steps(fn, steps, warmup, dist, device):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(device)
    ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def pinned(a):
    import torch
    t = torch.empty(a.shape, dtype=a.dtype, pin_memory=True)
    t.copy_(a)
    return t.numpy()


def run_ours(args, cfg):
    import torch
    import rayuela_b200 as rb
    from rayuela_b200 import core
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    pk, pk_src = peaks()
    n, d, m, nq, k = cfg["n"], cfg["d"], cfg["m"], cfg["nq"], cfg["k"]

    # ---- workload -----------------------------------------------------------------------------------------
    real = None
    if args.data_dir:
        # real SIFT1M (src/read_datasets.jl:63-85; demos/experiment_utils.jl:63-86) when the files are there: base,
        # queries and -- for the full base on one GPU -- the dataset's own ground truth, so Recall@1 is the true one
        from rayuela_b200 import demos
        if demos.have_sift1m(args.data_dir):
            a0 = rank * n
            Xb = demos.read_dataset("SIFT1M_base", (a0 + 1, a0 + n), False, args.data_dir)
            Xq = demos.read_dataset("SIFT1M_query", nq, False, args.data_dir)
            gtf = demos.read_dataset("SIFT1M_groundtruth", nq, False, args.data_dir)
            X = torch.from_numpy(np.ascontiguousarray(Xb.T)).to(device)
            Q = torch.from_numpy(np.ascontiguousarray(Xq.T)).to(device)
            real = {"dataset": "SIFT1M", "dir": args.data_dir,
                    "gt": torch.from_numpy(gtf[0, :nq].astype(np.int64)).to(device) if world == 1 and n == 1_000_000
                    else None}
            d = X.shape[1]
        else:
            print("bench: no SIFT1M files under %s (sift/sift_{learn,base,query}.fvecs, sift_groundtruth.ivecs); "
                  "using the synthetic base" % args.data_dir, file=sys.stderr)
    if real is None:
        X, Q = make_data(n, nq, d, seed=1000, device=device, rank=rank)
    C = train_codebooks(X[:50000], m, device)
    gB = torch.Generator(device=device).manual_seed(5 + rank)
    B0 = torch.randint(0, H, (n, m), generator=gB, device=device, dtype=torch.uint8)   # src/LSQ_GPU.jl:351
    g0 = rank * n
    Bwork = B0.clone()

    def icm_step():
        Bwork.copy_(B0)
        core.encode_icm(X, C, Bwork, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0,
                        inplace=True)

    out = {}
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = rb.launch_count()
    # ---- path (1): ICM encode -------------------------------------------------------------------------------
    icm_ms = timed_steps(icm_step, args.steps, args.warmup, dist, device)
    icm_launches = timed_steps.last_launches
    icm_per = icm_ms / args.steps
    icm_value = world * n / (icm_per * 1e-3)
    # dominant kernel alone (K3), timed by events on the launching stream: encode minus unary/table kernels is
    # not separable from outside, so time a second pass with ilsiter=0 (K0+K1+K2 + one cost pass) and subtract
    Bscratch = B0.clone()

    def icm_setup_only():
        core.encode_icm(X, C, Bscratch, 0, cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0, inplace=True)
    setup_ms = timed_steps(icm_setup_only, max(2, args.steps // 2), 1, dist, device) / max(2, args.steps // 2)
    k3_ms = max(icm_per - setup_ms, 1e-6)
    core.encode_icm(X, C, B0.clone(), cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0,
                    inplace=True, want_stats=True)          # untimed: executed-step count of the same workload
    steps_done, steps_total = core.last_icm_steps()
    steps_exact = core.last_icm_exact_steps()
    lib_ms = core.last_icm_timings()                        # the library's own CUDA-event phase timers (cross-check)
    # bytes the kernel actually gathers: a 512 B quantised row per (step, other codebook), and the 1 KB fp32 rows again for the steps the pre-filter left undecided AND whose near-tie held more
    # candidates than the windowed evaluation takes (steps_exact; the windowed near-ties read ~1 KB of sectors each and
    # are not counted); memoised steps read nothing
    # the unaries are read once per vector (the warp keeps them in shared memory as 16-bit integers); m > 8 runs one
    # uniform loop over all m rows (the diagonal one is a block of zero words)
    rows = (m - 1) if m <= 8 else m
    gather_bytes = float(steps_done) * rows * H * 2 + float(n) * m * H * 4 + float(steps_exact) * m * H * 4
    # SURVEY 8d algorithmic figure: every reference step gathers (m-1) fp32 rows of 256 entries
    gather_bytes_ref = float(n) * cfg["ilsiter"] * cfg["icmiter"] * m * (m - 1) * H * 4
    qerr = core.qerror(X, Bwork, C)
    qerr0 = core.qerror(X, B0, C)

    # opt-in fast mode (tcgen05 bf16x3 unaries; NOT bit-identical, so never the headline): same workload, reported aside
    Bfast = B0.clone()

    def icm_fast():
        Bfast.copy_(B0)
        core.encode_icm(X, C, Bfast, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0,
                        inplace=True, fast=True)
    fs = max(2, args.steps // 2)
    fast_ms = timed_steps(icm_fast, fs, 1, dist, device) / fs
    fast_obj = {"flag": "RAYUELA_FAST_UNARIES", "ms_per_step": fast_ms, "vectors_per_sec": world * n / (fast_ms * 1e-3),
                "qerror_after": core.qerror(X, Bfast, C),
                "vectors_differing_from_exact": float((Bfast != Bwork).any(1).float().mean().item()),
                "note": "unaries by a tcgen05 bf16x3 GEMM (TMEM accumulators) instead of the exact fp32 kernel; validated "
                        "by tolerance / qerror (tests/test_gpu_fast_mode.py), off by default"}
    del Bfast

    # e2e through the C ABI with pinned host buffers
    Xh, Ch, B0h = pinned(X.cpu()), pinned(C.cpu()), pinned(B0.cpu())
    Bh = pinned(B0.cpu())

    def icm_e2e():
        np.copyto(Bh, B0h)
        core.encode_icm(Xh, Ch, Bh, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0,
                        inplace=True)
    e2e_steps = max(1, min(args.steps, 3))
    icm_e2e_ms = wall_steps(icm_e2e, e2e_steps, 1, dist, device) / e2e_steps
    same = bool(np.array_equal(Bh, Bwork.cpu().numpy()))
    # the same call from PAGEABLE host arrays -- what a Julia caller passes (Julia arrays are not pinned)
    Xp, Cp_, Bp = np.array(Xh), np.array(Ch), np.array(B0h)

    def icm_e2e_pageable():
        core.encode_icm(Xp, Cp_, Bp, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0)
    icm_e2e_pageable_ms = wall_steps(icm_e2e_pageable, 2, 1, dist, device) / 2
    del Xh, Xp

    # ---- path (2): linscan_lsq over the encoded base -----------------------------------------------------------
    Cm = C.reshape(m, H, d)
    rec = torch.zeros(n, d, device=device)
    for j in range(m):
        rec += Cm[j][Bwork[:, j].long()]
    dbnorms = (rec * rec).sum(1).contiguous()
    del rec
    index = core.Index(core.SCAN_LSQ, Bwork, dbnorms, id_offset=g0)
    res = {}
    if os.environ.get("BENCH_DUMP_SCAN"):                 # the scan's inputs, for tools/scan_file.py
        np.savez(os.environ["BENCH_DUMP_SCAN"], B=Bwork.cpu().numpy(), nrm=dbnorms.cpu().numpy(), Q=Q.cpu().numpy(),
                 C=C.cpu().numpy())

    def scan_step():
        dl, il = index.search(Q, C, k)
        if dist is not None:
            gd = torch.empty((world * nq, k), device=device, dtype=dl.dtype)
            gi = torch.empty((world * nq, k), device=device, dtype=il.dtype)
            dist.all_gather_into_tensor(gd, dl)
            dist.all_gather_into_tensor(gi, il)
            dl, il = core.topk_merge(gd.view(world, nq, k), gi.view(world, nq, k))
        res["d"], res["i"] = dl, il

    l1 = rb.launch_count()
    scan_ms = timed_steps(scan_step, args.steps, args.warmup, dist, device)
    scan_launches = timed_steps.last_launches
    scan_per = scan_ms / args.steps
    scan_value = nq / (scan_per * 1e-3)
    # Recall@1 against exact fp32 brute force over the GLOBAL base (or the dataset's own ground truth)
    gt = real["gt"] if real is not None and real["gt"] is not None else exact_nn(X, Q)
    if dist is not None:
        # per-rank best -> global best: compare distances
        dbest = ((X[gt] - Q) ** 2).sum(1)
        alld = torch.empty(world * nq, device=device)
        alli = torch.empty(world * nq, device=device, dtype=torch.long)
        dist.all_gather_into_tensor(alld, dbest)
        dist.all_gather_into_tensor(alli, gt + g0)
        alld, alli = alld.view(world, nq), alli.view(world, nq)
        gt_global = alli.gather(0, alld.argmin(0, keepdim=True))[0]
    else:
        gt_global = gt
    recall1 = float((res["i"][:, 0].long() - 1 == gt_global).float().mean().item())   # ids are 1-based

    Qh, Ch2 = pinned(Q.cpu()), pinned(C.cpu())
    dh = torch.empty((nq, k), dtype=torch.float32, pin_memory=True).numpy()
    ih = torch.empty((nq, k), dtype=torch.int32, pin_memory=True).numpy()

    def scan_e2e():
        index.search(Qh, Ch2, k, out=(dh, ih))     # pinned host queries/codebooks in, pinned host results out
        res["dh"], res["ih"] = dh, ih
    scan_e2e_ms = wall_steps(scan_e2e, e2e_steps, 1, None, device) / e2e_steps

    # the demos' knn = 1000 (demos/demos_train_query_base.jl:16) on the same index, same exchange
    kk = 1000
    scan_k1000 = None
    if not args.no_strong and k != kk:
        def scan_step_k():
            dl, il = index.search(Q, C, kk)
            if dist is not None:
                gd = torch.empty((world * nq, kk), device=device, dtype=dl.dtype)
                gi = torch.empty((world * nq, kk), device=device, dtype=il.dtype)
                dist.all_gather_into_tensor(gd, dl)
                dist.all_gather_into_tensor(gi, il)
                dl, il = core.topk_merge(gd.view(world, nq, kk), gi.view(world, nq, kk))
            res["ik"] = il
        ks = max(2, args.steps // 2)
        msk = timed_steps(scan_step_k, ks, 1, dist, device) / ks
        hit = (res["ik"].long() - 1 == gt_global[:, None])
        scan_k1000 = {"k": kk, "ms_per_step": msk, "queries_per_sec": nq / (msk * 1e-3),
                      "recall_at_1000": float((hit.sum(1) == 1).float().mean().item()), "steps": ks}

    # opt-in tensor-core LUT (RAYUELA_FAST_LUT; not bit-identical, so never the headline)
    fast_scan = None
    if not args.no_strong:
        def scan_fast():
            res["f"] = index.search(Q, C, k, fast=True)
        fsn = max(2, args.steps // 2)
        msf = timed_steps(scan_fast, fsn, 1, None, device) / fsn
        fast_scan = {"flag": "RAYUELA_FAST_LUT", "ms_per_step_local_shard": msf, "queries_per_sec_local_shard": nq / (msf * 1e-3),
                     "top1_agreement_with_exact": float((res["f"][1][:, 0] == index.search(Q, C, k)[1][:, 0]).float().mean().item())}

    # sub-range of the timed workload for the untimed oracle parity check (host copies, taken before X is freed)
    off = (n // 2) // 8 * 8
    cnt = min(1536, n - off)
    parity_in = (X[off:off + cnt].cpu().numpy(), C.cpu().numpy(), B0[off:off + cnt].cpu().numpy(),
                 Bwork[off:off + cnt].cpu().numpy(), off)

    # ---- strong-scaling objects (every rank takes part; rank 0 keeps the result) ---------------------------------
    index.free()
    strong = {}
    if not args.no_strong and 8 % world != 0:
        strong["config4"] = strong["config5"] = {"skipped": "the strong-scaling bases are cut in 8 fixed blocks: N must divide 8"}
    elif not args.no_strong:
        del X, Bwork, B0, Bscratch
        torch.cuda.empty_cache()
        strong["config4"] = run_config4(dist, device, world, rank)
        torch.cuda.empty_cache()
        strong["config5"] = run_config5(dist, device, world, rank)
    clocks = sampler.stop() if sampler else {}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- untimed parity check of the timed configuration against the oracle (sub-range, full ilsiter) ---------------
    parity = None
    if not args.no_cpu_baseline:
        parity = parity_subrange(cfg, parity_in, g0)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample ---------------------------------------------------
    cpu = None
    cpu_scan = None
    if world == 1 and not args.no_cpu_baseline:
        cores, omp = force_host_threads()
        n_s, t, kind, ph = cpu_sized(lambda s_: cpu_icm_sample(cfg, s_), 16384, cap=n)
        cpu = {"value": n_s / t, "unit": "vectors/s", "cores": cores, "omp_max_threads": omp, "kind": kind,
               "sample": "%d of %d vectors, one encode at the same m/ilsiter/icmiter/npert (%.1f s)" % (n_s, n, t),
               "phases_s": ph, "note": CPU_NOTE}
        nq_s, t2, kind2, _ = cpu_sized(lambda s_: cpu_scan_sample(cfg, s_, n), max(64, 4 * cores), cap=nq)
        cpu_scan = {"value": nq_s / t2, "unit": "queries/s", "cores": cores, "omp_max_threads": omp, "kind": kind2,
                    "sample": "%d of %d queries over the full %d-code base (%.1f s)" % (nq_s, nq, n, t2)}

    sm_mhz = clocks.get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
    onchip_peak = 148 * 128 * sm_mhz * 1e6 / 1e9          # GB/s of L1/shared load bandwidth at the sampled clock
    tr = ncu_traffic("icm_m%d_n%d_ils%d" % (m, n, cfg["ilsiter"]))
    achieved = gather_bytes_ref / (k3_ms * 1e-3) / 1e9
    icm_roof = {"bound": "L2 gather bandwidth / issue (ncu at HEAD: lts throughput %s %%, issue active %s %%; DRAM ~1 %% busy, so "
                         "neither HBM nor tensor)" % ((tr or {}).get("lts_throughput_pct", "?"), (tr or {}).get("issue_active_pct", "?")),
                "achieved": achieved, "peak": onchip_peak, "unit": "GB/s", "frac": achieved / onchip_peak,
                "frac_actual": gather_bytes / (k3_ms * 1e-3) / 1e9 / onchip_peak,
                "traffic": tr["bytes"] if tr else None, "traffic_source": tr,
                "peak_source": "148 SM x 128 B/clk x sampled SM clock (SURVEY 8d: the path is not HBM-bound)",
                "hbm_peak": pk["hbm_gbs"], "hbm_peak_source": pk_src, "hbm_frac_of_algorithmic": achieved / pk["hbm_gbs"],
                "kernel": "icm_warp_kernel<%d,true,%s>" % (m, "true" if m <= 8 else "false"), "kernel_ms": k3_ms,
                "kernel_ms_library_events": lib_ms,
                "steps_executed": steps_done, "steps_reference": steps_total, "steps_exact_rows": steps_exact,
                "gathered_actual": gather_bytes / (k3_ms * 1e-3) / 1e9,
                "note": "achieved = SURVEY 8d algorithmic (work-equivalent) bytes n*ilsiter*icmiter*m*(m-1)*256*4 -- what "
                        "the reference's steps gather -- / K3 time, against the on-chip load ceiling.  frac_actual = "
                        "the bytes the kernel really gathers (memoised steps skipped, 512 B rows of 14-bit fields, "
                        "unaries once per vector, near-ties resolved on the window's candidates) against the same "
                        "ceiling; rows are served by L2, whose gather bandwidth (ncu lts throughput) is what binds "
                        "(profiles/).  `traffic` = ncu dram bytes of one K3 launch at HEAD"}
    scan_bytes = float(nq) * n * (m + 4)                      # per rank: its shard of the base
    tr2 = ncu_traffic("scan_m%d_n%d_nq%d_k%d" % (m, n, nq, k))
    lookups = float(nq) * n * m / (scan_per * 1e-3)
    prefilter = os.environ.get("RAYUELA_B200_SCAN_PREFILTER", "1") != "0"
    per_clk = 64 if prefilter else 32                         # (query, code byte) lookups per clock per SM the pipe can serve
    scan_roof = {"bound": "hbm", "achieved": scan_bytes / (scan_per * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                 "unit": "GB/s", "traffic": tr2["bytes"] if tr2 else None, "traffic_source": tr2, "peak_source": pk_src,
                 "kernel": "scanx_kernel<%d,true,%s,%s>" % (8 if m <= 8 else 16, "true" if k >= 16 else "false",
                                                            "true" if prefilter else "false"),
                 "smem_lookups_per_s": lookups, "smem_lookup_peak": 148 * per_clk * sm_mhz * 1e6,
                 "smem_lookup_frac": lookups / (148 * per_clk * sm_mhz * 1e6),
                 "note": "algorithmic bytes = nq*n*(m+4): what the reference streams per query "
                         "(pairwise_byte.cpp:56-83); whole search step (LUT + scan + merge) in the denominator; frac "
                         "can exceed 1 because codes are re-read from L2 across a query tile (see traffic); the "
                         "ceiling that binds is the shared-memory gather rate (smem_lookup_frac): 128 B/clk/SM = 32 "
                         "fp32 lookups, or 64 lookups of the quantised pre-filter's 2-byte entries (two queries per "
                         "fp32 word; survivors re-evaluated exactly, results bit-identical)"}
    scan_roof["frac"] = scan_roof["achieved"] / scan_roof["peak"]

    primary_icm = args.path == "icm"
    line = {
        "metric": "lsq_icm_encode_vectors_per_sec" if primary_icm else "linscan_lsq_queries_per_sec",
        "value": icm_value if primary_icm else scan_value,
        "unit": "vectors/s" if primary_icm else "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": icm_per if primary_icm else scan_per,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "SIFT1M (%s)" % args.data_dir if real is not None else "synthetic",
        "config": workload_config(cfg, args),
        "clocks": clocks,
        "e2e": ({"value": world * n / (icm_e2e_ms * 1e-3), "unit": "vectors/s",
                 "h2d_bytes_per_step": n * d * 4 + m * H * d * 4 + n * m, "d2h_bytes_per_step": n * m,
                 "ms_per_step": icm_e2e_ms, "codes_equal_device_path": same}
                if primary_icm else
                {"value": nq / (scan_e2e_ms * 1e-3), "unit": "queries/s",
                 "h2d_bytes_per_step": nq * d * 4 + m * H * d * 4, "d2h_bytes_per_step": nq * k * 8,
                 "ms_per_step": scan_e2e_ms}),
        "gpu_launches": icm_launches if primary_icm else scan_launches,
        "roofline": icm_roof if primary_icm else scan_roof,
        "cpu_baseline": cpu if primary_icm else cpu_scan,
        "parity_checked": parity,
        "icm": {"vectors_per_sec": icm_value, "vector_ils_iters_per_sec": icm_value * cfg["ilsiter"],
                "ms_per_step": icm_per, "setup_ms(K0+K1+K2+cost)": setup_ms, "qerror_before": qerr0,
                "qerror_after": qerr, "e2e_vectors_per_sec": world * n / (icm_e2e_ms * 1e-3),
                "e2e_pageable_host_arrays_vectors_per_sec": world * n / (icm_e2e_pageable_ms * 1e-3),
                "e2e_pageable_ms_per_step": icm_e2e_pageable_ms, "roofline": icm_roof,
                "cpu_baseline": cpu, "gpu_launches": icm_launches, "fast_mode": fast_obj},
        "linscan": {"metric": "linscan_lsq_queries_per_sec", "queries_per_sec": scan_value, "recall_at_1": recall1,
                    "k": k, "nq": nq, "n_base_total": world * n, "ms_per_step": scan_per,
                    "e2e_queries_per_sec": nq / (scan_e2e_ms * 1e-3), "roofline": scan_roof,
                    "cpu_baseline": cpu_scan, "gpu_launches": scan_launches,
                    "sharding": "base-sharded, one all-gather of per-shard top-k + merge" if world > 1 else "none",
                    "k1000": scan_k1000, "fast_mode": fast_scan},
    }
    line.update(strong)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def parity_subrange(cfg, parity_in, g0, count=1536):
    """Untimed: the oracle (fixed-order restatement + the reference's compiled `condition`) encodes `count` vectors
    of the TIMED workload at the timed ilsiter, and the GPU's codes for the same global indices must be identical."""
    from oracle import oracle as orc
    force_host_threads()                       # also under torchrun (OMP_NUM_THREADS=1): the oracle is OpenMP-parallel
    Xs, Cs, B0s, got, off = parity_in
    want = orc.encode_icm(Xs, Cs, B0s, cfg["ilsiter"], cfg["icmiter"], cfg["npert"], True, seed=2024, g0=g0 + off,
                          use_ref_step=orc.have_ref())
    eq = bool(np.array_equal(want["B"], got))
    return {"vectors": int(Xs.shape[0]), "first_global_index": int(g0 + off), "ilsiter": cfg["ilsiter"],
            "codes_bit_identical_to_oracle": eq,
            "mismatching_vectors": int((want["B"] != got).any(1).sum()),
            "oracle": "oracle.encode_icm(use_ref_step=%s)" % orc.have_ref()}


# ---------------------------------------------------------------------------------------------------------
# strong scaling: BASELINE.json configs[3] and configs[4]
# ---------------------------------------------------------------------------------------------------------
def _blocks_of(rank, world, nblocks=8):
    per = nblocks // world
    return list(range(rank * per, (rank + 1) * per))


def _allsum_i64(v, dist, device):
    import torch
    t = torch.tensor([v], device=device, dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(t)
    return int(t.item())


def run_config4(dist, device, world, rank, n_total=1_000_000, m=16, d=128, ilsiter=32, steps=2):
    """configs[3]: LSQ++ m=16 h=256, 1M x 128 in total, encoding sharded over the N ranks (n/N vectors each, contiguous
    splitarray slices, RNG keyed on the global index -> codes independent of N: `codes_checksum` must agree for
    every N).  The base is generated in 8 fixed blocks of 125k vectors so every N sees the same data."""
    import torch
    from rayuela_b200 import core
    assert 8 % world == 0 and n_total % 8 == 0
    nb = n_total // 8
    blocks = _blocks_of(rank, world)
    Xt, _ = make_data(50000, 1, d, seed=4000, device=device, rank=0)
    C = train_codebooks(Xt, m, device)                              # identical on every rank
    X = torch.cat([make_data(nb, 1, d, seed=5000 + b, device=device, rank=0)[0] for b in blocks])
    B0 = torch.cat([torch.randint(0, H, (nb, m), device=device, dtype=torch.uint8,
                                  generator=torch.Generator(device=device).manual_seed(6000 + b)) for b in blocks])
    g0 = blocks[0] * nb
    Bw = B0.clone()

    def step():
        Bw.copy_(B0)
        core.encode_icm(X, C, Bw, ilsiter, 4, 4, True, seed=2024, g0=g0, inplace=True)
    ms = timed_steps(step, steps, 1, dist, device) / steps
    idx = torch.arange(g0, g0 + Bw.shape[0], device=device, dtype=torch.int64)
    local = int((Bw.long().sum(1) * (idx % 1000003 + 1)).sum().item())     # order-independent, position-sensitive
    checksum = _allsum_i64(local, dist, device) % (1 << 61)
    q0 = core.qerror(X, B0, C)
    q1 = core.qerror(X, Bw, C)
    qsum = torch.tensor([q0 * Bw.shape[0], q1 * Bw.shape[0]], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(qsum)
    return {"workload": "BASELINE.json configs[3]: LSQ++ m=16 h=256 icmiter=4 npert=4 randord ilsiter=%d, %dx%d in "
                        "total, %d vectors per GPU" % (ilsiter, n_total, d, n_total // world),
            "scaling": "strong", "n_total": n_total, "n_per_gpu": n_total // world, "ms_per_step": ms,
            "vectors_per_sec": n_total / (ms * 1e-3), "steps": steps, "warmup": 1,
            "qerror_before": float(qsum[0].item() / n_total), "qerror_after": float(qsum[1].item() / n_total),
            "codes_checksum": checksum, "collective": "none (disjoint code slices)"}


def _exact_scan_topk(codes, norms, Q, C, m, k, id0):
    """Independent restatement of linscan_aqd_query_extra_byte (deps/src/linscan_aqd_pairwise_byte.cpp:42-83) in
    torch element-wise fp32 ops in the reference's operation order (LUT: t -= (2*q[s])*c[s] for ascending s;
    distance ((t_0 + t_1) + ...) + norm), so it is bit-identical; returns the k smallest signed 64-bit keys
    (ordered(dist) << 32 | 1-based global id) per query."""
    import torch
    nqv, d = Q.shape
    lut = torch.zeros(nqv, m * H, device=Q.device)
    for s in range(d):
        lut = lut - (2.0 * Q[:, s:s + 1]) * C[None, :, s]
    lut = lut.view(nqv, m, H)
    best = None
    step = 1 << 20
    for a in range(0, codes.shape[0], step):
        cb = codes[a:a + step].long()
        acc = lut[:, 0, :][:, cb[:, 0]]
        for j in range(1, m):
            acc = acc + lut[:, j, :][:, cb[:, j]]
        acc = acc + norms[None, a:a + step] + 0.0
        bits = acc.view(torch.int32).long()
        ordered = torch.where(bits < 0, bits ^ 0x7FFFFFFF, bits)                    # signed order == float order
        ids = torch.arange(a, a + cb.shape[0], device=Q.device, dtype=torch.int64) + (id0 + 1)
        keys = (ordered << 32) | ids[None, :]
        cat = keys if best is None else torch.cat([best, keys], 1)
        best = torch.topk(cat, min(k, cat.shape[1]), dim=1, largest=False, sorted=True).values
    return best


def run_config5(dist, device, world, rank, n_total=100_000_000, m=8, d=128, nq=10_000, steps=2, nverify=64):
    """configs[4]: linscan_lsq over 100M x 8 B uniform random codes (+ fp32 norms), 10k queries, base sharded over the
    N ranks (global ids), one all-gather of the per-shard top-k + merge; k = 1 and k = 1000 (the demos' knn).  The ids
    of a 64-query subset are checked against an independent bit-exact restatement of the reference's scan."""
    import torch
    from rayuela_b200 import core
    assert 8 % world == 0 and n_total % 8 == 0
    nb = n_total // 8
    blocks = _blocks_of(rank, world)
    gq = torch.Generator(device=device).manual_seed(7000)
    Q = torch.randn(nq, d, generator=gq, device=device)
    C = (torch.randn(m * H, d, generator=gq, device=device) / m ** 0.5).contiguous()
    codes = torch.cat([torch.randint(0, H, (nb, m), device=device, dtype=torch.uint8,
                                     generator=torch.Generator(device=device).manual_seed(8000 + b)) for b in blocks])
    norms = torch.cat([torch.randn(nb, device=device,
                                   generator=torch.Generator(device=device).manual_seed(9000 + b)) * 4 + 30
                       for b in blocks]).contiguous()
    id0 = blocks[0] * nb
    index = core.Index(core.SCAN_LSQ, codes, norms, id_offset=id0)
    out = {"workload": "BASELINE.json configs[4]: linscan_lsq, %d x %d B uniform random codes + fp32 norms in total, "
                       "%d queries, base sharded over the GPUs (%d codes each), all-gather of per-shard top-k + merge"
                       % (n_total, m, nq, n_total // world),
           "scaling": "strong", "n_total": n_total, "n_per_gpu": n_total // world, "nq": nq, "steps": steps,
           "warmup": 1}
    res = {}
    for k in (1, 1000):
        gd = torch.empty((world * nq, k), device=device, dtype=torch.float32)
        gi = torch.empty((world * nq, k), device=device, dtype=torch.int32)

        def search():
            res["l"] = index.search(Q, C, k)

        def gather():
            if dist is not None:
                dist.all_gather_into_tensor(gd, res["l"][0])
                dist.all_gather_into_tensor(gi, res["l"][1])

        def merge():
            res["g"] = core.topk_merge(gd.view(world, nq, k), gi.view(world, nq, k)) if dist is not None else res["l"]

        def step():
            search()
            gather()
            merge()
        ms = timed_steps(step, steps, 1, dist, device) / steps
        # where the step goes (separately timed sections, same inputs)
        ms_search = timed_steps(search, 1, 0, dist, device)
        ms_gather = timed_steps(gather, 1, 0, dist, device) if dist is not None else 0.0
        ms_merge = timed_steps(merge, 1, 0, dist, device) if dist is not None else 0.0
        dg, ig = res["g"]
        # verification: 64-query subset against the bit-exact restatement, merged over the ranks
        qs = torch.arange(0, nq, nq // nverify, device=device)[:nverify]
        keys = _exact_scan_topk(codes, norms, Q[qs], C, m, k, id0)
        if dist is not None:
            allk = torch.empty((world,) + tuple(keys.shape), device=device, dtype=keys.dtype)
            dist.all_gather_into_tensor(allk.view(-1, keys.shape[1]), keys)
            keys = torch.topk(allk.permute(1, 0, 2).reshape(keys.shape[0], -1), k, dim=1, largest=False,
                              sorted=True).values
        want_ids = (keys & 0xFFFFFFFF).to(torch.int32)
        ok = bool(torch.equal(want_ids, ig[qs]))
        out["k%d" % k] = {"ms_per_step": ms, "queries_per_sec": nq / (ms * 1e-3), "ms_search_local": ms_search,
                          "ms_allgather": ms_gather, "ms_merge": ms_merge,
                          "allgather_bytes_per_rank": nq * k * 8 if world > 1 else 0,
                          "ids_checksum": int(ig.long().sum().item()) % (1 << 61),
                          "ids_verified_queries": int(qs.numel()), "ids_equal_exact_restatement": ok,
                          "algorithmic_GBps": float(nq) * n_total * (m + 4) / (ms * 1e-3) / 1e9}
    index.free()
    return out


def main():
    # stdout carries exactly one JSON line: route everything else (NCCL banners, library chatter) to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="icm", choices=["icm", "linscan"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--ilsiter", type=int, default=32)
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the config4 / config5 / k=1000 objects")
    ap.add_argument("--data-dir", default=os.environ.get("RAYUELA_DATA_DIR", ""),
                    help="directory holding sift/sift_{learn,base,query}.fvecs + sift_groundtruth.ivecs: bench on the "
                         "real SIFT1M base (true Recall@1) instead of the synthetic one")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference" or os.environ.get("BENCH_ALLOW_SHORT"), "warmup must be >= 3"
    cfg = dict(n=args.n, nq=args.nq, m=args.m, d=args.d, k=args.k, ilsiter=args.ilsiter, icmiter=4, npert=4)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
