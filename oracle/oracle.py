"""ctypes front-end to oracle/liboracle.so and oracle/_ref/*.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under rayuela.jl_b200/ does.

Array conventions (numpy, C-contiguous) are the memory images Julia's ccall would pass:
    X      (n, d)    float32   == Julia d-by-n
    C      (m*h, d)  float32   == Julia hcat(C...) d-by-(m*h)      (LSQ/CQ paths)
    Cpq    (m*h, sub) float32  == Julia cat(C..., dims=3) sub-by-h-by-m   (PQ path)
    B      (n, m)    uint8 0-based == Julia m-by-n
    dists  (nq, k)   float32   == Julia k-by-nq
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = ct.POINTER(ct.c_float)
_u8p = ct.POINTER(ct.c_uint8)
_i32p = ct.POINTER(ct.c_int32)

CONDITION_FN = ct.CFUNCTYPE(None, _u8p, _f32p, _f32p, _f32p, _i32p, _i32p, ct.c_int, ct.c_int, ct.c_int)


def build(force=False):
    """Compile liboracle.so (+ _ref/ when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rayuela_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/deps/src") and (force or not have_ref()):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def have_ref():
    return all(os.path.exists(os.path.join(_HERE, "_ref", f))
               for f in ("encode_icm.so", "linscan_aqd.so", "linscan_aqd_pairwise_byte.so"))


_lib = None
_ref = {}


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ct.CDLL(os.path.join(_HERE, "liboracle.so"))
        _lib.orc_qerror.restype = ct.c_double
    return _lib


def ref(name):
    """The reference's own compiled C++ (oracle/_ref/<name>.so), or None."""
    if name not in _ref:
        p = os.path.join(_HERE, "_ref", name + ".so")
        _ref[name] = ct.CDLL(p) if os.path.exists(p) else None
    return _ref[name]


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def philox(ctr, k0, k1):
    c = np.array(ctr, dtype=np.uint32)
    lib().orc_philox(c.ctypes.data_as(ct.POINTER(ct.c_uint32)), ct.c_uint32(k0), ct.c_uint32(k1))
    return c


def randperm(seed, it, m):
    perm = np.zeros(m, dtype=np.int32)
    lib().orc_randperm(ct.c_uint64(seed), ct.c_int(it), ct.c_int(m), _p(perm, _i32p))
    return perm


def perturb_codes(B, h, npert, seed, it, g0=0):
    B = np.ascontiguousarray(B, dtype=np.uint8).copy()
    n, m = B.shape
    lib().orc_perturb_codes(_p(B, _u8p), ct.c_int64(n), m, h, npert, ct.c_uint64(seed), it, ct.c_int64(g0))
    return B


def get_unaries(X, C, m, h=256):
    X, C = _f32(X), _f32(C)
    n, d = X.shape
    U = np.empty((m, n, h), dtype=np.float32)
    lib().orc_get_unaries(_p(X, _f32p), _p(C, _f32p), ct.c_int64(n), d, m, h, _p(U, _f32p))
    return U


def get_binaries(C, m, h=256):
    C = _f32(C)
    d = C.shape[1]
    ncbi = m * (m - 1) // 2
    b = np.empty((max(ncbi, 1), h, h), dtype=np.float32)
    bt = np.empty((max(ncbi, 1), h, h), dtype=np.float32)
    cbi = np.zeros((max(ncbi, 1), 2), dtype=np.int32)
    lib().orc_get_binaries(_p(C, _f32p), d, m, h, _p(b, _f32p), _p(bt, _f32p), _p(cbi, _i32p))
    return b[:ncbi], bt[:ncbi], cbi[:ncbi]


def condition(B, ub, binaries, binaries_t, pair2idx, to_condition, j, use_ref=False):
    """One ICM step in place on (B, ub).  use_ref -> the reference's compiled `condition`."""
    n, m = B.shape
    fn = ref("encode_icm").condition if use_ref else lib().orc_condition
    fn(_p(B, _u8p), _p(ub, _f32p), _p(binaries, _f32p), _p(binaries_t, _f32p),
       _p(np.ascontiguousarray(pair2idx, dtype=np.int32), _i32p),
       _p(np.ascontiguousarray(to_condition, dtype=np.int32), _i32p), ct.c_int(j), ct.c_int(n), ct.c_int(m))


def veccost(X, B, C, h=256):
    X, C = _f32(X), _f32(C)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    n, d = X.shape
    m = B.shape[1]
    cost = np.empty(n, dtype=np.float32)
    lib().orc_veccost(_p(X, _f32p), _p(B, _u8p), _p(C, _f32p), ct.c_int64(n), d, m, h, _p(cost, _f32p))
    return cost


def qerror(X, B, C, h=256):
    return float(np.mean(veccost(X, B, C, h).astype(np.float64)))


def set_num_threads(t):
    """Force the OpenMP team size used by liboracle AND oracle/_ref (one libgomp per process) and the BLAS pool
    numpy uses; returns what OpenMP reports afterwards.  torchrun exports OMP_NUM_THREADS=1, and an environment
    variable set after the runtimes are loaded changes nothing, hence the runtime calls."""
    lib().orc_set_num_threads(ct.c_int(int(t)))
    try:
        from threadpoolctl import threadpool_limits
        global _blas_limit
        _blas_limit = threadpool_limits(limits=int(t))      # kept alive: the limit lasts as long as the object
    except Exception:
        pass
    # spin the BLAS pool up now, not inside a timed call: the first dozen sgemm calls of a process take ~0.1 s each
    import time
    w = np.ones((256, 128), dtype=np.float32)
    fast = 0
    for _ in range(64):
        t0 = time.perf_counter()
        np.matmul(w, w.T)
        fast = fast + 1 if time.perf_counter() - t0 < 5e-3 else 0
        if fast >= 3:
            break
    return max_threads()


_blas_limit = None


def max_threads():
    return int(lib().orc_get_max_threads())


def blas_unaries(X, C, m, h=256):
    """get_unaries as the reference computes it (src/utils.jl:135-144): one sgemm per codebook,
    unaries[i] = (-2*C[i]')*X, then + diag(C[i]'C[i]).  TIMING ARM ONLY: the BLAS summation order is not the
    oracle's pinned order.  Returns U (m, n, h)."""
    X, C = _f32(X), _f32(C)
    n = X.shape[0]
    U = np.empty((m, n, h), dtype=np.float32)
    for i in range(m):
        Ci = C[i * h:(i + 1) * h]
        np.matmul(X, (np.float32(-2) * Ci).T, out=U[i])
        U[i] += np.einsum("cd,cd->c", Ci, Ci)[None, :]
    return U


def blas_binaries(C, m, h=256):
    """get_binaries + binaries_t via sgemm (src/utils.jl:164, src/LSQ.jl:180-183).  TIMING ARM ONLY."""
    C = _f32(C)
    ncbi = m * (m - 1) // 2
    b = np.empty((max(ncbi, 1), h, h), dtype=np.float32)
    bt = np.empty((max(ncbi, 1), h, h), dtype=np.float32)
    idx = 0
    for i in range(m):
        for j in range(i + 1, m):
            # column-major h-by-h 2*C_i'*C_j -> C-order image [b][a] = 2<C_i[:,a], C_j[:,b]>
            np.matmul(np.float32(2) * C[j * h:(j + 1) * h], C[i * h:(i + 1) * h].T, out=b[idx])
            bt[idx] = b[idx].T
            idx += 1
    return b, bt


def encode_icm(X, C, B, ilsiter, icmiter, npert, randord, seed=0, g0=0, orders=None, h=256,
               snap_iters=None, use_ref_step=False, blas=False, want_phases=False):
    """encode_icm_fully! restatement.  Returns dict(B, cost, stats, B_snap, objs[, phases]).
    blas=True (bench.py's CPU arm only): unaries and tables come from sgemm like the reference's, instead of the
    pinned fixed-order chains; phases = seconds in {tables, unaries, icm_steps, veccost, other}."""
    import time
    X, C = _f32(X), _f32(C)
    B = np.ascontiguousarray(B, dtype=np.uint8).copy()
    n, d = X.shape
    m = B.shape[1]
    cost = np.empty(n, dtype=np.float32)
    stats = np.zeros((max(ilsiter, 1), 2), dtype=np.int32)
    if orders is not None:
        orders = np.ascontiguousarray(orders, dtype=np.int32)
        assert orders.shape == (ilsiter, m)
    snaps = np.ascontiguousarray(snap_iters if snap_iters is not None else [], dtype=np.int32)
    ns = len(snaps)
    Bs = np.zeros((max(ns, 1), n, m), dtype=np.uint8)
    objs = np.zeros(max(ns, 1), dtype=np.float32)
    step = None
    if use_ref_step:
        step = ct.cast(ref("encode_icm").condition, CONDITION_FN)
    U = bn = bt = None
    t_tab = t_un = 0.0
    if blas:
        t0 = time.perf_counter()
        bn, bt = blas_binaries(C, m, h)
        t_tab = time.perf_counter() - t0
        t0 = time.perf_counter()
        U = blas_unaries(X, C, m, h)
        t_un = time.perf_counter() - t0
    ph = np.zeros(5, dtype=np.float64)
    f64p = ct.POINTER(ct.c_double)
    fn = lib().orc_encode_icm_fully_ex
    fn.argtypes = [_f32p, _f32p, _u8p, ct.c_int64, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int,
                   ct.c_int, ct.c_uint64, ct.c_int64, _i32p, CONDITION_FN, _i32p, ct.c_int, _u8p, _f32p, _f32p, _i32p,
                   _f32p, _f32p, _f32p, f64p]
    rc = fn(_p(X, _f32p), _p(C, _f32p), _p(B, _u8p), n, d, m, h, ilsiter, icmiter, npert, int(bool(randord)),
            seed, g0, _p(orders, _i32p) if orders is not None else None,
            step if step is not None else ct.cast(None, CONDITION_FN),
            _p(snaps, _i32p) if ns else None, ns, _p(Bs, _u8p), _p(objs, _f32p), _p(cost, _f32p), _p(stats, _i32p),
            _p(U, _f32p) if blas else None, _p(bn, _f32p) if blas else None, _p(bt, _f32p) if blas else None,
            ph.ctypes.data_as(f64p))
    if rc != 0:
        raise RuntimeError("orc_encode_icm_fully failed: %d" % rc)
    out = dict(B=B, cost=cost, stats=stats[:ilsiter], B_snap=Bs[:ns], objs=objs[:ns])
    if want_phases:
        out["phases"] = {"tables": float(ph[0] + t_tab), "unaries": float(ph[1] + t_un), "icm_steps": float(ph[2]),
                         "veccost": float(ph[3]), "other": float(ph[4])}
    return out


LSQ, CQ, PQ = 0, 1, 2


def linscan(kind, B, Xq, codebooks, k, dbnorms=None, id_offset=0, h=256):
    """Restated scan.  kind LSQ/CQ: codebooks (m*h, d); kind PQ: (m*h, sub).  Returns dists, idx (nq,k)."""
    B = np.ascontiguousarray(B, dtype=np.uint8)
    Xq, codebooks = _f32(Xq), _f32(codebooks)
    n, m = B.shape
    nq, d = Xq.shape
    dists = np.zeros((nq, k), dtype=np.float32)
    idx = np.zeros((nq, k), dtype=np.int32)
    nrm = _f32(dbnorms) if dbnorms is not None else None
    rc = lib().orc_linscan(kind, _p(dists, _f32p), _p(idx, _i32p), _p(B, _u8p), _p(Xq, _f32p), _p(codebooks, _f32p),
                           _p(nrm, _f32p) if nrm is not None else None, nq, ct.c_int64(n), m, h, d, k,
                           ct.c_int64(id_offset))
    if rc != 0:
        raise RuntimeError("orc_linscan failed: %d" % rc)
    return dists, idx


def ref_linscan(kind, B, Xq, codebooks, k, dbnorms=None, h=256):
    """The reference's own compiled scan symbols, called with the buffers Julia passes
    (src/Linscan.jl:19-23, 135-141, 173-179)."""
    B = np.ascontiguousarray(B, dtype=np.uint8)
    Xq, codebooks = _f32(Xq), _f32(codebooks)
    n, m = B.shape
    nq, d = Xq.shape
    dists = np.zeros((nq, k), dtype=np.float32)
    if kind == PQ:
        res = np.zeros((nq, k), dtype=np.uint32)
        ref("linscan_aqd").linscan_aqd_query(
            _p(dists, _f32p), res.ctypes.data_as(ct.POINTER(ct.c_uint32)), _p(B, _u8p), _p(codebooks, _f32p),
            _p(Xq, _f32p), ct.c_int(n), ct.c_uint32(nq), ct.c_int(8 * m), ct.c_int(k), ct.c_int(m), ct.c_int(d),
            ct.c_int(d // m))
        return dists, res.astype(np.int32)
    idx = np.zeros((nq, k), dtype=np.int32)
    so = ref("linscan_aqd_pairwise_byte")
    if kind == LSQ:
        nrm = _f32(dbnorms)
        so.linscan_aqd_query_extra_byte(_p(dists, _f32p), _p(idx, _i32p), _p(B, _u8p), _p(Xq, _f32p),
                                        _p(codebooks, _f32p), _p(nrm, _f32p), nq, n, m, h, d, k)
    else:
        so.linscan_aqd_cq_query_extra_byte(_p(dists, _f32p), _p(idx, _i32p), _p(B, _u8p), _p(Xq, _f32p),
                                           _p(codebooks, _f32p), nq, n, m, h, d, k)
    return dists, idx


def quantize_pq(X, Cpq, m, h=256):
    X, Cpq = _f32(X), _f32(Cpq)
    n, d = X.shape
    B = np.zeros((n, m), dtype=np.uint8)
    lib().orc_quantize_pq(_p(X, _f32p), _p(Cpq, _f32p), ct.c_int64(n), d, m, h, _p(B, _u8p))
    return B


def eval_recall(gt, idx, k):
    """eval_recall (src/Linscan.jl:196-234): rank = position of the true NN if it appears EXACTLY
    once in the list, else k+1; recall@i = fraction of queries with rank <= i.  gt, idx share a base."""
    gt = np.asarray(gt).reshape(-1)
    idx = np.asarray(idx)
    nq = idx.shape[0]
    hit = idx[:, :k] == gt[:, None]
    cnt = hit.sum(1)
    rank = np.where(cnt == 1, hit.argmax(1) + 1, k + 1)
    return np.array([(rank <= i).sum() / nq for i in range(1, k + 1)])


def fast_bin_matmul(X, B, h=256, rho=1e-4):
    """fast_bin_matmul restatement: returns A ((m*h, m*h) float64, symmetric) and b as its column-major image
    viewed C-order, i.e. shape (d, m*h) with b[t, i*h + c]."""
    X = _f32(X)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    n, d = X.shape
    m = B.shape[1]
    A = np.zeros((m * h, m * h), dtype=np.float64)
    b = np.zeros((d, m * h), dtype=np.float64)
    f64p = ct.POINTER(ct.c_double)
    lib().orc_fast_bin_matmul(_p(X, _f32p), _p(B, _u8p), ct.c_int64(n), d, m, h, ct.c_double(rho),
                              A.ctypes.data_as(f64p), b.ctypes.data_as(f64p))
    return A, b


def update_codebooks_fast_bin(X, B, h=256, rho=1e-4):
    """update_codebooks_fast_bin (src/codebook_update.jl:175-204): LAPACK getrf + getrs on (A, b), result
    converted to Float32.  Returns the codebooks as the (m*h, d) image of hcat(C...)."""
    from scipy.linalg import lu_factor, lu_solve
    A, b = fast_bin_matmul(X, B, h, rho)
    C = lu_solve(lu_factor(A), b.T)          # (m*h, d)
    return C.astype(np.float32)


def quantize_norms(B, C, cbnorms=None, h=256):
    """quantize_norms restatement: returns (norm_codes uint8 0-based or None, norms float32)."""
    B = np.ascontiguousarray(B, dtype=np.uint8)
    C = _f32(C)
    n, m = B.shape
    d = C.shape[1]
    norms = np.empty(n, dtype=np.float32)
    codes = np.empty(n, dtype=np.uint8) if cbnorms is not None else None
    cb = _f32(cbnorms) if cbnorms is not None else None
    lib().orc_quantize_norms(_p(B, _u8p), _p(C, _f32p), _p(cb, _f32p) if cb is not None else None, ct.c_int64(n), d,
                             m, h, _p(codes, _u8p) if codes is not None else None, _p(norms, _f32p))
    return codes, norms


VITERBI_FN = ct.CFUNCTYPE(None, _u8p, _f32p, _f32p, ct.c_int, ct.c_int)


def viterbi_encoding(unaries, binaries, m, use_ref=False):
    """viterbi_encoding with the reference's argument layout: unaries (n, m*256), binaries (m-1, 256, 256) with
    binaries[i][j, k] = cost of going from state k of codebook i to state j of codebook i+1."""
    U = _f32(unaries)
    bb = _f32(binaries)
    n = U.shape[0]
    B = np.zeros((n, m), dtype=np.uint8)
    fn = ref("encode_icm").viterbi_encoding if use_ref else lib().orc_viterbi_encoding
    fn(_p(B, _u8p), _p(U, _f32p), _p(bb, _f32p), ct.c_int(n), ct.c_int(m))
    return B


def quantize_chainq(X, C, m, h=256, use_ref=False):
    X, C = _f32(X), _f32(C)
    n, d = X.shape
    B = np.zeros((n, m), dtype=np.uint8)
    fn = lib().orc_quantize_chainq
    fn.argtypes = [_f32p, _f32p, ct.c_int64, ct.c_int, ct.c_int, ct.c_int, _u8p, VITERBI_FN]
    vit = ct.cast(ref("encode_icm").viterbi_encoding, VITERBI_FN) if use_ref else ct.cast(None, VITERBI_FN)
    fn(_p(X, _f32p), _p(C, _f32p), n, d, m, h, _p(B, _u8p), vit)
    return B
