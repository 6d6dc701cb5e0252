"""Memory-image level API over librayuela_b200.so.

Arrays are either numpy (host; uploaded/downloaded inside the call) or torch CUDA tensors (device resident;
the call is enqueued on torch's current stream and nothing crosses PCIe).  Shapes are the C views of the
reference's column-major Julia arrays:
    X (n, d) f32      C (m*256, d) f32      B (n, m) u8 0-based      queries (nq, d) f32
    Cpq (m*256, d/m) f32 (PQ codebooks)     dists / idx (nq, k)
"""
import ctypes as ct

import numpy as np

from . import _lib
from ._lib import DEVICE_PTRS, FAST_LUT, FAST_UNARIES, SCAN_CQ, SCAN_LSQ, SCAN_PQ, RayuelaError, check

H = 256

try:  # torch is plumbing (device memory, streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP2T = {}
if torch is not None:
    _NP2T = {np.float32: torch.float32, np.uint8: torch.uint8, np.int32: torch.int32}


def _is_dev(a):
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


class _Args:
    """Collects array arguments of one call; all must live on the same side (host or one device)."""

    def __init__(self):
        self.dev = None
        self.keep = []

    def _side(self, dev):
        if self.dev is None:
            self.dev = dev
        elif self.dev != dev:
            raise RayuelaError("mixing host (numpy) and device (torch.cuda) arrays in one call")

    def inp(self, a, dtype, shape=None):
        if a is None:
            return None
        if _is_dev(a):
            self._side(True)
            if a.dtype != _NP2T[dtype] or not a.is_contiguous():
                a = a.to(_NP2T[dtype]).contiguous()
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise RayuelaError("bad shape %s, expected %s" % (tuple(a.shape), tuple(shape)))
            self.keep.append(a)
            return a.data_ptr()
        if torch is not None and isinstance(a, torch.Tensor):
            a = a.numpy()
        self._side(False)
        a = np.ascontiguousarray(a, dtype=dtype)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise RayuelaError("bad shape %s, expected %s" % (tuple(a.shape), tuple(shape)))
        self.keep.append(a)
        return a.ctypes.data

    def out(self, a, dtype, shape):
        """In/out or output array: must already be contiguous and of the right dtype (written in place)."""
        if _is_dev(a):
            self._side(True)
            ok = a.dtype == _NP2T[dtype] and a.is_contiguous() and tuple(a.shape) == tuple(shape)
            ptr = a.data_ptr()
        else:
            self._side(False)
            ok = isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and \
                a.flags.writeable and tuple(a.shape) == tuple(shape)
            ptr = a.ctypes.data if ok else None
        if not ok:
            raise RayuelaError("output array must be contiguous %s of shape %s" % (np.dtype(dtype), tuple(shape)))
        self.keep.append(a)
        return ptr

    def new(self, like_dev, dtype, shape, device=None):
        if like_dev:
            a = torch.empty(shape, dtype=_NP2T[dtype], device=device)
        else:
            a = np.empty(shape, dtype=dtype)
        return a, self.out(a, dtype, shape)

    @property
    def flags(self):
        return DEVICE_PTRS if self.dev else 0

    @property
    def stream(self):
        if self.dev:
            return ct.c_void_p(torch.cuda.current_stream().cuda_stream)
        return None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def encode_icm(X, C, B, ilsiter, icmiter, npert, randord, seed=0, g0=0, orders=None, snap_iters=None,
               want_cost=False, want_stats=False, inplace=False, h=H, fast=False):
    """encode_icm_fully! (src/LSQ.jl:152-252) on the GPU.  Returns dict(B, cost, stats, B_snap, objs).
    h = 256: the tuned kernels; h < 256: the reference's any-h path (iterated_conditional_modes!, src/LSQ.jl:83-149).
    fast=True: opt-in tensor-core (tcgen05 bf16x3) unaries -- no longer bit-identical to the oracle."""
    L = _lib.lib()
    dev = _is_dev(X)
    n, d = X.shape
    m = B.shape[1]
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    cp = a.inp(C, np.float32, (m * h, d))
    if not inplace:
        B = B.clone() if _is_dev(B) else np.array(B, dtype=np.uint8, order="C", copy=True)
    bp = a.out(B, np.uint8, (n, m))
    ordp = None
    if orders is not None:
        orders = _i32(orders)
        if orders.shape != (ilsiter, m):
            raise RayuelaError("orders must be ilsiter-by-m")
        ordp = orders.ctypes.data
    snaps = _i32(snap_iters if snap_iters is not None else [])
    ns = int(snaps.size)
    Bs = Bsp = objs = None
    if ns:
        Bs, Bsp = a.new(dev, np.uint8, (ns, n, m), device=X.device if dev else None)
        objs = np.zeros(ns, dtype=np.float32)
    cost = costp = None
    if want_cost:
        cost, costp = a.new(dev, np.float32, (n,), device=X.device if dev else None)
    stats = np.zeros((max(ilsiter, 1), 2), dtype=np.int32) if want_stats else None
    check(L.rayuela_encode_icm(xp, cp, bp, n, d, m, h, ilsiter, icmiter, npert, int(bool(randord)), seed, g0, ordp,
                               snaps.ctypes.data if ns else None, ns, Bsp, objs.ctypes.data if ns else None, costp,
                               stats.ctypes.data if want_stats else None, a.flags | (FAST_UNARIES if fast else 0),
                               a.stream))
    return dict(B=B, cost=cost, stats=stats[:ilsiter] if want_stats else None, B_snap=Bs, objs=objs)


def get_unaries(X, C, m, fast=False):
    """get_unaries (src/utils.jl:121-149) on the GPU: U (n, m*256) with U[l, j*256 + c] = unaries[j][c, l]."""
    L = _lib.lib()
    n, d = X.shape
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    cp = a.inp(C, np.float32, (m * H, d))
    U, up = a.new(_is_dev(X), np.float32, (n, m * H), device=X.device if _is_dev(X) else None)
    check(L.rayuela_get_unaries(xp, cp, n, d, m, H, up, a.flags | (FAST_UNARIES if fast else 0), a.stream))
    return U


def last_icm_steps():
    """(executed, total) conditioning steps of the last encode_icm(..., want_stats=True) call."""
    a, b = ct.c_uint64(0), ct.c_uint64(0)
    check(_lib.lib().rayuela_encode_icm_steps(ct.addressof(a), ct.addressof(b)))
    return int(a.value), int(b.value)


def last_icm_timings():
    """CUDA-event times (ms) of the last encode_icm(..., want_stats=True): dict(setup, unaries, icm, total)."""
    t = (ct.c_float * 4)()
    check(_lib.lib().rayuela_encode_icm_timings(ct.addressof(t)))
    return dict(setup=float(t[0]), unaries=float(t[1]), icm=float(t[2]), total=float(t[3]))


def last_icm_exact_steps():
    """Executed steps of that call that re-read the exact fp32 rows (the quantised pre-filter left a near-tie)."""
    a = ct.c_uint64(0)
    check(_lib.lib().rayuela_encode_icm_exact_steps(ct.addressof(a)))
    return int(a.value)


def veccost(X, B, C, want_mean=False, h=H):
    """veccost / qerror (src/qerrors.jl:36-74)."""
    L = _lib.lib()
    n, d = X.shape
    m = B.shape[1]
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    bp = a.inp(B, np.uint8, (n, m))
    cp = a.inp(C, np.float32, (m * h, d))
    cost, costp = a.new(_is_dev(X), np.float32, (n,), device=X.device if _is_dev(X) else None)
    mean = ct.c_double(0.0)
    check(L.rayuela_veccost(xp, bp, cp, n, d, m, h, costp, ct.addressof(mean) if want_mean else None, a.flags,
                            a.stream))
    return (cost, mean.value) if want_mean else cost


def qerror(X, B, C, h=H):
    return veccost(X, B, C, want_mean=True, h=h)[1]


def quantize_pq(X, Cpq, m):
    """quantize_pq (src/PQ.jl:18-48): nearest centroid per subspace.  Returns B (n, m) u8."""
    L = _lib.lib()
    n, d = X.shape
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    cp = a.inp(Cpq, np.float32, (m * H, d // m))
    B, bp = a.new(_is_dev(X), np.uint8, (n, m), device=X.device if _is_dev(X) else None)
    check(L.rayuela_quantize_pq(xp, cp, n, d, m, H, bp, a.flags, a.stream))
    return B


class Index:
    """Encoded base set resident on the GPU (rayuela_index_*): upload / re-layout once, scan many times."""

    def __init__(self, kind, codes, dbnorms=None, id_offset=0, h=H):
        """h: entries per codebook (1..256; the codes are bytes either way)."""
        L = _lib.lib()
        n, m = codes.shape
        a = _Args()
        cp = a.inp(codes, np.uint8, (n, m))
        np_ = a.inp(dbnorms, np.float32, (n,)) if dbnorms is not None else None
        hd = ct.c_void_p()
        check(L.rayuela_index_create(ct.byref(hd), kind, cp, np_, n, m, h, id_offset, a.flags, a.stream))
        self._h, self.kind, self.n, self.m, self.id_offset, self.h = hd, kind, n, m, id_offset, h

    def search(self, queries, codebooks, k, out=None, fast=False):
        """Top-k of every query.  fast=True (LSQ only): opt-in tensor-core LUT build, not bit-identical.  out=(dists, idx): preallocated (nq, k) float32 / int32 result arrays on the same
        side as the queries (e.g. pinned host memory), like the caller-allocated outputs of src/Linscan.jl:132-133."""
        L = _lib.lib()
        if self._h is None:
            raise RayuelaError("index already freed")
        nq, d = queries.shape
        a = _Args()
        qp = a.inp(queries, np.float32, (nq, d))
        cols = d // self.m if self.kind == SCAN_PQ else d
        cp = a.inp(codebooks, np.float32, (self.m * self.h, cols))
        dev = _is_dev(queries)
        if out is not None:
            dists, idx = out
            dp, ip = a.out(dists, np.float32, (nq, k)), a.out(idx, np.int32, (nq, k))
        else:
            dists, dp = a.new(dev, np.float32, (nq, k), device=queries.device if dev else None)
            idx, ip = a.new(dev, np.int32, (nq, k), device=queries.device if dev else None)
        check(L.rayuela_index_search(self._h, qp, cp, nq, d, k, dp, ip, a.flags | (FAST_LUT if fast else 0), a.stream))
        return dists, idx

    def free(self):
        if self._h is not None:
            _lib.lib().rayuela_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def topk_merge(dists, idx):
    """Merge S per-shard sorted lists [S, nq, k] into the global top-k [nq, k] by (dist, id)."""
    L = _lib.lib()
    S, nq, k = dists.shape
    a = _Args()
    dp = a.inp(dists, np.float32, (S, nq, k))
    ip = a.inp(idx, np.int32, (S, nq, k))
    dev = _is_dev(dists)
    do, dop = a.new(dev, np.float32, (nq, k), device=dists.device if dev else None)
    io, iop = a.new(dev, np.int32, (nq, k), device=dists.device if dev else None)
    check(L.rayuela_topk_merge(dp, ip, S, nq, k, dop, iop, a.flags, a.stream))
    return do, io


# ---- the reference's own four C symbols, exact signatures (host numpy only) ------------------------------
def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def c_linscan_aqd_query(B, Xq, centers, k):
    """linscan_aqd_query as src/Linscan.jl:19-23 calls it.  Returns dists, res (0-based ids)."""
    B, Xq, centers = _np(B, np.uint8), _np(Xq, np.float32), _np(centers, np.float32)
    n, m = B.shape
    nq, d = Xq.shape
    dists = np.zeros((nq, k), dtype=np.float32)
    res = np.zeros((nq, k), dtype=np.uint32)
    _lib.lib().linscan_aqd_query(dists.ctypes.data, res.ctypes.data, B.ctypes.data, centers.ctypes.data,
                                 Xq.ctypes.data, n, nq, 8 * m, k, m, d, d // m)
    return dists, res


def c_linscan_aqd_query_extra_byte(B, Xq, codebooks, dbnorms, k, h=H):
    """linscan_aqd_query_extra_byte as src/Linscan.jl:135-141 calls it (1-based ids)."""
    B, Xq, codebooks, dbnorms = _np(B, np.uint8), _np(Xq, np.float32), _np(codebooks, np.float32), \
        _np(dbnorms, np.float32)
    n, m = B.shape
    nq, d = Xq.shape
    dists = np.zeros((nq, k), dtype=np.float32)
    idx = np.zeros((nq, k), dtype=np.int32)
    _lib.lib().linscan_aqd_query_extra_byte(dists.ctypes.data, idx.ctypes.data, B.ctypes.data, Xq.ctypes.data,
                                            codebooks.ctypes.data, dbnorms.ctypes.data, nq, n, m, h, d, k)
    return dists, idx


def c_linscan_aqd_cq_query_extra_byte(B, Xq, codebooks, k):
    """linscan_aqd_cq_query_extra_byte as src/Linscan.jl:173-179 calls it (1-based ids)."""
    B, Xq, codebooks = _np(B, np.uint8), _np(Xq, np.float32), _np(codebooks, np.float32)
    n, m = B.shape
    nq, d = Xq.shape
    dists = np.zeros((nq, k), dtype=np.float32)
    idx = np.zeros((nq, k), dtype=np.int32)
    _lib.lib().linscan_aqd_cq_query_extra_byte(dists.ctypes.data, idx.ctypes.data, B.ctypes.data, Xq.ctypes.data,
                                               codebooks.ctypes.data, nq, n, m, H, d, k)
    return dists, idx


def c_condition(B, ub, binaries, binaries_t, pair2idx, to_condition, j):
    """`condition` as src/LSQ.jl:71-75 calls it; B and ub are modified in place."""
    n, m = B.shape
    assert B.dtype == np.uint8 and ub.dtype == np.float32 and B.flags.c_contiguous and ub.flags.c_contiguous
    bins, bins_t = _np(binaries, np.float32), _np(binaries_t, np.float32)
    p2i, tc = _i32(pair2idx), _i32(to_condition)
    _lib.lib().condition(B.ctypes.data, ub.ctypes.data, bins.ctypes.data, bins_t.ctypes.data, p2i.ctypes.data,
                         tc.ctypes.data, j, n, m)


def fast_bin_matmul(X, B, rho=1e-4):
    """fast_bin_matmul (src/codebook_update.jl:96-171) on the GPU: A = B'B + rho*I ((m*h, m*h) float64) and
    b = B'X' returned as the C-order image (d, m*h) of the reference's (m*h)-by-d column-major matrix."""
    L = _lib.lib()
    n, d = X.shape
    m = B.shape[1]
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    bp = a.inp(B, np.uint8, (n, m))
    dev = _is_dev(X)
    if dev:
        A = torch.empty((m * H, m * H), dtype=torch.float64, device=X.device)
        bb = torch.empty((d, m * H), dtype=torch.float64, device=X.device)
        a._side(True)
        ap, bbp = A.data_ptr(), bb.data_ptr()
    else:
        A = np.empty((m * H, m * H), dtype=np.float64)
        bb = np.empty((d, m * H), dtype=np.float64)
        ap, bbp = A.ctypes.data, bb.ctypes.data
    check(L.rayuela_fast_bin_matmul(xp, bp, n, d, m, H, ct.c_double(rho), ap, bbp, a.flags, a.stream))
    return A, bb


def update_codebooks_fast_bin(X, B, rho=1e-4):
    """update_codebooks_fast_bin (src/codebook_update.jl:175-204): GPU fast_bin_matmul, then the reference's own
    dense solve -- LAPACK getrf + getrs in Float64 (scipy on host arrays, torch.linalg on device tensors) --
    and conversion to Float32.  Returns the (m*h, d) image of hcat(C...)."""
    A, bb = fast_bin_matmul(X, B, rho)
    if _is_dev(X):
        lu, piv = torch.linalg.lu_factor(A)
        return torch.linalg.lu_solve(lu, piv, bb.T.contiguous()).to(torch.float32).contiguous()
    from scipy.linalg import lu_factor, lu_solve
    return np.ascontiguousarray(lu_solve(lu_factor(A), bb.T).astype(np.float32))


def quantize_norms(B, C, cbnorms=None):
    """quantize_norms (src/utils.jl:29-59) on the GPU: (norm_codes u8 0-based or None, norms f32)."""
    L = _lib.lib()
    n, m = B.shape
    d = C.shape[1]
    a = _Args()
    bp = a.inp(B, np.uint8, (n, m))
    cp = a.inp(C, np.float32, (m * H, d))
    cbp = a.inp(cbnorms, np.float32, (H,)) if cbnorms is not None else None
    dev = _is_dev(B)
    norms, npp = a.new(dev, np.float32, (n,), device=B.device if dev else None)
    codes = cdp = None
    if cbnorms is not None:
        codes, cdp = a.new(dev, np.uint8, (n,), device=B.device if dev else None)
    check(L.rayuela_quantize_norms(bp, cp, cbp, n, d, m, H, cdp, npp, a.flags, a.stream))
    return codes, norms


def quantize_chainq(X, C, m):
    """quantize_chainq (src/ChainQ.jl:287-348) on the GPU: exact Viterbi over the chain.  Returns B (n, m) u8."""
    L = _lib.lib()
    n, d = X.shape
    a = _Args()
    xp = a.inp(X, np.float32, (n, d))
    cp = a.inp(C, np.float32, (m * H, d))
    B, bp = a.new(_is_dev(X), np.uint8, (n, m), device=X.device if _is_dev(X) else None)
    check(L.rayuela_quantize_chainq(xp, cp, n, d, m, H, bp, a.flags, a.stream))
    return B


def c_viterbi_encoding(unaries, binaries, m):
    """`viterbi_encoding` as src/ChainQ.jl:26-28 calls it (host arrays): unaries (n, m*256), binaries (m-1,256,256)."""
    U, bb = _np(unaries, np.float32), _np(binaries, np.float32)
    n = U.shape[0]
    B = np.zeros((n, m), dtype=np.uint8)
    _lib.lib().viterbi_encoding(B.ctypes.data, U.ctypes.data, bb.ctypes.data, n, m)
    return B
