"""Host-side mirror of Rayuela.jl's API for the two hot paths: same function names, positional arguments,
shapes and index bases as the Julia package, so scripts written against the reference (e.g.
demos/demos_train_query_base.jl) translate line by line and the parity tests read like its own would.

Julia conventions kept at this level:
    X   d-by-n Float32        C  list of m d-by-h Float32 matrices       R  d-by-d Float32
    B   m-by-n Int16, ONE-based codes                 results k-by-nq, ids ONE-based
A Julia column-major d-by-n array and a numpy C-order (n, d) array are the same bytes, so passing
`np.asfortranarray(X)` (or a transposed C-order array) costs no copy.  (Julia is not installed in this image;
julia/RayuelaB200.jl holds the equivalent ccall shim.)
"""
import numpy as np

from . import core
from .core import H, SCAN_CQ, SCAN_LSQ, SCAN_PQ, RayuelaError

_state = {"seed": 0}


def seed_b200(seed):
    """Reproducible stream for the ILS perturbations / visiting orders (the reference draws from Julia's
    global MersenneTwister; here every encode call consumes one seed from this counter)."""
    _state["seed"] = int(seed)


def _next_seed():
    s = _state["seed"]
    _state["seed"] = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    return s


def _img(A):
    """Julia d-by-n matrix -> its memory image as a C-order (n, d) float32 array (no copy when possible)."""
    return np.ascontiguousarray(np.asarray(A, dtype=np.float32).T)


def _hcat(C):
    """hcat(C...) of m d-by-h codebooks -> memory image (m*h, d)."""
    return np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32).T for c in C], axis=0))


def _codes0(B):
    """m-by-n one-based Int16 -> (n, m) uint8 zero-based (convert(Matrix{UInt8}, B .- 1), src/LSQ.jl:228)."""
    B = np.asarray(B)
    if B.min() < 1 or B.max() > H:
        raise RayuelaError("codes must be in 1..256")
    return np.ascontiguousarray((B.T - 1).astype(np.uint8))


def _codes1(B0):
    return np.asfortranarray(B0.T.astype(np.int16) + 1)


# ---- path (1) -----------------------------------------------------------------------------------------
def encoding_icm(X, oldB, C, ilsiter, icmiter, randord, npert, cpp=True, V=False):
    """encoding_icm(X, oldB, C, ilsiter, icmiter, randord, npert, cpp=true, V=true) -> B   (src/LSQ.jl:272-294)
    oldB is updated in place like the reference does (src/LSQ.jl:248)."""
    d, n = np.shape(X)
    m = len(C)
    h = np.shape(C[0])[1]
    if h > H:
        raise RayuelaError("The B200 implementation of ICM encoding stores codes in one byte: h must be <= 256")
    # h == 256 is the reference's cpp=true path (src/LSQ.jl:42-80, 256-only :173-175); other h its cpp=false path
    # iterated_conditional_modes! (:83-149) -- same results either way, `cpp` only selects the implementation there
    r = core.encode_icm(_img(X), _hcat(C), _codes0(oldB), ilsiter, icmiter, npert, randord, seed=_next_seed(),
                        want_stats=V, inplace=True, h=h)
    if V:
        for it, (neq, nbet) in enumerate(r["stats"]):
            print(" ILS iteration %d/%d done. %5.2f%% new codes are equal. %5.2f%% new codes are better."
                  % (it + 1, ilsiter, 100.0 * neq / n, 100.0 * nbet / n))
    B = _codes1(r["B"])
    if isinstance(oldB, np.ndarray) and oldB.shape == B.shape:
        oldB[...] = B
    return B


def encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, nsplits=2, V=False):
    """encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, nsplits=2, V=false) -> Bs, objs
    (src/LSQ_GPU.jl:218-264).  `nsplits` is accepted and ignored: the library tiles the base set itself."""
    d, n = np.shape(RX)
    ilsiters = [int(i) for i in ilsiters]
    r = core.encode_icm(_img(RX), _hcat(C), _codes0(B), max(ilsiters), icmiter, npert, randord, seed=_next_seed(),
                        snap_iters=ilsiters, want_stats=V, inplace=True)
    if V:
        for it, (neq, nbet) in enumerate(r["stats"]):
            print(" ILS iteration %d/%d done. %5.2f%% new codes are equal. %5.2f%% new codes are better."
                  % (it + 1, max(ilsiters), 100.0 * neq / n, 100.0 * nbet / n))
    return [_codes1(b) for b in r["B_snap"]], r["objs"].copy()


def veccost(X, B, C):
    """veccost(X, B, C) (src/qerrors.jl:36-66)."""
    return core.veccost(_img(X), _codes0(B), _hcat(C), h=np.shape(C[0])[1])


def qerror(X, B, C):
    """qerror(X, B, C) = mean(veccost(X, B, C)) (src/qerrors.jl:69-74)."""
    return np.float32(core.qerror(_img(X), _codes0(B), _hcat(C), h=np.shape(C[0])[1]))


def _pq_as_full(C, d):
    """PQ codebooks (m blocks of sub-by-h) embedded as d-by-h additive codebooks (zeros off-block)."""
    m = len(C)
    out, at = [], 0
    for c in C:
        c = np.asarray(c, dtype=np.float32)
        f = np.zeros((d, c.shape[1]), dtype=np.float32)
        f[at:at + c.shape[0]] = c
        at += c.shape[0]
        out.append(f)
    return out


def qerror_opq(X, B, C, R):
    """qerror_opq(X, B, C, R) = mean ||R*CB - X||^2 (src/qerrors.jl:77-90) = qerror of R'X against the
    block codebooks (R orthogonal)."""
    X = np.asarray(X, dtype=np.float32)
    RX = np.asarray(R, dtype=np.float32).T @ X
    return qerror(RX, B, _pq_as_full(C, X.shape[0]))


def qerror_pq(X, B, C):
    """qerror_pq(X, B, C) (src/qerrors.jl:93-100)."""
    return qerror(X, B, _pq_as_full(C, np.shape(X)[0]))


# ---- PQ / OPQ encode ------------------------------------------------------------------------------------
def _cat3(C):
    """cat(C..., dims=3) of m sub-by-h matrices -> memory image (m*h, sub)."""
    return np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32).T for c in C], axis=0))


def quantize_pq(X, C, V=False):
    """quantize_pq(X, C, V=false) -> B (m-by-n Int16, one-based)   (src/PQ.jl:18-48)."""
    d, n = np.shape(X)
    m = len(C)
    if any(np.shape(c) != (d // m, H) for c in C) or d % m:
        raise RayuelaError("quantize_pq needs m codebooks of size (d/m)-by-256")
    return _codes1(core.quantize_pq(_img(X), _cat3(C), m))


def quantize_opq(X, R, C, V=False):
    """quantize_opq(X, R, C, V=false) = quantize_pq(R'X, C)   (src/OPQ.jl:19-27)."""
    RX = np.asarray(R, dtype=np.float32).T @ np.asarray(X, dtype=np.float32)
    return quantize_pq(RX, C, V)


# ---- path (2) -----------------------------------------------------------------------------------------
def _scan_codes(B):
    B = np.asarray(B)
    if B.dtype == np.uint8:      # the UInt8 methods take zero-based codes (src/Linscan.jl:5,93,118,160)
        return np.ascontiguousarray(B.T)
    return _codes0(B)            # the Integer methods subtract one (src/Linscan.jl:35,113,155,191)


def _out(dists, idx):
    return np.asfortranarray(dists.T), np.asfortranarray(idx.T.astype(np.uint32))


def linscan_pq(B, X, C, b, k=10000):
    """linscan_pq(B, X, C, b, k) -> dists, res (k-by-nq; res one-based)   (src/Linscan.jl:5-37)."""
    codes = _scan_codes(B)
    m = codes.shape[1]
    if b != 8 * m:
        raise RayuelaError("b must be log2(h)*m = 8*m")
    ix = core.Index(SCAN_PQ, codes)
    dists, idx = ix.search(_img(X), _cat3(C), k)
    ix.free()
    return _out(dists, idx + 1)          # res .+= 1, src/Linscan.jl:25


def linscan_opq(B, X, C, b, R, k=10000):
    """linscan_opq(B, X, C, b, R, k) = linscan_pq(B, R'X, C, b, k)   (src/Linscan.jl:93-115)."""
    RX = np.asarray(R, dtype=np.float32).T @ np.asarray(X, dtype=np.float32)
    return linscan_pq(B, RX, C, b, k)


def linscan_lsq(B, X, C, dbnorms, R, k=10000):
    """linscan_lsq(B, X, C, dbnorms, R, k) -> dists, res (one-based)   (src/Linscan.jl:118-157)."""
    RX = np.asarray(R, dtype=np.float32).T @ np.asarray(X, dtype=np.float32)
    ix = core.Index(SCAN_LSQ, _scan_codes(B), np.asarray(dbnorms, dtype=np.float32).reshape(-1), h=np.shape(C[0])[1])
    dists, idx = ix.search(_img(RX), _hcat(C), k)
    ix.free()
    return _out(dists, idx)


def linscan_cq(B, X, C, k=10000):
    """linscan_cq(B, X, C, k) -> dists, res (one-based)   (src/Linscan.jl:160-193)."""
    ix = core.Index(SCAN_CQ, _scan_codes(B), h=np.shape(C[0])[1])
    dists, idx = ix.search(_img(X), _hcat(C), k)
    ix.free()
    return _out(dists, idx)


def eval_recall(ids_gnd, ids_predicted, k, V=True):
    """eval_recall(ids_gnd, ids_predicted, k) -> recall_at_i (k-vector)   (src/Linscan.jl:196-234).
    A query counts as found at rank r only if the true id appears exactly once in its list (:208-214)."""
    gt = np.asarray(ids_gnd).reshape(-1)
    P = np.asarray(ids_predicted)
    nq = P.shape[1]
    assert nq == gt.size
    hit = P[:k, :] == gt[None, :]
    cnt = hit.sum(0)
    ranks = np.where(cnt == 1, hit.argmax(0) + 1, k + 1)
    recall = np.array([(ranks <= i).sum() / nq for i in range(1, k + 1)])
    if V:
        for i in (1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000):
            if i <= k:
                print("r@%d = %s" % (i, recall[i - 1] * 100))
    return recall


# ---- LSQ++ stochastic relaxations (host side, between encodes; src/SR_perturbations.jl) -------------------
def apply_schedule(stdev, it, niter, schedule=1, p=0.5):
    """apply_schedule (src/SR_perturbations.jl:4-24)."""
    stdev = np.asarray(stdev, dtype=np.float64)
    if schedule == 1:
        return stdev * (1 - it / niter) ** p
    if schedule == 2:
        return stdev / ((1 + it) ** p)
    if schedule == 3:
        return stdev * p ** (it / 2)
    raise RayuelaError("Schedule unknown: %s" % schedule)


def SR_D_perturb(C, it, niter, schedule=1, p=0.5, rng=None):
    """SR_D_perturb (src/SR_perturbations.jl:27-49): C[i][j,:] += randn(h) * std_j / m, scheduled."""
    rng = rng or np.random.default_rng(_next_seed() & 0xFFFFFFFF)
    m = len(C)
    allc = np.concatenate([np.asarray(c, dtype=np.float32) for c in C], axis=1)
    stdc = apply_schedule(allc.std(axis=1, ddof=1) / m, it, niter, schedule, p)   # Statistics.std is corrected
    out = []
    for c in C:
        c = np.asarray(c, dtype=np.float32)
        noise = rng.standard_normal(c.shape) * stdc[:, None]
        out.append((c + noise).astype(np.float32))
    return out


def SR_C_perturb(X, it, niter, schedule=1, p=0.5, rng=None):
    """SR_C_perturb (src/SR_perturbations.jl:52-73): X[i,:] += randn(n) * std_i, scheduled."""
    rng = rng or np.random.default_rng(_next_seed() & 0xFFFFFFFF)
    X = np.asarray(X, dtype=np.float32)
    stdx = apply_schedule(X.std(axis=1, ddof=1), it, niter, schedule, p)
    return (X + rng.standard_normal(X.shape) * stdx[:, None]).astype(np.float32)


# ---- codebook update and the trainers that alternate it with the encoder ("next" row 1) ---------
def _unhcat(Cimg, m):
    return [np.asfortranarray(Cimg[i * H:(i + 1) * H].T) for i in range(m)]


def update_codebooks_fast_bin(X, B, h, V=False, rho=1e-4):
    """update_codebooks_fast_bin(X, B, h, V=false, rho=1e-4) -> C   (src/codebook_update.jl:175-204)."""
    if h != H:
        raise RayuelaError("only h = 256 is supported")
    m = np.shape(B)[0]
    return _unhcat(core.update_codebooks_fast_bin(_img(X), _codes0(B), rho), m)


def update_codebooks(X, B, h, V=False, method="fastbin"):
    """update_codebooks(X, B, h, V=false, method="fastbin")   (src/codebook_update.jl:235-278); only the method
    every LSQ / LSQ++ trainer uses ("fastbin") is provided."""
    if method != "fastbin":
        raise RayuelaError("Codebook update method not available on B200: " + method)
    return update_codebooks_fast_bin(X, B, h, V)


def train_lsq(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, cpp=True, V=True):
    """train_lsq(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, cpp=true, V=true) -> C, B, obj
    (src/LSQ.jl:323-372; train_lsq_cuda src/LSQ_GPU.jl:267-319 is the same alternation)."""
    X = np.asarray(X, dtype=np.float32)
    R = np.asarray(R, dtype=np.float32)
    RX = R.T @ X
    C = update_codebooks(RX, B, h, V, "fastbin")                     # src/LSQ.jl:343-344
    C = [R @ c for c in C]                                           # :347
    if V:
        print("%3d %e" % (-2, qerror(X, B, C)))
    B = encoding_icm(X, np.array(B, dtype=np.int16), C, ilsiter, icmiter, randord, npert, cpp, V)   # :351
    if V:
        print("%3d %e" % (-1, qerror(X, B, C)))
    obj = np.zeros(niter, dtype=np.float32)
    for it in range(niter):
        obj[it] = qerror(X, B, C)                                    # :357
        if V:
            print("%3d %e" % (it + 1, obj[it]))
        C = update_codebooks(X, B, h, V, "fastbin")                  # :361
        B = encoding_icm(X, B, C, ilsiter, icmiter, randord, npert, cpp, V)   # :364
    return C, B, obj


def train_lsq_cuda(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, nsplits=1, V=False):
    """train_lsq_cuda (src/LSQ_GPU.jl:267-319): same alternation; nsplits is accepted and ignored."""
    return train_lsq(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, True, V)


def train_sr_cuda(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, method, schedule, p=0.5, nsplits=1,
                  V=False):
    """train_sr_cuda(...) -> C, B, objarray   (src/SR.jl:88-175): LSQ++ = LSQ with SR-C (noise on X before the
    codebook update) or SR-D (noise on C before the encode)."""
    if method not in ("SR_C", "SR_D"):
        raise RayuelaError("SR method unknown")
    X = np.asarray(X, dtype=np.float32)
    R = np.asarray(R, dtype=np.float32)
    RX = np.ascontiguousarray(R.T @ X)
    if method == "SR_C":
        C = update_codebooks(SR_C_perturb(RX, 0, niter, schedule, p), B, h, V, "fastbin")        # :118-121
    else:
        C = update_codebooks(RX, B, h, V, "fastbin")                                               # :124
        C = SR_D_perturb(C, 1, niter, schedule, p)                                                 # :127
    Bs, _ = encode_icm_cuda(RX, B, C, [ilsiter], icmiter, npert, randord, nsplits, V)              # :134
    B = Bs[-1]
    objarray = np.zeros(niter + 1, dtype=np.float32)
    for it in range(1, niter + 1):
        objarray[it - 1] = qerror(RX, B, C)                                                        # :146-147
        if V:
            print("%3d %e" % (it, objarray[it - 1]))
        if method == "SR_C":
            C = update_codebooks(SR_C_perturb(RX, it, niter, schedule, p), B, h, V, "fastbin")   # :152-154
        else:
            C = update_codebooks(RX, B, h, V, "fastbin")                                           # :157
            C = SR_D_perturb(C, it, niter, schedule, p)                                            # :158
        Bs, _ = encode_icm_cuda(RX, B, C, [ilsiter], icmiter, npert, randord, nsplits, V)          # :162
        B = Bs[-1]
        C = update_codebooks(RX, B, h, V, "fastbin")                                               # :166
    objarray[niter] = qerror(RX, B, C)                                                             # :169
    C = [R @ c for c in C]                                                                         # :172
    return C, B, objarray


def _rand_codes(h, m, n):
    """convert(Matrix{Int16}, rand(1:h, m, n)) (src/LSQ_GPU.jl:351) from the seeded stream."""
    rng = np.random.default_rng(_next_seed() & 0xFFFFFFFF)
    return np.asfortranarray(rng.integers(1, h + 1, (m, n)).astype(np.int16))


def _encode_base_and_query(C, B, R, train_error, Xb, Xq, gt, m, h, ilsiter, icmiter, randord, npert, knn,
                           nsplits_base, V):
    """Shared tail of experiment_lsq_cuda / experiment_sr_cuda (src/LSQ_GPU.jl:347-367, src/SR.jl:280-305)."""
    Xb = np.asarray(Xb, dtype=np.float32)
    Xq = np.asarray(Xq, dtype=np.float32)
    d = Xb.shape[0]
    norms_B, norms_C = get_norms_codebook(B, C)                                                    # :347 / :280
    B_base = _rand_codes(h, m, Xb.shape[1])                                                        # :351 / :283
    Bs_base, _ = encode_icm_cuda(Xb, B_base, C, [ilsiter * 4], icmiter, npert, randord, nsplits_base, V)
    B_base = Bs_base[-1]
    base_error = qerror(Xb, B_base, C)
    if V:
        print("Error in base is %e" % base_error)
    B_base_norms, db_norms_X = quantize_norms(B_base, C, norms_C)                                 # :358 / :295
    db_norms = np.asarray(norms_C, dtype=np.float32)[np.asarray(B_base_norms, dtype=np.int64) - 1]  # vec(norms_C[...])
    dists, idx = linscan_lsq(B_base, Xq, C, db_norms, np.eye(d, dtype=np.float32), knn)           # :362 / :300
    recall = eval_recall(gt, idx, knn, V)
    return C, B, R, train_error, B_base, recall


def experiment_lsq_cuda(Xt, B, C, R, Xb, Xq, gt, m, h, niter=25, ilsiter=8, icmiter=4, randord=True, npert=4,
                        knn=1000, nsplits_train=1, nsplits_base=1, V=False):
    """experiment_lsq_cuda(Xt, B, C, R, Xb, Xq, gt, m, h, niter=25, ilsiter=8, icmiter=4, randord=true, npert=4,
    knn=1000, nsplits_train=1, nsplits_base=1, V=false) -> C, B, R, train_error, B_base, recall
    (src/LSQ_GPU.jl:322-368): train on Xt, encode the base from random codes with 4*ilsiter ILS iterations,
    quantize the database norms, linscan_lsq, eval_recall."""
    C, B, train_error = train_lsq_cuda(Xt, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, nsplits_train, V)
    return _encode_base_and_query(C, B, R, train_error, Xb, Xq, gt, m, h, ilsiter, icmiter, randord, npert, knn,
                                  nsplits_base, V)


def experiment_sr_cuda(Xt, B, C, R, Xb, Xq, gt, m, h, niter=25, ilsiter=8, icmiter=4, randord=True, npert=4,
                       knn=1000, nsplits_train=1, nsplits_base=1, sr_method="SR_D", schedule=1, p=0.5, V=False):
    """experiment_sr_cuda(..., sr_method="SR_D", schedule=1, p=0.5, V=false) -> C, B, R, train_error, B_base, recall
    (src/SR.jl:247-306): the LSQ++ experiment the demos run for SR-D and SR-C."""
    if V:
        print("\nRunning LSQ++ (%s) with %d codebooks, %d perturbations, %d icm iterations and random order = %s"
              % (sr_method, m, npert, icmiter, randord))
    C, B, train_error = train_sr_cuda(Xt, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, sr_method,
                                      schedule, p, nsplits_train, V)
    return _encode_base_and_query(C, B, R, train_error, Xb, Xq, gt, m, h, ilsiter, icmiter, randord, npert, knn,
                                  nsplits_base, V)


# ---- norm quantization ("next" row 2): produces the dbnorms linscan_lsq consumes ------------------------------
def kmeans_1d(x, k, rng, maxiter=100):
    """1-D k-means (k-means++ seeding, Lloyd) -- the role Clustering.kmeans(dbnorms, h) plays at
    src/utils.jl:20.  The reference's result depends on Julia's global RNG and is not reproducible; this one is
    seeded.  Returns (assignments 0-based, centers)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    n = x.size
    centers = np.empty(k)
    centers[0] = x[rng.integers(n)]
    d2 = (x - centers[0]) ** 2
    for i in range(1, k):
        tot = d2.sum()
        centers[i] = x[rng.integers(n)] if tot <= 0 else x[np.searchsorted(np.cumsum(d2), rng.random() * tot)]
        d2 = np.minimum(d2, (x - centers[i]) ** 2)
    for _ in range(maxiter):
        centers.sort()
        edges = (centers[1:] + centers[:-1]) / 2
        a = np.searchsorted(edges, x)
        sums = np.bincount(a, weights=x, minlength=k)
        cnts = np.bincount(a, minlength=k)
        new = np.where(cnts > 0, sums / np.maximum(cnts, 1), centers)
        if np.allclose(new, centers, rtol=0, atol=1e-12):
            centers = new
            break
        centers = new
    centers.sort()
    a = np.searchsorted((centers[1:] + centers[:-1]) / 2, x)
    return a, centers.astype(np.float32)


def get_norms_codebook(B, C):
    """get_norms_codebook(B, C) -> norms_codes (1-based), norms_codebook   (src/utils.jl:4-26)."""
    m = len(C)
    assert np.shape(B)[0] == m
    _, dbnorms = core.quantize_norms(_codes0(B), _hcat(C))            # reconstruct + sum of squares on the GPU
    a, centers = kmeans_1d(dbnorms, np.shape(C[0])[1], np.random.default_rng(_next_seed() & 0xFFFFFFFF))
    return (a + 1).astype(np.int16), centers


def quantize_norms(B, C, cbnorms):
    """quantize_norms(B, C, cbnorms) -> dbnormsB (1-based), dbnormsX   (src/utils.jl:29-59)."""
    codes, norms = core.quantize_norms(_codes0(B), _hcat(C), np.asarray(cbnorms, dtype=np.float32))
    return codes.astype(np.int16) + 1, norms


# ---- ChainQ Viterbi encode ("next" row 3) --------------------------------------------------------------------
def quantize_chainq(X, C, use_cuda=False, use_cpp=False):
    """quantize_chainq(X, C, use_cuda=false, use_cpp=false) -> B, ellapsed   (src/ChainQ.jl:287-348).
    The three reference implementations give identical codes (test/chainq.jl:27-39); here there is one."""
    import time
    t0 = time.perf_counter()
    B0 = core.quantize_chainq(_img(X), _hcat(C), len(C))
    return _codes1(B0), time.perf_counter() - t0
