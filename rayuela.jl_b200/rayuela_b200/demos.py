"""demos/demos_train_query_base.jl `run_demos` (:9-105) replayed end to end on the GPU library: the acceptance script
of the drop-in.  Every ENCODE and SEARCH step below goes through librayuela_b200.so via julia_api (quantize_pq /
quantize_opq / quantize_chainq / encoding_icm / encode_icm_cuda / quantize_norms / linscan_* / qerror*).

The reference's *trainers* for the initialisers -- train_pq (src/PQ.jl:68-99), train_opq (src/OPQ.jl:49-139),
train_chainq (src/ChainQ.jl:373-431) -- are outside the hot-path scope; the
`harness_*` functions here are small numpy stand-ins with the same signatures and return tuples (Lloyd iterations,
Procrustes rotation, codebook update + Viterbi encode), good enough to initialise LSQ exactly where the demo does.
RVQ / ERVQ (`experiment_rvq`, `experiment_ervq`, :39-48) are out of scope and skipped.

Data: read_dataset / load_experiment_data (src/read_datasets.jl:63-85, demos/experiment_utils.jl:63-86) read the
real SIFT1M files when a data directory holds them; otherwise `synthetic_sift` makes a SIFT-like set with exact
brute-force ground truth.
"""
import os

import numpy as np

from . import julia_api as J
from .xvecs import fvecs_read, ivecs_read

H = 256


# ---- datasets ------------------------------------------------------------------------------------
_SIFT_FILES = {"SIFT1M": "sift/sift_learn.fvecs", "SIFT1M_base": "sift/sift_base.fvecs",
               "SIFT1M_query": "sift/sift_query.fvecs", "SIFT1M_groundtruth": "sift/sift_groundtruth.ivecs"}


def read_dataset(dname, nvectors, V=False, data_dir="./data"):
    """read_dataset(dname, nvectors, V) (src/read_datasets.jl:5-244), the SIFT1M entries (:63-85): the first
    `nvectors` vectors of ./data/sift/sift_{learn,base,query}.fvecs / sift_groundtruth.ivecs, d-by-n."""
    if dname not in _SIFT_FILES:
        raise J.RayuelaError("dataset unknown: %s (only the SIFT1M files are wired up)" % dname)
    fname = os.path.join(data_dir, _SIFT_FILES[dname])
    if V:
        print("Loading %s from %s" % (dname, fname))
    bounds = tuple(int(v) for v in nvectors) if isinstance(nvectors, (tuple, list)) else int(nvectors)
    return (ivecs_read if fname.endswith(".ivecs") else fvecs_read)(bounds, fname)


def have_sift1m(data_dir="./data"):
    return all(os.path.exists(os.path.join(data_dir, f)) for f in _SIFT_FILES.values())


def load_experiment_data(dataset_name, ntrain, nbase, nquery, V=False, data_dir="./data"):
    """load_experiment_data (demos/experiment_utils.jl:63-86) -> Xt, Xb, Xq, gt (one-based ids of the true NN)."""
    Xt = read_dataset(dataset_name, ntrain, V, data_dir)
    Xb = read_dataset(dataset_name + "_base", nbase, V, data_dir)
    Xq = read_dataset(dataset_name + "_query", nquery, V, data_dir)[:, :nquery]
    gt = read_dataset(dataset_name + "_groundtruth", nquery, V, data_dir)
    if dataset_name in ("SIFT1M", "GIST1M"):
        gt = gt + 1                                      # zero-based in the files, :73-76
    gt = np.asarray(gt[0, :nquery], dtype=np.uint32)    # keep only the top neighbour, :79-80
    if nbase < 1_000_000:                                # a truncated base invalidates the file's ground truth
        gt = exact_ground_truth(Xb, Xq)
    return Xt, Xb, Xq, gt


def exact_ground_truth(Xb, Xq, block=4096):
    """One-based id of the exact fp32 nearest neighbour of every query (first minimum)."""
    Xb = np.asarray(Xb, dtype=np.float32)
    Xq = np.asarray(Xq, dtype=np.float32)
    bn = (Xb.astype(np.float64) ** 2).sum(0)
    out = np.empty(Xq.shape[1], dtype=np.uint32)
    for a in range(0, Xq.shape[1], block):
        q = Xq[:, a:a + block].astype(np.float64)
        d2 = bn[:, None] - 2.0 * (Xb.T.astype(np.float64) @ q)
        out[a:a + block] = d2.argmin(0) + 1
    return out


def synthetic_sift(ntrain, nbase, nquery, d=128, seed=0):
    """SIFT-like synthetic set (non-negative, integer valued, clustered, decaying spectrum): Xt, Xb, Xq, gt."""
    r = np.random.default_rng(seed)
    s = np.arange(1, d + 1, dtype=np.float32) ** -0.6
    rot, _ = np.linalg.qr(r.standard_normal((d, d)))
    centres = r.standard_normal((256, d)).astype(np.float32) * s

    def draw(k):
        z = centres[r.integers(0, 256, k)] + r.standard_normal((k, d)).astype(np.float32) * s
        x = (z @ rot.astype(np.float32)) * 25.0 + 40.0
        return np.asfortranarray(np.clip(np.rint(x), 0, 218).astype(np.float32).T)
    Xt, Xb, Xq = draw(ntrain), draw(nbase), draw(nquery)
    return Xt, Xb, Xq, exact_ground_truth(Xb, Xq)


# ---- harness stand-ins for the out-of-scope trainers -----------------------------------------------------------------
def _subdims(d, m):
    per, xtra = divmod(d, m)
    out, at = [], 0
    for i in range(m):
        ln = per + (1 if i < xtra else 0)
        out.append(slice(at, at + ln))
        at += ln
    return out


def _lloyd_update(Xs, b, C):
    """update_centers!: mean of the members; empty clusters keep their centre."""
    h = C.shape[1]
    sums = np.zeros((h, Xs.shape[0]), dtype=np.float64)
    np.add.at(sums, b, Xs.T.astype(np.float64))
    cnt = np.bincount(b, minlength=h)
    out = C.copy()
    nz = cnt > 0
    out[:, nz] = (sums[nz] / cnt[nz, None]).T.astype(np.float32)
    return out


def _pq_assign(X, C, sd):
    """quantize_pq for possibly ragged subspaces (d % m != 0, e.g. the demo's OPQ with m-1 = 7 codebooks over 128
    dimensions, splitarray src/utils.jl:179-203): every subspace is zero-padded to the widest one -- the extra
    +0*0 terms leave each fp32 distance unchanged -- so the library's equal-width kernel applies."""
    d = X.shape[0]
    m = len(C)
    if d % m == 0:
        return J.quantize_pq(X, C)
    w = max(s.stop - s.start for s in sd)
    Xp = np.zeros((m * w, X.shape[1]), dtype=np.float32)
    Cp = []
    for i, s in enumerate(sd):
        Xp[i * w:i * w + (s.stop - s.start)] = X[s]
        c = np.zeros((w, C[i].shape[1]), dtype=np.float32)
        c[:s.stop - s.start] = C[i]
        Cp.append(c)
    return J.quantize_pq(Xp, Cp)


def harness_train_pq(X, m, h, niter=25, V=False, rng=None):
    """train_pq(X, m, h, niter, V) -> C, B, error (src/PQ.jl:68-99): per-subspace Lloyd, random-sample seeding;
    the assignment step is the library's quantize_pq."""
    rng = rng or np.random.default_rng(1)
    X = np.asarray(X, dtype=np.float32)
    d, n = X.shape
    sd = _subdims(d, m)
    perm = rng.choice(n, h, replace=False)
    C = [np.asfortranarray(X[s][:, perm]) for s in sd]
    B = _pq_assign(X, C, sd)
    for _ in range(niter):
        C = [np.asfortranarray(_lloyd_update(X[s], B[i].astype(np.int64) - 1, C[i])) for i, s in enumerate(sd)]
        B = _pq_assign(X, C, sd)
    return C, B, float(J.qerror_pq(X, B, C))


def harness_train_opq(X, m, h, niter, init="natural", V=False, rng=None):
    """train_opq(X, m, h, niter, init, V) -> C, B, R, obj (src/OPQ.jl:49-139): alternate the Procrustes rotation
    (svd(X*CB'), :113-114) with one Lloyd step per subspace on R'X; assignments by the library's quantize_pq."""
    rng = rng or np.random.default_rng(2)
    X = np.asarray(X, dtype=np.float32)
    d, n = X.shape
    sd = _subdims(d, m)
    if init == "natural":
        R = np.eye(d, dtype=np.float32)
    elif init == "random":
        R = np.linalg.svd(rng.standard_normal((d, d)))[0].astype(np.float32)
    else:
        raise J.RayuelaError("Intialization %s unknown" % init)
    RX = R.T @ X
    perm = rng.choice(n, h, replace=False)
    C = [np.asfortranarray(RX[s][:, perm]) for s in sd]
    B = _pq_assign(RX, C, sd)
    obj = np.zeros(niter + 1, dtype=np.float32)
    for it in range(niter + 1):
        CB = np.concatenate([C[i][:, B[i].astype(np.int64) - 1] for i in range(m)], axis=0)
        obj[it] = float(((R @ CB - X) ** 2).sum() / n)
        if V:
            print("%3d %e" % (it, obj[it]))
        U, _, Vt = np.linalg.svd(X.astype(np.float64) @ CB.T.astype(np.float64), full_matrices=False)
        R = (U @ Vt).astype(np.float32)
        RX = R.T @ X
        C = [np.asfortranarray(_lloyd_update(RX[s], B[i].astype(np.int64) - 1, C[i])) for i, s in enumerate(sd)]
        B = _pq_assign(RX, C, sd)
    return C, B, R, obj


def harness_train_chainq(X, m, h, R, B, C, niter, V=False):
    """train_chainq(X, m, h, R, B, C, niter, V) -> C, B, R, obj (src/ChainQ.jl:373-431): alternate a codebook update
    with the exact chain (Viterbi) encoder.  C arrives as PQ-style blocks (sub-by-h) from OPQ and is expanded to full
    d-by-h codebooks by the first update.  The update here is the generic least-squares update_codebooks_fast_bin
    (the reference uses the chain-structured variant, src/codebook_update.jl:280-412); the encoder is the library's."""
    X = np.asarray(X, dtype=np.float32)
    R = np.asarray(R, dtype=np.float32)
    RX = np.ascontiguousarray(R.T @ X)
    obj = np.zeros(niter, dtype=np.float32)
    B = np.asfortranarray(np.asarray(B, dtype=np.int16))
    for it in range(niter):
        C = J.update_codebooks(RX, B, h, V, "fastbin")
        B, _ = J.quantize_chainq(RX, C)
        obj[it] = J.qerror(RX, B, C)
        if V:
            print("%3d %e" % (it + 1, obj[it]))
    C = [R @ c for c in C]
    return C, B, R, obj


# ---- experiments (src/PQ.jl:103-131, src/OPQ.jl:141-178) ------------------------------------------------------------------
def experiment_pq(Xt, Xb, Xq, gt, m, h, niter=25, knn=1000, V=False):
    """experiment_pq -> C, B, train_error, B_base, recall (src/PQ.jl:103-131)."""
    C, B, train_error = harness_train_pq(Xt, m, h, niter, V)
    B_base = J.quantize_pq(Xb, C, V)
    base_error = J.qerror_pq(Xb, B_base, C)
    if V:
        print("Error in training is %e\nError in base is %e" % (train_error, base_error))
    b = int(np.log2(h) * m)
    dists, idx = J.linscan_pq(B_base, np.asarray(Xq, dtype=np.float32), C, b, knn)
    return C, B, train_error, B_base, J.eval_recall(gt, idx, knn, V)


def experiment_opq(Xt, Xb, Xq, gt, m, h, init, niter=25, knn=1000, V=False):
    """experiment_opq -> C, B, R, train_error, B_base, recall (src/OPQ.jl:141-178)."""
    C, B, R, train_error = harness_train_opq(Xt, m, h, niter, init, V)
    B_base = J.quantize_opq(Xb, R, C, V)
    base_error = J.qerror_opq(Xb, B_base, C, R)
    if V:
        print("Error in base is %e" % base_error)
    b = int(np.log2(h) * m)
    dists, idx = J.linscan_opq(B_base, np.asarray(Xq, dtype=np.float32), C, b, R, knn)
    return C, B, R, train_error, B_base, J.eval_recall(gt, idx, knn, V)


def run_demos(dataset_name="SIFT1M", ntrain=int(1e5), m=8, h=256, niter=25, nquery=int(1e4), nbase=int(1e6),
              knn=int(1e3), data_dir="./data", verbose=True, seed=0, ilsiter=8, icmiter=4, randord=True, npert=4,
              sr_methods=("SR_D", "SR_C")):
    """run_demos (demos/demos_train_query_base.jl:9-105): PQ -> OPQ -> (OPQ m-1 -> ChainQ init) -> LSQ -> LSQ++ (SR-D,
    SR-C), each trained on Xt, base encoded, norms quantised, searched, recall evaluated.  Returns a dict of results
    (the reference writes HDF5 files, :31,36,71-74; there is no h5py here).  dataset_name "synthetic" (or missing
    SIFT1M files) uses synthetic_sift."""
    J.seed_b200(seed)
    if dataset_name != "synthetic" and have_sift1m(data_dir):
        Xt, Xb, Xq, gt = load_experiment_data(dataset_name, ntrain, nbase, nquery, verbose, data_dir)
    else:
        Xt, Xb, Xq, gt = synthetic_sift(ntrain, nbase, nquery, seed=seed)
    out = {"gt": gt}
    # (semi-)orthogonal methods, :29-37
    C, B, err, B_base, recall = experiment_pq(Xt, Xb, Xq, gt, m, h, niter, knn, verbose)
    out["pq"] = dict(train_error=err, recall=recall)
    C, B, R, obj, B_base, recall = experiment_opq(Xt, Xb, Xq, gt, m, h, "natural", niter, knn, verbose)
    out["opq"] = dict(train_error=float(obj[-1]), recall=recall)
    # init for LSQ / SR: OPQ with m-1 codebooks, then ChainQ, :51-59
    C, B, R, obj = harness_train_opq(Xt, m - 1, h, niter, "natural", verbose)
    C, B, R, chainq_error = harness_train_chainq(Xt, m - 1, h, R, B, C, niter, verbose)
    out["chainq"] = dict(train_error=float(chainq_error[-1]))
    nsplits_train, nsplits_base = 1, (2 if m <= 8 else 4)                                          # :61-62
    # GPU LSQ, :70-75
    Cl, Bl, Rl, train_error, B_base, recall = J.experiment_lsq_cuda(
        Xt, B, C, R, Xb, Xq, gt, m - 1, h, niter, ilsiter, icmiter, randord, npert, knn, nsplits_train, nsplits_base,
        verbose)
    out["lsq"] = dict(train_error=float(train_error[-1]), recall=recall, B_base=B_base)
    # GPU LSQ++, SR-D and SR-C, :80-95
    for sr_method in sr_methods:
        Cs, Bs, Rs, train_error, B_base, recall = J.experiment_sr_cuda(
            Xt, B, C, R, Xb, Xq, gt, m - 1, h, niter, ilsiter, icmiter, randord, npert, knn, nsplits_train,
            nsplits_base, sr_method, 1, 0.5, verbose)
        out[sr_method.lower()] = dict(train_error=float(train_error[-1]), recall=recall, B_base=B_base)
    return out


def high_recall_experiments(C, B, Xb, Xq, gt, m, h=256, ilsiters=(1, 2, 4, 8, 16, 32, 64, 128, 256), icmiter=4,
                            npert=4, randord=True, knn=1000, nsplits_base=2, V=True):
    """The per-trial body of high_recall_experiments (demos/demos_train_query_base.jl:107-158): given trained codebooks
    C and training codes B (the demo loads them from its HDF5 results, :127), encode the base ONCE with
    encode_icm_cuda's `ilsiters` snapshots (:134), then for every snapshot: qerror, quantize_norms, linscan_lsq,
    eval_recall (:136-152).  Returns {ilsiter: dict(base_error, recall)}."""
    Xb = np.asarray(Xb, dtype=np.float32)
    Xq = np.asarray(Xq, dtype=np.float32)
    d = Xb.shape[0]
    norms_B, norms_C = J.get_norms_codebook(B, C)                                                  # :128
    B_base = J._rand_codes(h, m, Xb.shape[1])                                                      # :131
    Bs_base, _ = J.encode_icm_cuda(Xb, B_base, C, list(ilsiters), icmiter, npert, randord, nsplits_base, V)
    out = {}
    for idx, ilsiter in enumerate(ilsiters):
        B_base = Bs_base[idx]
        base_error = float(J.qerror(Xb, B_base, C))
        if V:
            print("Error in base is %e" % base_error)
        B_base_norms, _ = J.quantize_norms(B_base, C, norms_C)
        db_norms = np.asarray(norms_C, dtype=np.float32)[np.asarray(B_base_norms, dtype=np.int64) - 1]
        dists, ids = J.linscan_lsq(B_base, Xq, C, db_norms, np.eye(d, dtype=np.float32), knn)
        out[int(ilsiter)] = dict(base_error=base_error, recall=J.eval_recall(gt, ids, knn, V))
    return out
