"""One process per GPU (torch.distributed) sharding of the two hot paths.

Neither path exists in distributed form in the reference (its only split is `nsplits`, serial chunks through one
GPU, src/LSQ_GPU.jl:236-255); the partitioning below reuses its `splitarray` rule (src/utils.jl:179-203).

  * encode: vectors are independent (deps/src/encode_icm.cpp:26-59), so each rank encodes a contiguous slice.
    The perturbation RNG is keyed on the GLOBAL vector index (g0 = slice start), so the codes are bit-identical
    for every world size.  No data-path collective; `gather=True` adds one all_gather of the code slices.
  * scan, base-sharded: every rank scans its slice of the base for all queries and returns its local top-k with
    global ids (id_offset = slice start); ONE all_gather of (dists, ids) and a k-way merge by the (dist, id) total
    order give exactly the single-GPU result.
  * scan, query-sharded: queries are split, the base is replicated; results are concatenated.

`backend` carries the four primitives (encode, index, merge, empty-like); the default is the CUDA library.
Tests inject a CPU stand-in to exercise the partition / exchange logic over gloo.
"""
import numpy as np

from . import core

try:
    import torch
    import torch.distributed as tdist
except Exception:  # pragma: no cover
    torch = None
    tdist = None


def splitarray(n, nparts):
    """Reference rule (src/utils.jl:179-203): the first n % nparts parts get one extra element.
    Returns [(start, stop)] * nparts (0-based, half-open)."""
    per, xtra = divmod(n, nparts)
    out, at = [], 0
    for i in range(nparts):
        ln = per + (1 if i < xtra else 0)
        out.append((at, at + ln))
        at += ln
    return out


class CudaBackend:
    """The product backend: librayuela_b200.so on this rank's GPU."""

    def encode(self, X, C, B, ilsiter, icmiter, npert, randord, seed, g0):
        return core.encode_icm(X, C, B, ilsiter, icmiter, npert, randord, seed=seed, g0=g0)["B"]

    def make_index(self, kind, codes, norms, id_offset):
        return core.Index(kind, codes, norms, id_offset=id_offset)

    def merge(self, dists, idx):
        return core.topk_merge(dists, idx)

    def to_tensor(self, a):
        if isinstance(a, np.ndarray):
            return torch.from_numpy(a).cuda()
        return a


def _world(group):
    if tdist is None or not tdist.is_initialized():
        return 0, 1
    return tdist.get_rank(group), tdist.get_world_size(group)


def _all_gather_cat(t, sizes, group):
    """all_gather of first-dim-ragged tensors (pads to the longest slice)."""
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in sizes]
    tdist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


def sharded_encode_icm(X, C, B, ilsiter, icmiter, npert, randord, seed=0, group=None, gather=False, backend=None,
                       local_slice=None):
    """Encode this rank's slice of a base set of n vectors.

    X, B: either the FULL arrays (every rank passes the same; the rank's slice is taken here) or, with
    local_slice=(start, stop, n_total), just this rank's rows.  Returns (B_local, (start, stop)), or the full
    code matrix on every rank when gather=True."""
    backend = backend or CudaBackend()
    rank, world = _world(group)
    if local_slice is None:
        n = X.shape[0]
        a, b = splitarray(n, world)[rank]
        Xl, Bl = X[a:b], B[a:b]
    else:
        a, b, n = local_slice
        Xl, Bl = X, B
    out = backend.encode(Xl, C, Bl, ilsiter, icmiter, npert, randord, seed, a)
    if not gather or world == 1:
        return out, (a, b)
    t = backend.to_tensor(out)
    sizes = [e - s for s, e in splitarray(n, world)]
    return _all_gather_cat(t, sizes, group), (0, n)


class ShardedIndex:
    """Base-sharded index: this rank holds rows [start, stop) of the encoded base."""

    def __init__(self, kind, codes_local, norms_local, start, group=None, backend=None):
        self.backend = backend or CudaBackend()
        self.group = group
        self.rank, self.world = _world(group)
        self.index = self.backend.make_index(kind, codes_local, norms_local, start)

    def search(self, queries, codebooks, k):
        """Top-k over the WHOLE base for all queries, identical on every rank."""
        d_loc, i_loc = self.index.search(queries, codebooks, k)
        if self.world == 1:
            return d_loc, i_loc
        d_loc, i_loc = self.backend.to_tensor(d_loc), self.backend.to_tensor(i_loc)
        nq = d_loc.shape[0]
        gd = torch.empty((self.world * nq, k), dtype=d_loc.dtype, device=d_loc.device)
        gi = torch.empty((self.world * nq, k), dtype=i_loc.dtype, device=i_loc.device)
        tdist.all_gather_into_tensor(gd, d_loc.contiguous(), group=self.group)   # the one exchange step
        tdist.all_gather_into_tensor(gi, i_loc.contiguous(), group=self.group)
        return self.backend.merge(gd.view(self.world, nq, k), gi.view(self.world, nq, k))


def query_sharded_search(index, queries, codebooks, k, group=None, backend=None):
    """Replicated base, queries split by splitarray; returns the full (nq, k) result on every rank."""
    backend = backend or CudaBackend()
    rank, world = _world(group)
    nq = queries.shape[0]
    parts = splitarray(nq, world)
    a, b = parts[rank]
    d_loc, i_loc = index.search(queries[a:b], codebooks, k)
    if world == 1:
        return d_loc, i_loc
    sizes = [e - s for s, e in parts]
    return (_all_gather_cat(backend.to_tensor(d_loc), sizes, group),
            _all_gather_cat(backend.to_tensor(i_loc), sizes, group))
