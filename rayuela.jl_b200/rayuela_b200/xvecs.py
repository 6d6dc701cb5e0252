""".fvecs / .ivecs / .bvecs readers and writers ("next" row 4): src/xvecs_read.jl, src/xvecs_write.jl.

Format: every vector is an int32 dimension d followed by d components (float32 / int32 / uint8).
Readers return the Julia-shaped d-by-n matrix (Fortran order, so its bytes are the hot paths' (n, d) image);
`bounds` follows the reference: an int n reads vectors 1..n, a (a, b) pair reads a..b (1-based, inclusive).
"""
import os

import numpy as np


def _read(filename, bounds, dtype, itemsize):
    with open(filename, "rb") as f:
        d = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        vecsize = 4 + d * itemsize
        total = os.path.getsize(filename) // vecsize
        if bounds is None:
            a, b = 1, total
        elif isinstance(bounds, (tuple, list, range)):
            a, b = (bounds[0], bounds[-1])
        else:
            a, b = 1, int(bounds)
        assert a >= 1 and b <= total, "bounds outside the file"          # src/xvecs_read.jl:17
        n = b - a + 1
        f.seek((a - 1) * vecsize)
        raw = np.fromfile(f, dtype=np.uint8, count=n * vecsize).reshape(n, vecsize)
    dims = raw[:, :4].copy().view(np.int32).reshape(-1)
    assert np.all(dims == d), "inconsistent vector dimension in file"      # src/xvecs_read.jl:41-44
    body = np.ascontiguousarray(raw[:, 4:]).view(dtype).reshape(n, d)
    return np.asfortranarray(body.T)


def fvecs_read(bounds=None, filename=None):
    """fvecs_read(n | (a, b), filename) -> d-by-n Float32   (src/xvecs_read.jl)."""
    return _read(filename, bounds, np.float32, 4)


def ivecs_read(bounds=None, filename=None):
    """ivecs_read(n | (a, b), filename) -> d-by-n Int32."""
    return _read(filename, bounds, np.int32, 4)


def bvecs_read(bounds=None, filename=None):
    """bvecs_read(n | (a, b), filename) -> d-by-n UInt8   (src/xvecs_read.jl:14-52)."""
    return _read(filename, bounds, np.uint8, 1)


def _write(X, filename, dtype):
    X = np.asarray(X)
    d, n = X.shape
    out = np.empty((n, d + 1), dtype=dtype)
    out[:, 0] = np.array([d], dtype=np.int32).view(dtype)[0]              # reinterpret(Float32, Int32(d)), :13
    out[:, 1:] = X.T
    out.tofile(filename)


def fvecs_write(X, filename):
    """fvecs_write(X::Matrix{Float32}, filename)   (src/xvecs_write.jl:10-16)."""
    _write(np.asarray(X, dtype=np.float32), filename, np.float32)


def ivecs_write(X, filename):
    """ivecs_write(X::Matrix{Int32}, filename)   (src/xvecs_write.jl:19-25)."""
    _write(np.asarray(X, dtype=np.int32), filename, np.int32)
