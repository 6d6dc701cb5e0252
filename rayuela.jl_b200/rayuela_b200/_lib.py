"""ctypes binding of librayuela_b200.so (include/rayuela_b200.h).

There is no CPU fallback: if the shared library is missing, or it cannot reach a CUDA device, every call
raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C rayuela.jl_b200`.
"""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAYUELA_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "librayuela_b200.so")

DEVICE_PTRS = 1
FAST_UNARIES = 2
FAST_LUT = 4
SCAN_LSQ, SCAN_CQ, SCAN_PQ = 0, 1, 2

_vp = ct.c_void_p
_i64 = ct.c_int64
_int = ct.c_int

# name -> (restype, argtypes); also the list tests check against include/rayuela_b200.h
SIGNATURES = {
    "rayuela_last_error": (ct.c_char_p, []),
    "rayuela_set_device": (_int, [_int]),
    "rayuela_launch_count": (ct.c_uint64, []),
    "rayuela_init": (_int, [_vp, _int]),
    "rayuela_shutdown": (_int, []),
    "rayuela_device_count": (_int, []),
    "rayuela_encode_icm": (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _int, ct.c_uint64, _i64,
                                  _vp, _vp, _int, _vp, _vp, _vp, _vp, ct.c_uint, _vp]),
    "rayuela_encode_icm_steps": (_int, [_vp, _vp]),
    "rayuela_encode_icm_exact_steps": (_int, [_vp]),
    "rayuela_encode_icm_timings": (_int, [_vp]),
    "rayuela_get_unaries": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, ct.c_uint, _vp]),
    "rayuela_veccost": (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, ct.c_uint, _vp]),
    "condition": (None, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int]),
    "linscan_aqd_query": (None, [_vp, _vp, _vp, _vp, _vp, _int, ct.c_uint, _int, _int, _int, _int, _int]),
    "linscan_aqd_query_extra_byte": (None, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "linscan_aqd_cq_query_extra_byte": (None, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "rayuela_index_create": (_int, [ct.POINTER(_vp), _int, _vp, _vp, _i64, _int, _int, _i64, ct.c_uint, _vp]),
    "rayuela_index_search": (_int, [_vp, _vp, _vp, _int, _int, _int, _vp, _vp, ct.c_uint, _vp]),
    "rayuela_index_free": (_int, [_vp]),
    "rayuela_topk_merge": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp, ct.c_uint, _vp]),
    "rayuela_quantize_pq": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, ct.c_uint, _vp]),
    "rayuela_quantize_chainq": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, ct.c_uint, _vp]),
    "viterbi_encoding": (None, [_vp, _vp, _vp, _int, _int]),
    "rayuela_quantize_norms": (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, ct.c_uint, _vp]),
    "rayuela_fast_bin_matmul": (_int, [_vp, _vp, _i64, _int, _int, _int, ct.c_double, _vp, _vp, ct.c_uint, _vp]),
}

_lib = None


class RayuelaError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RayuelaError(
                "librayuela_b200.so is not built (%s). This package has no CPU fallback; run "
                "`make -C rayuela.jl_b200` (needs nvcc) first." % LIB_PATH)
        L = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().rayuela_last_error().decode("utf-8", "replace")
        raise RayuelaError("librayuela_b200 error %d: %s" % (rc, msg))


def launch_count():
    return int(lib().rayuela_launch_count())


def init(devices=None):
    """rayuela_init: configure the device set used by HOST-array calls (encode_icm splits the base over the
    devices, Index becomes base-sharded).  devices=None or a single device -> single-device mode.  A device may be
    listed more than once (two shards on one GPU)."""
    devs = list(devices or [])
    arr = (ct.c_int * max(len(devs), 1))(*devs)
    check(lib().rayuela_init(ct.cast(arr, ct.c_void_p), len(devs)))


def shutdown():
    check(lib().rayuela_shutdown())


def device_count():
    return int(lib().rayuela_device_count())
