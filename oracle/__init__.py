"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

ctypes binding over oracle/liboracle.so (built from oracle/oracle.c by oracle/Makefile).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the timed CPU baseline.  The product package
(metrovector_b200) never imports it.

Parity status: "parity unpinned" for search results -- see the header of oracle/oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

F32, F16, I8, U8 = 0, 1, 2, 3          # schema/types.fbs:3-11
L2, IP, COS = 0, 1, 2                  # schema/types.fbs:19-25
NEAREST, AS_WRITTEN = 0, 1

_NP = {F32: np.float32, F16: np.float16, I8: np.int8, U8: np.uint8}


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def _load():
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "oracle.c")):
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    c = ctypes
    lib.oracle_f16_to_f32.restype = c.c_float
    lib.oracle_f16_to_f32.argtypes = [c.c_uint16]
    lib.oracle_f32_to_f16.restype = c.c_uint16
    lib.oracle_f32_to_f16.argtypes = [c.c_float]
    lib.oracle_l2_distance_f32.restype = c.c_float
    lib.oracle_l2_distance_f32.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32]
    lib.oracle_ordered_key_f32.restype = c.c_uint32
    lib.oracle_ordered_key_f32.argtypes = [c.c_float]
    lib.oracle_topk.restype = c.c_int
    lib.oracle_topk.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.c_uint32, c.c_int, c.c_void_p,
                                c.c_int, c.c_uint32, c.c_int, c.c_uint64, c.c_void_p, c.c_void_p,
                                c.c_void_p]
    lib.oracle_score_row.restype = c.c_int
    lib.oracle_score_row.argtypes = [c.c_void_p, c.c_int, c.c_uint32, c.c_int, c.c_void_p, c.c_int,
                                     c.c_void_p, c.c_void_p, c.c_void_p]
    lib.oracle_row_norms.restype = None
    lib.oracle_row_norms.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.c_uint32, c.c_void_p,
                                     c.c_void_p, c.c_void_p]
    lib.oracle_synth_fill.restype = None
    lib.oracle_synth_fill.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.c_uint64, c.c_uint64,
                                      c.c_uint32]
    lib.oracle_synth_fill_mt.restype = None
    lib.oracle_synth_fill_mt.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.c_uint64, c.c_uint64,
                                         c.c_uint32, c.c_int]
    lib.oracle_scan_timed.restype = c.c_double
    lib.oracle_scan_timed.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.c_uint32, c.c_int,
                                      c.c_void_p, c.c_int, c.c_uint32, c.c_uint32, c.c_int, c.c_int,
                                      c.c_void_p, c.c_void_p]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def np_dtype(dtype):
    return _NP[dtype]


def query_dtype_for(dtype, query):
    """f32 or f16 queries for float spaces; the space's own dtype for integer spaces."""
    if dtype in (F32, F16):
        return F16 if query.dtype == np.float16 else F32
    return dtype


def topk(db, dtype, metric, query, k, mode=NEAREST, row_base=0):
    """One query.  db: (n, d) array in the space's dtype.  Returns (ids u64, scores f32, iscores i32)."""
    db = np.ascontiguousarray(db, dtype=_NP[dtype])
    n, d = db.shape
    if dtype in (F32, F16):
        query = np.ascontiguousarray(query, dtype=np.float16 if query.dtype == np.float16 else np.float32)
    else:
        query = np.ascontiguousarray(query, dtype=_NP[dtype])
    assert query.shape == (d,)
    ids = np.zeros(k, dtype=np.uint64)
    scores = np.zeros(k, dtype=np.float32)
    iscores = np.zeros(k, dtype=np.int32)
    cnt = lib().oracle_topk(_ptr(db), dtype, n, d, metric, _ptr(query), query_dtype_for(dtype, query),
                            k, mode, row_base, _ptr(ids), _ptr(scores), _ptr(iscores))
    assert cnt >= 0
    return ids[:cnt], scores[:cnt], iscores[:cnt]


def score_row(row, dtype, metric, query):
    row = np.ascontiguousarray(row, dtype=_NP[dtype])
    d = row.shape[0]
    if dtype in (F32, F16):
        query = np.ascontiguousarray(query, dtype=np.float16 if query.dtype == np.float16 else np.float32)
    else:
        query = np.ascontiguousarray(query, dtype=_NP[dtype])
    s = ctypes.c_float()
    i = ctypes.c_int32()
    key = ctypes.c_uint32()
    lib().oracle_score_row(_ptr(row), dtype, d, metric, _ptr(query), query_dtype_for(dtype, query),
                           ctypes.byref(s), ctypes.byref(i), ctypes.byref(key))
    return s.value, i.value, key.value


def row_norms(db, dtype):
    db = np.ascontiguousarray(db, dtype=_NP[dtype])
    n, d = db.shape
    norm = np.zeros(n, dtype=np.float32)
    sqf = np.zeros(n, dtype=np.float32)
    sqi = np.zeros(n, dtype=np.int32)
    lib().oracle_row_norms(_ptr(db), dtype, n, d, _ptr(norm), _ptr(sqf), _ptr(sqi))
    return norm, sqf, sqi


def synth(dtype, seed, row0, nrows, d, threads=1):
    out = np.empty((nrows, d), dtype=_NP[dtype])
    if threads > 1:
        lib().oracle_synth_fill_mt(_ptr(out), dtype, seed, row0, nrows, d, threads)
    else:
        lib().oracle_synth_fill(_ptr(out), dtype, seed, row0, nrows, d)
    return out


def scan_timed(db, dtype, metric, queries, k, threads, faithful=False):
    """Times nq full scans on the CPU.  Returns (seconds, ids, scores) of the last query."""
    db = np.ascontiguousarray(db, dtype=_NP[dtype])
    n, d = db.shape
    if dtype in (F32, F16):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        qd = F32
    else:
        queries = np.ascontiguousarray(queries, dtype=_NP[dtype])
        qd = dtype
    nq = queries.shape[0]
    ids = np.zeros(k, dtype=np.uint64)
    scores = np.zeros(k, dtype=np.float32)
    sec = lib().oracle_scan_timed(_ptr(db), dtype, n, d, metric, _ptr(queries), qd, nq, k, threads,
                                  1 if faithful else 0, _ptr(ids), _ptr(scores))
    return sec, ids, scores


def f16_to_f32(h):
    return lib().oracle_f16_to_f32(int(h))


def f32_to_f16(f):
    return lib().oracle_f32_to_f16(float(f))


def l2_distance_f32(q, x):
    q = np.ascontiguousarray(q, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    return lib().oracle_l2_distance_f32(_ptr(q), _ptr(x), min(q.shape[0], x.shape[0]))
