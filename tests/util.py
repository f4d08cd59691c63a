"""Shared checkers for the parity tests: the CUDA path against the CPU oracle."""
import struct

import numpy as np

import oracle

INVALID = np.uint64(0xFFFFFFFFFFFFFFFF)

# Float tolerances: compared on the pre-sqrt / pre-divide quantity.
TOL_L2_REL = 1e-5      # |d2_gpu - d2_ref| <= 1e-5 * (|q|^2 + |x|^2)
TOL_IP_REL = 1e-5      # |dot_gpu - dot_ref| <= 1e-5 * |q| * |x|
TOL_COS_ABS = 2e-5     # |cos_gpu - cos_ref| <= 2e-5


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def as_query(dtype, row):
    """Queries are f32 for float spaces, the space's own dtype for integer spaces."""
    if dtype in (oracle.F32, oracle.F16):
        return np.asarray(row).astype(np.float32)
    return np.asarray(row).astype(oracle.np_dtype(dtype))


def _rank_value(metric, score, is_int):
    """The quantity ranking happens on, in float64: L2 -> d^2 (ints already squared), IP -> -dot, COS -> 1-cos."""
    s = float(score)
    if metric == oracle.L2:
        return s if is_int else s * s
    if metric == oracle.IP:
        return -s
    return s


def check_one(db, dtype, metric, query, k, ids, scores, iscores):
    """Asserts one query's GPU result against the oracle.  db is the full (global) matrix."""
    n, d = db.shape
    is_int = dtype in (oracle.I8, oracle.U8)
    kk = min(k, n)
    o_ids, o_sc, o_isc = oracle.topk(db, dtype, metric, query, min(n, k + 32))
    valid = ids != INVALID
    assert int(valid.sum()) == kk, "expected %d results, got %d" % (kk, int(valid.sum()))
    assert valid[:kk].all()
    g_ids = ids[:kk].astype(np.int64)
    assert len(set(g_ids.tolist())) == kk, "duplicate ids"
    assert (g_ids >= 0).all() and (g_ids < n).all()
    if is_int:
        # bit-exact: ids, integer scores (L2 / InnerProduct), float score bits (Cosine); ties -> lowest index
        assert np.array_equal(g_ids, o_ids[:kk].astype(np.int64)), (g_ids, o_ids[:kk])
        if metric != oracle.COS:
            assert np.array_equal(iscores[:kk], o_isc[:kk])
            assert np.array_equal(scores[:kk], o_isc[:kk].astype(np.float32))
        else:
            assert np.array_equal(scores[:kk].view(np.uint32), o_sc[:kk].view(np.uint32))
        return
    # floats: tolerance on the ranking quantity
    qf = np.asarray(query, dtype=np.float64)
    qn2 = float(np.dot(qf, qf))
    r_gpu = np.array([_rank_value(metric, s, False) for s in scores[:kk]])
    assert np.all(np.diff(r_gpu) >= 0), "result not sorted best-first"
    r_ref = np.empty(kk)
    tol = np.empty(kk)
    for j, gid in enumerate(g_ids):
        row = db[gid]
        s_ref, _, _ = oracle.score_row(row, dtype, metric, query)
        r_ref[j] = _rank_value(metric, s_ref, False)
        xf = row.astype(np.float64)
        xn2 = float(np.dot(xf, xf))
        if metric == oracle.L2:
            tol[j] = TOL_L2_REL * (qn2 + xn2)
        elif metric == oracle.IP:
            tol[j] = TOL_IP_REL * np.sqrt(qn2 * xn2)
        else:
            tol[j] = TOL_COS_ABS
    assert np.all(np.abs(r_gpu - r_ref) <= tol + 1e-30), (r_gpu, r_ref, tol)
    # id rule: nothing returned is worse than the oracle's k-th by more than the tolerance, and
    # everything clearly better than the oracle's k-th is returned
    o_r = np.array([_rank_value(metric, s, False) for s in o_sc])
    kth = o_r[kk - 1]
    tmax = float(tol.max()) if kk else 0.0
    assert np.all(r_ref <= kth + 2 * tmax)
    must = set(int(i) for i, r in zip(o_ids, o_r) if r < kth - 2 * tmax)
    assert must.issubset(set(int(i) for i in g_ids)), (must, g_ids)


def check_batch(db, dtype, metric, queries, k, ids, scores, iscores):
    for qi in range(len(queries)):
        check_one(db, dtype, metric, queries[qi], k, ids[qi], scores[qi], iscores[qi])


def ids_agree_within_tolerance(metric, ids_a, sc_a, ids_b, sc_b):
    """Two float result sets of the same queries (e.g. scan kernel vs tensor kernel, whose summation orders differ):
    the id rule of SURVEY.md section 8a without an oracle -- every id one side ranks clearly inside its top-k (better
    than BOTH k-th scores by more than twice the tolerance) must be in the other side's list.  Positions may differ
    only among scores closer than that."""
    for q in range(ids_a.shape[0]):
        if np.array_equal(ids_a[q], ids_b[q]):
            continue
        ra = np.array([_rank_value(metric, s, False) for s in sc_a[q]])
        rb = np.array([_rank_value(metric, s, False) for s in sc_b[q]])
        scale = max(1.0, float(np.abs(ra).max()), float(np.abs(rb).max()))
        tol = (TOL_COS_ABS if metric == oracle.COS else 1e-5) * scale
        kth = min(ra[-1], rb[-1])
        clear_a = set(int(i) for i, r in zip(ids_a[q], ra) if r < kth - 2 * tol)
        clear_b = set(int(i) for i, r in zip(ids_b[q], rb) if r < kth - 2 * tol)
        if not clear_a.issubset(set(int(i) for i in ids_b[q])) or not clear_b.issubset(set(int(i) for i in ids_a[q])):
            return False
    return True
