"""Parity at BASELINE.json's full sizes, through size-independent properties (the CPU oracle cannot
scan 1e6..1e9 rows in a test, but it can regenerate and score any single row of the synthetic spaces).

Properties checked on device-generated spaces:
  * planted queries: a query equal to stored row r returns r first with the metric's identity score;
  * score consistency: every returned (id, score) pair equals the oracle's score of that regenerated row
    (bit-exact for integers, within the float tolerance otherwise), and the list is sorted best-first;
  * sampled completeness: no row of a 200k-row oracle-scanned sample beats the returned k-th score;
  * shard additivity: searching two half spaces (row_base applied) and merging on the host equals
    searching the whole space -- the single-GPU form of the multi-GPU identity.
"""
import numpy as np
import pytest

import metrovector_b200 as mv
import oracle
from util import as_query, ids_agree_within_tolerance

pytestmark = pytest.mark.gpu
SEED = 0x4D564631


def _check_consistency(dtype, metric, d, q, ids, sc, isc):
    is_int = dtype in (oracle.I8, oracle.U8)
    for gid, s, si in zip(ids, sc, isc):
        row = mv.synth_fill_host(dtype, SEED, int(gid), 1, d)[0]
        ref, iref, _ = oracle.score_row(row, dtype, metric, q)
        if is_int and metric != oracle.COS:
            assert si == iref and s == np.float32(iref)
        elif is_int:
            assert np.float32(s).view(np.uint32) == np.float32(ref).view(np.uint32)
        else:
            assert abs(float(s) - ref) <= 2e-5 * max(1.0, abs(ref)), (gid, s, ref)
    rank = -sc if metric == oracle.IP else sc
    assert np.all(np.diff(rank) >= 0)


@pytest.mark.parametrize("name,n,d,dtype,metric,k", [
    ("B", 1_000_000, 768, oracle.F32, oracle.L2, 10),            # BASELINE configs[1], full size
    ("A", 100_000, 128, oracle.F32, oracle.COS, 10),             # configs[0], full size
    ("D-slab", 12_500_000, 768, oracle.F16, oracle.COS, 10),     # one 8-GPU shard of configs[3]
    ("E-slab", 125_000_000, 128, oracle.U8, oracle.L2, 10),      # one 8-GPU shard of configs[4]
])
def test_full_size_properties(engine, name, n, d, dtype, metric, k):
    gs = engine.synthetic(mv.DataType(dtype), mv.DistanceMetric(metric), n, d, SEED)
    planted = [0, n // 3, n - 1]
    qrows = [mv.synth_fill_host(dtype, SEED, r, 1, d)[0] for r in planted]
    qrand = mv.synth_fill_host(dtype, SEED + 1, 0, 2, d)
    queries = np.stack([as_query(dtype, r) for r in qrows + list(qrand)])
    ids, sc, isc = gs.search(queries, k)
    for i, r in enumerate(planted):
        assert ids[i][0] == r, (name, r, ids[i])
        if metric == oracle.L2:
            assert sc[i][0] == 0.0
        else:
            assert abs(sc[i][0]) <= 2e-5
    for i in range(len(queries)):
        assert len(set(ids[i].tolist())) == k and ids[i].max() < n
        _check_consistency(dtype, metric, d, queries[i], ids[i], sc[i], isc[i])
    # sampled completeness on the two random queries
    sample_n = min(n, 200_000)
    sample = mv.synth_fill_host(dtype, SEED, n - sample_n, sample_n, d)
    for i in (3, 4):
        o_ids, o_sc, o_isc = oracle.topk(sample, dtype, metric, queries[i], 1, row_base=n - sample_n)
        best_in_sample = o_sc[0]
        kth = sc[i][k - 1]
        if metric == oracle.IP:
            assert best_in_sample <= sc[i][0] + 1e-3 * abs(sc[i][0])
        else:
            assert best_in_sample >= sc[i][0] - 2e-5 * max(1.0, abs(float(sc[i][0])))
            if best_in_sample < kth - 2e-5 * max(1.0, abs(float(kth))):
                assert int(o_ids[0]) in ids[i].tolist()
    gs.free()


@pytest.mark.parametrize("n,d,dtype,metric", [(1_000_000, 768, oracle.F32, oracle.L2), (20_000_000, 128, oracle.U8, oracle.L2)])
def test_shard_additivity_at_size(engine, n, d, dtype, metric):
    k, nq = 10, 3
    queries = np.stack([as_query(dtype, r) for r in mv.synth_fill_host(dtype, SEED + 1, 0, nq, d)])
    whole = engine.synthetic(mv.DataType(dtype), mv.DistanceMetric(metric), n, d, SEED)
    w_ids, w_sc, w_isc = whole.search(queries, k)
    whole.free()
    # two "ranks" on one GPU: upload each half with its row_base, merge (score, id) on the host
    half = n // 2
    parts = []
    for base, cnt in ((0, half), (half, n - half)):
        rows = mv.synth_fill_host(dtype, SEED, base, cnt, d)
        gs = engine.from_host(rows, mv.DataType(dtype), mv.DistanceMetric(metric), row_base=base, global_n=n)
        parts.append(gs.search(queries, k))
        gs.free()
    for qi in range(nq):
        cand = {}
        for ids, sc, isc in parts:
            for gid, s, si in zip(ids[qi], sc[qi], isc[qi]):
                cand[int(gid)] = (float(s), int(si))
        merged = sorted(cand.items(), key=lambda kv: (kv[1][0], kv[0]))[:k]
        assert [m[0] for m in merged] == w_ids[qi].tolist()
        assert [m[1][0] for m in merged] == [float(x) for x in w_sc[qi]]


# ---- the batched (tensor-core) path at BASELINE.json's full sizes ---------------------------------------------------
@pytest.mark.parametrize("name,n,d,dtype,metric,k,nq", [
    ("C", 10_000_000, 768, oracle.F16, oracle.COS, 100, 4096),          # BASELINE configs[2], full size
    ("D1024-slab", 12_500_000, 768, oracle.F16, oracle.COS, 10, 1024),  # one 8-GPU shard of configs[3], batch 1024
    ("E1024-slab", 125_000_000, 128, oracle.U8, oracle.L2, 10, 1024),   # one 8-GPU shard of configs[4], batch 1024
    ("B1024", 1_000_000, 768, oracle.F32, oracle.L2, 10, 1024),         # configs[1]'s space, batch 1024 (split tf32)
])
def test_full_size_tensor_path(engine, name, n, d, dtype, metric, k, nq):
    gs = engine.synthetic(mv.DataType(dtype), mv.DistanceMetric(metric), n, d, SEED)
    q = mv.synth_fill_host(dtype, SEED + 1, 0, nq, d)
    planted = {0: 0, 1: n // 3, 127: n // 2, nq - 128: 12345, nq - 1: n - 1}   # query slot -> stored row, first and last query tile
    for slot, r in planted.items():
        q[slot] = mv.synth_fill_host(dtype, SEED, r, 1, d)[0]
    try:
        engine.set_path(0)
        ids, sc, isc = gs.search(q, k)
        assert engine.profile_read()["last_path"] == 2, "the batch did not take the tensor-core path"
        for slot, r in planted.items():
            assert ids[slot][0] == r, (name, slot, r, ids[slot])
            if metric != oracle.L2:
                assert abs(sc[slot][0]) <= 2e-5
            elif dtype in (oracle.I8, oracle.U8):
                assert sc[slot][0] == 0.0 and isc[slot][0] == 0
            else:       # |q|^2 + |x|^2 - 2 q.x cancels to within the stated tolerance on d^2, not to an exact zero
                qn2 = float(np.dot(q[slot].astype(np.float64), q[slot].astype(np.float64)))
                assert float(sc[slot][0]) ** 2 <= 1e-5 * 2 * qn2, (name, slot, sc[slot][0])
        assert all(len(set(ids[i].tolist())) == k and ids[i].max() < n for i in range(nq))
        # every query of the first and of the last 128-query tile again through the bandwidth scan (itself checked
        # against the oracle at these sizes above): bit-identical for integers, the id rule for floats
        engine.set_path(1)
        for lo in (0, nq - 128):
            a = gs.search(q[lo:lo + 128], k)
            assert engine.profile_read()["last_path"] == 1
            if dtype in (oracle.I8, oracle.U8):
                assert np.array_equal(a[0], ids[lo:lo + 128]) and np.array_equal(a[2], isc[lo:lo + 128]), (name, lo)
            else:
                if metric == oracle.L2:      # the tolerance is stated on d^2: 1e-5 (|q|^2 + |x|^2); near d = 0 the root amplifies it
                    qn2 = (q[lo:lo + 128].astype(np.float64) ** 2).sum(1)[:, None]
                    assert (np.abs(a[1].astype(np.float64) ** 2 - sc[lo:lo + 128].astype(np.float64) ** 2) <= 1e-5 * 2.5 * qn2).all(), (name, lo)
                else:
                    assert np.allclose(a[1], sc[lo:lo + 128], rtol=1e-4, atol=2e-5), (name, lo)
                assert ids_agree_within_tolerance(metric, a[0], a[1], ids[lo:lo + 128], sc[lo:lo + 128]), (name, lo)
        # and the oracle's own score for every returned row of a few queries
        for i in (2, 64, nq - 2):
            qq = q[i].astype(np.float32) if dtype in (oracle.F32, oracle.F16) else q[i]
            _check_consistency(dtype, metric, d, qq, ids[i], sc[i], isc[i])
    finally:
        engine.set_path(0)
        gs.free()
