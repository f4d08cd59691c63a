"""The oracle against every known answer we have for this path (CPU only).

The reference holds no golden vector for the scan, so the anchors are
(1) the values derived by hand from the reference's example datasets with its exact arithmetic
(restated below as literals), (2) IEEE f16 conversion facts pinned by
src/vectors/vector.rs:227-241, (3) independent numpy restatements for the metrics we define.
"""
import json
import os

import numpy as np
import pytest

import oracle
from util import f32_bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def clustered_rows():
    f = np.float32
    rows = []
    for base in ((1, 1, 1, 1), (5, 5, 5, 5)):
        for i in range(20):
            n = f(i) * f(0.1)
            rows.append([f(base[0]) + n, f(base[1]) - n, f(base[2]) + n * f(0.5), f(base[3]) - n * f(0.5)])
    for i in range(20):
        n = f(i) * f(0.1)
        rows.append([f(-2.0) + n, f(3.0) - n, f(0.0) + n, f(4.0) - n * f(0.5)])
    return np.array(rows, dtype=f)


def test_simple_rs_known_answer():
    # examples/simple.rs:15-21,70 ; SURVEY.md Appendix B.1
    rows = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [2, 4, 6, 8], [1, 3, 5, 7]], dtype=np.float32)
    q = np.array([2.5, 4.5, 6.5, 8.5], dtype=np.float32)
    expect = [0x40cce665, 0x40400000, 0x4123f383, 0x3f800000, 0x40400000]
    for i in range(5):
        assert f32_bits(oracle.l2_distance_f32(q, rows[i])) == expect[i]
    ids, sc, _ = oracle.topk(rows, oracle.F32, oracle.L2, q, 3)
    assert ids.tolist() == [3, 1, 4]          # ids 1 and 4 tie at exactly 3.0 -> lowest index first
    ids, _, _ = oracle.topk(rows, oracle.F32, oracle.L2, q, 1)
    assert ids.tolist() == [3]                # simple.rs:73-94 strict-< argmin


@pytest.mark.parametrize("query,ids,bits", [
    ([1, 1, 1, 1], [0, 1, 2, 3, 4], [0x00000000, 0x3e21e89c, 0x3ea1e89c, 0x3ef2dce7, 0x3f21e89b]),
    ([5, 5, 5, 5], [20, 21, 22, 23, 24], [0x00000000, 0x3e21e89c, 0x3ea1e891, 0x3ef2dcf2, 0x3f21e89b]),
    ([-2, 3, 0, 4], [40, 41, 42, 43, 44], [0x00000000, 0x3e389aab, 0x3eb89aaf, 0x3f0a7403, 0x3f389ab0]),
    ([0, 0, 0, 0], [0, 1, 2, 3, 4], [0x40000000, 0x4000663e, 0x40019713, 0x40038cfe, 0x40063f5e]),
])
def test_similarity_search_rs_known_answers(query, ids, bits):
    # examples/similarity_search.rs:43-76,103-116 ; SURVEY.md Appendix B.2
    got_ids, sc, _ = oracle.topk(clustered_rows(), oracle.F32, oracle.L2, np.array(query, dtype=np.float32), 5)
    assert got_ids.tolist() == ids
    assert [f32_bits(s) for s in sc] == bits


def test_as_written_keeps_the_farthest():
    # the reversed Ord + pop (similarity_search.rs:23-31,165-168) keeps the k FARTHEST rows
    db = clustered_rows()
    ids, sc, _ = oracle.topk(db, oracle.F32, oracle.L2, np.array([1, 1, 1, 1], dtype=np.float32), 5, mode=oracle.AS_WRITTEN)
    assert ids.tolist() == [35, 36, 37, 38, 39]
    assert 8.34 < sc[0] < 8.35 and 8.54 < sc[4] < 8.55
    ids, _, _ = oracle.topk(db, oracle.F32, oracle.L2, np.array([5, 5, 5, 5], dtype=np.float32), 5, mode=oracle.AS_WRITTEN)
    assert ids.tolist() == [19, 43, 42, 41, 40]


def test_committed_known_answers_reproduce():
    with open(os.path.join(GOLD, "known_answers.json")) as fh:
        known = json.load(fh)
    rows = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [2, 4, 6, 8], [1, 3, 5, 7]], dtype=np.float32)
    c = known["simple"]
    ids, sc, _ = oracle.topk(rows, oracle.F32, oracle.L2, np.array(c["query"], dtype=np.float32), c["k"])
    assert ids.tolist() == c["ids"] and [f32_bits(s) for s in sc] == c["score_bits"]
    for c in known["seeded"]:
        db = oracle.synth(c["dtype"], c["seed"], 0, c["n"], c["d"])
        q = oracle.synth(c["dtype"], c["qseed"], 0, 2, c["d"])[c["qrow"]]
        if c["planted"] is not None:
            q = db[c["planted"]]
        if c["dtype"] in (oracle.F32, oracle.F16):
            q = q.astype(np.float32)
        ids, sc, isc = oracle.topk(db, c["dtype"], c["metric"], q, c["k"])
        assert ids.tolist() == c["ids"]
        assert [f32_bits(s) for s in sc] == c["score_bits"]
        assert isc.tolist() == c["iscores"]
        if c["planted"] is not None and c["metric"] == oracle.L2:
            assert ids[0] == c["planted"] and sc[0] == 0.0


def test_f16_conversion_facts():
    # src/vectors/vector.rs:227-241: f16(pi), f16(e) -> f32
    assert oracle.f32_to_f16(3.14159265) == 0x4248 and oracle.f16_to_f32(0x4248) == 3.140625
    assert oracle.f32_to_f16(2.718281828) == 0x4170 and oracle.f16_to_f32(0x4170) == 2.71875
    allh = np.arange(65536, dtype=np.uint16)
    ref = allh.view(np.float16).astype(np.float32)
    mine = np.array([oracle.f16_to_f32(int(h)) for h in allh], dtype=np.float32)
    ok = ~np.isnan(ref)
    assert np.array_equal(ref.view(np.uint32)[ok], mine.view(np.uint32)[ok])
    assert np.isnan(mine[~ok]).all()
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1, 1e-5, 1e-7, 6e4, 1e-3)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = np.array([oracle.f32_to_f16(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("dtype", [oracle.U8, oracle.I8])
@pytest.mark.parametrize("metric", [oracle.L2, oracle.IP, oracle.COS])
def test_integer_metrics_against_numpy(dtype, metric):
    n, d, k = 500, 48, 20
    db = oracle.synth(dtype, 5, 0, n, d)
    q = oracle.synth(dtype, 6, 0, 1, d)[0]
    ids, sc, isc = oracle.topk(db, dtype, metric, q, k)
    a = db.astype(np.int64)
    b = q.astype(np.int64)
    dot = a @ b
    if metric == oracle.L2:
        val = ((a - b) ** 2).sum(1)
    elif metric == oracle.IP:
        val = -dot
    else:
        den = np.sqrt(np.float32((b * b).sum())) * np.sqrt((a * a).sum(1).astype(np.float32))
        val = (np.float32(1.0) - dot.astype(np.float32) / den).astype(np.float32)
    order = np.lexsort((np.arange(n), val))[:k]
    assert ids.tolist() == order.tolist()
    if metric == oracle.L2:
        assert isc.tolist() == val[order].tolist()
    elif metric == oracle.IP:
        assert isc.tolist() == dot[order].tolist()
    else:
        assert np.array_equal(sc.view(np.uint32), val[order].view(np.uint32))


@pytest.mark.parametrize("dtype", [oracle.F32, oracle.F16])
@pytest.mark.parametrize("metric", [oracle.L2, oracle.IP, oracle.COS])
def test_float_metrics_against_numpy_f64(dtype, metric):
    n, d, k = 400, 96, 10
    db = oracle.synth(dtype, 15, 0, n, d)
    q = oracle.synth(dtype, 16, 0, 1, d)[0].astype(np.float32)
    ids, sc, _ = oracle.topk(db, dtype, metric, q, k)
    a = db.astype(np.float64)
    b = q.astype(np.float64)
    if metric == oracle.L2:
        val = np.sqrt(((a - b) ** 2).sum(1))
    elif metric == oracle.IP:
        val = a @ b
    else:
        val = 1 - (a @ b) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b))
    order = np.argsort(-val if metric == oracle.IP else val, kind="stable")[:k]
    assert ids.tolist() == order.tolist()
    assert np.allclose(sc, val[order], rtol=1e-5, atol=1e-6)


def test_ties_lowest_index_and_k_larger_than_n():
    db = np.array([[3, 3], [1, 1], [3, 3], [1, 1], [2, 2]], dtype=np.uint8)
    q = np.array([1, 1], dtype=np.uint8)
    ids, _, isc = oracle.topk(db, oracle.U8, oracle.L2, q, 10)
    assert ids.tolist() == [1, 3, 4, 0, 2] and isc.tolist() == [0, 0, 2, 8, 8]
    zero = np.zeros((3, 4), dtype=np.float32)
    ids, sc, _ = oracle.topk(zero, oracle.F32, oracle.COS, np.ones(4, dtype=np.float32), 2)
    assert ids.tolist() == [0, 1] and sc.tolist() == [1.0, 1.0]      # zero-norm rows -> distance 1


def test_threaded_baseline_matches_single_scan():
    db = oracle.synth(oracle.F32, 3, 0, 5000, 64)
    q = oracle.synth(oracle.F32, 4, 0, 2, 64)
    ids, sc, _ = oracle.topk(db, oracle.F32, oracle.L2, q[1], 10)
    for faithful in (False, True):
        _, ids2, sc2 = oracle.scan_timed(db, oracle.F32, oracle.L2, q, 10, threads=4, faithful=faithful)
        assert ids2.tolist() == ids.tolist() and np.array_equal(sc2, sc)


def test_oracle_bits_equal_a_python_float32_restatement():
    """Independent restatement of the reference expression in numpy float32 scalars, one rounded op at a
    time (examples/similarity_search.rs:152-157: t = a - b; s = s + t*t; sqrt): the C oracle must agree
    bit for bit, which also proves it was built without FMA contraction or reassociation."""
    f = np.float32
    rng = np.random.default_rng(11)
    for dtype in (oracle.F32, oracle.F16):
        db = oracle.synth(dtype, 91, 0, 40, 51)
        q = oracle.synth(dtype, 92, 0, 1, 51)[0].astype(np.float32)
        want = []
        for i in range(40):
            x = db[i].astype(np.float32)
            s = f(0.0)
            for j in range(51):
                t = f(q[j] - x[j])
                s = f(s + f(t * t))
            want.append(f(np.sqrt(s)))
        want = np.array(want, dtype=np.float32)
        order = np.lexsort((np.arange(40), want))
        ids, sc, _ = oracle.topk(db, dtype, oracle.L2, q, 40)
        assert ids.tolist() == order.tolist()
        assert np.array_equal(sc.view(np.uint32), want[order].view(np.uint32))
        # cosine and dot, same discipline (ours)
        dots, nx = [], []
        for i in range(40):
            x = db[i].astype(np.float32)
            s = f(0.0); n2 = f(0.0)
            for j in range(51):
                s = f(s + f(q[j] * x[j])); n2 = f(n2 + f(x[j] * x[j]))
            dots.append(s); nx.append(n2)
        qn = f(0.0)
        for j in range(51):
            qn = f(qn + f(q[j] * q[j]))
        cos = np.array([f(f(1.0) - f(dd / f(f(np.sqrt(qn)) * f(np.sqrt(n2))))) for dd, n2 in zip(dots, nx)], dtype=np.float32)
        ids, sc, _ = oracle.topk(db, dtype, oracle.COS, q, 40)
        order = np.lexsort((np.arange(40), cos))
        assert ids.tolist() == order.tolist() and np.array_equal(sc.view(np.uint32), cos[order].view(np.uint32))
        ids, sc, _ = oracle.topk(db, dtype, oracle.IP, q, 40)
        dd = np.array(dots, dtype=np.float32)
        order = np.lexsort((np.arange(40), -dd))
        assert ids.tolist() == order.tolist() and np.array_equal(sc.view(np.uint32), dd[order].view(np.uint32))
