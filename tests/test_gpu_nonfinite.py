"""Inf / NaN rows and queries (a3 of SURVEY.md section 8: the reference's `partial_cmp(..).unwrap_or(Equal)`,
examples/similarity_search.rs:29, leaves their order to the heap; DESIGN.md section 2 defines it: NaN last, ties by
index).  Both entry paths -- the scan kernel and a batch that qualifies for the tensor path -- must give exactly the
oracle's answer, including when fewer than k rows have a finite score.  Float16 spaces hold Inf/NaN bit patterns
legally; Float32 ones too."""
import numpy as np
import pytest

import metrovector_b200 as mv
import oracle

pytestmark = pytest.mark.gpu
DT = {oracle.F32: mv.DataType.Float32, oracle.F16: mv.DataType.Float16}


def _same(got, want):
    """scores equal where both finite or both the same infinity; NaN matches NaN"""
    got, want = np.asarray(got, dtype=np.float32), np.asarray(want, dtype=np.float32)
    nan = np.isnan(want)
    return bool(np.array_equal(np.isnan(got), nan) and np.allclose(got[~nan], want[~nan], rtol=1e-5, atol=1e-6))


def _check(gs, eng, db, dtype, queries, k, metric):
    ids, sc, _ = gs.search(queries, k, metric=metric)
    for qi in range(len(queries)):
        o_ids, o_sc, _ = oracle.topk(db, dtype, metric, queries[qi], k)
        assert ids[qi][:len(o_ids)].tolist() == o_ids.tolist(), (metric, qi, ids[qi], o_ids, sc[qi], o_sc)
        assert _same(sc[qi][:len(o_ids)], o_sc), (metric, qi, sc[qi], o_sc)


@pytest.mark.parametrize("dtype", [oracle.F32, oracle.F16])
@pytest.mark.parametrize("path", [1, 0, 2])
def test_nonfinite_rows_sort_last_on_every_path(engine, dtype, path):
    # 6000 rows of which only 7 are finite, k = 10: the tail of every result is made of +inf / NaN rows in index order
    n, d, k = 6000, 64, 10
    npdt = np.float32 if dtype == oracle.F32 else np.float16
    db = mv.synth_fill_host(dtype, 71, 0, n, d).astype(npdt)
    rng = np.random.default_rng(5)
    finite_rows = sorted(rng.choice(n, 7, replace=False).tolist())
    bad = np.ones(n, dtype=bool)
    bad[finite_rows] = False
    for r in np.flatnonzero(bad):
        db[r, rng.integers(0, d)] = [np.inf, -np.inf, np.nan][r % 3]
    gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
    assert gs.path_facts()["nonfinite_rows"] == n - 7
    engine.set_path(path)
    try:
        for nq in (1, 40):                                                    # 40 queries would qualify for the tensor path
            q = mv.synth_fill_host(dtype, 72, 0, nq, d)
            queries = q if (dtype == oracle.F16 and nq >= 4) else q.astype(np.float32)
            for metric in (oracle.L2, oracle.COS, oracle.IP):
                _check(gs, engine, db, dtype, queries, k, metric)
                assert engine.profile_read()["last_path"] == 1                # such a space stays on the scan kernel
    finally:
        engine.set_path(0)
        gs.free()


@pytest.mark.parametrize("dtype", [oracle.F32, oracle.F16])
def test_nonfinite_queries_in_a_tensor_sized_batch(engine, dtype):
    # a finite space, a batch large enough for the tensor path, a few queries holding Inf / NaN: the batch is answered
    # by the scan kernel (host-buffer and device-buffer entry points alike) and equals the oracle, query by query
    n, d, k, nq = 20000, 128, 10, 48
    npdt = np.float32 if dtype == oracle.F32 else np.float16
    db = mv.synth_fill_host(dtype, 73, 0, n, d).astype(npdt)
    db[17] = 0                                                                # a zero-norm row (cosine: distance 1)
    gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.Cosine)
    q = mv.synth_fill_host(dtype, 74, 0, nq, d).astype(npdt)
    for metric in (oracle.L2, oracle.COS, oracle.IP):
        ids, sc, _ = gs.search(q, k, metric=metric)                          # all finite: tensor path
        assert engine.profile_read()["last_path"] == 2
    qb = q.copy()
    qb[3, 5] = np.inf
    qb[9, 0] = np.nan
    qb[20, 7] = -np.inf
    qb[21, :] = np.inf
    for metric in (oracle.L2, oracle.COS, oracle.IP):
        ids, sc, _ = gs.search(qb, k, metric=metric)
        assert engine.profile_read()["last_path"] == 1
        for qi in (0, 3, 9, 20, 21, 47):
            o_ids, o_sc, _ = oracle.topk(db, dtype, metric, qb[qi], k)
            assert ids[qi].tolist() == o_ids.tolist(), (metric, qi, ids[qi], o_ids, sc[qi], o_sc)
            assert _same(sc[qi], o_sc), (metric, qi, sc[qi], o_sc)
    # the device-pointer entry point makes the same choice (its finite check runs on the GPU)
    import torch
    qd = torch.from_numpy(qb).cuda()
    ids_d = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    sc_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    gs.search_device(qd.data_ptr(), int(dtype), nq, k, oracle.L2, ids_d.data_ptr(), sc_d.data_ptr())
    engine.synchronize()
    assert engine.profile_read()["last_path"] == 1
    ids_h, sc_h, _ = gs.search(qb, k, metric=oracle.L2)
    assert np.array_equal(ids_d.cpu().numpy().astype(np.uint64), ids_h)
    gs.free()


def test_fewer_than_k_rows_and_all_nan(engine):
    db = np.full((5, 32), np.nan, dtype=np.float32)
    db[2] = 1.0
    gs = engine.from_host(db, mv.DataType.Float32, mv.DistanceMetric.L2)
    ids, sc, _ = gs.search(np.ones((1, 32), dtype=np.float32), 8)
    assert ids[0][:5].tolist() == [2, 0, 1, 3, 4] and (ids[0][5:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert sc[0][0] == 0.0 and np.isnan(sc[0][1:5]).all()
    gs.free()
