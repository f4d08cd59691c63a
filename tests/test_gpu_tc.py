"""K2, the tcgen05 tensor-core path, against the CPU oracle (and numpy for the raw contraction)."""
import numpy as np
import pytest

import metrovector_b200 as mv
import oracle
from util import check_batch, ids_agree_within_tolerance

pytestmark = pytest.mark.gpu
DT = {oracle.F32: mv.DataType.Float32, oracle.F16: mv.DataType.Float16, oracle.I8: mv.DataType.Int8, oracle.U8: mv.DataType.UInt8}
FLOATS = (oracle.F32, oracle.F16)       # Float32 spaces run the split-tf32 form (K2'), Float16 ones kind::f16


def _rank_values(db, q, dtype, metric):
    a = db.astype(np.float64)
    b = q.astype(np.float64)
    dot = b @ a.T
    if metric == oracle.L2:
        return (a * a).sum(1)[None, :] - 2 * dot
    if metric == oracle.IP:
        return -dot
    nrm = np.sqrt((a * a).sum(1))
    inv = np.where(nrm > 0, 1.0 / np.maximum(nrm, 1e-30), 0.0)
    return -dot * inv[None, :]


@pytest.mark.parametrize("dtype,d", [(oracle.F16, 64), (oracle.F16, 768), (oracle.F16, 200), (oracle.U8, 128), (oracle.I8, 96),
                                     (oracle.U8, 300), (oracle.F32, 32), (oracle.F32, 768), (oracle.F32, 100)])
def test_raw_contraction_matches_numpy(engine, dtype, d):
    n, nq = 700, 130                       # 3 database tiles (last one partial), 2 query tiles (last one partial)
    db = mv.synth_fill_host(dtype, 31, 0, n, d)
    q = mv.synth_fill_host(dtype, 32, 0, nq, d)
    gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
    for metric in (oracle.L2, oracle.IP, oracle.COS):
        got = gs.tc_dump(q, k=10, metric=metric)[:nq, :n]
        want = _rank_values(db, q, dtype, metric)
        if dtype not in FLOATS and metric == oracle.COS:
            # integer cosine: the common path's approximate ranking value float(a) * (-1 / |x|) (the candidates' keys are exact)
            assert np.allclose(got, want, rtol=1e-5, atol=1e-6 * np.abs(want).max()), np.abs(got - want).max()
        elif dtype in FLOATS:
            scale = np.abs(want).max()
            # f16 products are exact in f32; the split-tf32 product drops a 2^-22 term: both far inside 1e-5 of the scale
            assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * scale), (metric, np.abs(got - want).max())
            if dtype == oracle.F32 and metric != oracle.COS:
                # the stated tolerance itself (tests/util.py): |dot error| <= 1e-5 |q||x|.  What is left of it is
                # mostly the tensor core's f32 accumulation over 3 * d / 8 MMA steps, not the dropped lo.lo term
                qx = np.sqrt((q.astype(np.float64) ** 2).sum(1))[:, None] * np.sqrt((db.astype(np.float64) ** 2).sum(1))[None, :]
                err = np.abs(got - want) / (2.0 if metric == oracle.L2 else 1.0)
                assert (err <= 1e-5 * qx).all(), (metric, float((err / qx).max()))
        else:
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), metric
    gs.free()


@pytest.mark.parametrize("dtype,d,n", [(oracle.F16, 128, 20000), (oracle.F16, 768, 9000), (oracle.U8, 128, 30000),
                                       (oracle.I8, 64, 12345), (oracle.F16, 72, 5000), (oracle.F32, 128, 20000),
                                       (oracle.F32, 768, 6000), (oracle.F32, 72, 5000)])
def test_tensor_path_search_matches_oracle(engine, dtype, d, n):
    db = mv.synth_fill_host(dtype, 41, 0, n, d)
    gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
    engine.set_path(2)
    try:
        metrics = (oracle.L2, oracle.IP, oracle.COS)      # integers too: their cosine keys are formed exactly for the candidates
        for nq, k in ((200, 10), (64, 100), (3, 1), (129, 33), (130, 128), (130, 129), (140, 257), (64, 300), (260, 512)):
            q = mv.synth_fill_host(dtype, 42, 0, nq, d)
            q[0] = db[n // 2]
            # one query tile of a Float32 space with k > 56 has no split-tf32 launch shape (96 KiB stages beside a
            # 512-entry sort scratch): that combination stays on the scan kernel by design
            want_path = 1 if (dtype == oracle.F32 and nq <= 128 and k > 56) else 2
            for metric in metrics:
                ids, sc, isc = gs.search(q, k, metric=metric)
                assert engine.profile_read()["last_path"] == want_path
                qq = q.astype(np.float32) if dtype == oracle.F16 else q
                check_batch(db, dtype, metric, qq[:16], k, ids[:16], sc[:16], isc[:16])
                check_batch(db, dtype, metric, qq[-3:], k, ids[-3:], sc[-3:], isc[-3:])
    finally:
        engine.set_path(0)
        gs.free()


@pytest.mark.parametrize("dtype,n,d,nq,k", [(oracle.U8, 50000, 256, 96, 10), (oracle.U8, 3_000_000, 128, 200, 10),
                                            (oracle.I8, 2_000_000, 96, 130, 33), (oracle.F16, 1_000_000, 128, 130, 10),
                                            (oracle.F32, 1_000_000, 128, 130, 10), (oracle.F32, 400_000, 768, 300, 100),
                                            # k beyond 128: 512- and 1024-entry candidate buffers
                                            (oracle.U8, 2_000_000, 128, 200, 300), (oracle.F16, 1_000_000, 128, 300, 512),
                                            (oracle.F16, 600_000, 768, 260, 200), (oracle.F32, 300_000, 256, 260, 400)])
def test_tensor_and_scan_paths_agree(engine, dtype, n, d, nq, k):
    # sizes beyond the CPU oracle: the two GPU paths (each checked against the oracle at small sizes) must agree
    # on every metric.  Many database tiles per CTA exercise the threshold / early-out logic of the K2 epilogue.
    gs = engine.synthetic(DT[dtype], mv.DistanceMetric.L2, n, d, 51)
    q = mv.synth_fill_host(dtype, 52, 0, nq, d)
    try:
        for metric in (oracle.L2, oracle.IP, oracle.COS):
            engine.set_path(1)
            a = gs.search(q, k, metric=metric)
            engine.set_path(2)
            b = gs.search(q, k, metric=metric)
            assert engine.profile_read()["last_path"] == 2
            if dtype in FLOATS:                                              # accumulation order differs
                assert np.allclose(a[1], b[1], rtol=1e-4, atol=1e-5), metric
                assert ids_agree_within_tolerance(metric, a[0], a[1], b[0], b[1]), metric
            else:                                                            # integers: bit-identical (cosine: the float score bits)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]), metric
                assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), metric
    finally:
        engine.set_path(0)
        gs.free()


def test_ties_and_degenerate_queries_on_the_tensor_path(engine):
    # all-equal ranking values (zero query under cosine, duplicated rows, 2-valued bytes): exact lowest-index
    # tie-breaking must survive the threshold logic (strict vs. the local list, non-strict vs. other CTAs' bounds)
    engine.set_path(2)
    try:
        base = mv.synth_fill_host(oracle.F16, 3, 0, 500, 64)
        db = np.concatenate([base] * 40)                                  # every row appears 40 times
        gs = engine.from_host(db, mv.DataType.Float16, mv.DistanceMetric.Cosine)
        q = np.concatenate([np.zeros((1, 64), dtype=np.float16), base[:70]])
        for metric in (oracle.COS, oracle.L2, oracle.IP):
            ids, sc, isc = gs.search(q, 50, metric=metric)
            check_batch(db, oracle.F16, metric, q.astype(np.float32)[:6], 50, ids[:6], sc[:6], isc[:6])
        ids, sc, _ = gs.search(q[:1], 10, metric=oracle.COS)              # zero query: every distance is 1 -> ids 0..9
        assert ids[0].tolist() == list(range(10)) and (sc[0] == 1.0).all()
        gs.free()
        rng = np.random.default_rng(4)
        dbu = rng.integers(0, 2, (60000, 128)).astype(np.uint8)
        gu = engine.from_host(dbu, mv.DataType.UInt8, mv.DistanceMetric.L2)
        qu = rng.integers(0, 2, (130, 128)).astype(np.uint8)
        ids, sc, isc = gu.search(qu, 64)
        check_batch(dbu, oracle.U8, oracle.L2, qu[:4], 64, ids[:4], sc[:4], isc[:4])
        check_batch(dbu, oracle.U8, oracle.L2, qu[-2:], 64, ids[-2:], sc[-2:], isc[-2:])
        gu.free()
        # integer cosine: few distinct bytes -> many exactly equal keys; zero rows (distance 1) and a zero query
        dbc = rng.integers(0, 3, (50000, 64)).astype(np.uint8)
        dbc[100:140] = 0
        dbc[20000:20040] = dbc[:40]
        gc = engine.from_host(dbc, mv.DataType.UInt8, mv.DistanceMetric.Cosine)
        qc = rng.integers(0, 3, (140, 64)).astype(np.uint8)
        qc[0] = 0
        qc[1] = dbc[5]
        ids, sc, isc = gc.search(qc, 40)
        assert engine.profile_read()["last_path"] == 2
        check_batch(dbc, oracle.U8, oracle.COS, qc[:6], 40, ids[:6], sc[:6], isc[:6])
        check_batch(dbc, oracle.U8, oracle.COS, qc[-3:], 40, ids[-3:], sc[-3:], isc[-3:])
        assert ids[0].tolist() == list(range(40)) and (sc[0] == 1.0).all()
        gc.free()
    finally:
        engine.set_path(0)


@pytest.mark.parametrize("env", [{"MVF_TC_PAIR": "2"}, {"MVF_TC_PAIR": "0"}, {"MVF_TC_CLUSTER": "1"},
                                 {"MVF_TC_PAIR": "2", "MVF_TC_SYNC_WINDOW": "2"}, {"MVF_TC_PAIR": "0", "MVF_TC_SYNC_WINDOW": "1"}])
def test_every_kernel_form_matches_oracle(env, monkeypatch):
    # the knobs are read when an engine first uses the tensor path: CTA-pair MMAs (cta_group::2), multicast
    # clusters, single CTAs, and the producer pacing, each forced on shapes that would not pick it by default
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    eng = mv.GpuEngine(0)
    eng.set_path(2)
    try:
        for dtype, d, n in ((oracle.U8, 128, 40000), (oracle.I8, 320, 9000), (oracle.F16, 96, 30000), (oracle.F32, 96, 30000)):
            db = mv.synth_fill_host(dtype, 61, 0, n, d)
            gs = eng.from_host(db, DT[dtype], mv.DistanceMetric.L2)
            q = mv.synth_fill_host(dtype, 62, 0, 300, d)
            for metric in (oracle.L2, oracle.IP, oracle.COS):
                ids, sc, isc = gs.search(q, 10, metric=metric)
                assert eng.profile_read()["last_path"] == 2
                qq = q.astype(np.float32) if dtype == oracle.F16 else q
                check_batch(db, dtype, metric, qq[:8], 10, ids[:8], sc[:8], isc[:8])
                check_batch(db, dtype, metric, qq[-4:], 10, ids[-4:], sc[-4:], isc[-4:])
            gs.free()
    finally:
        eng.shutdown()


def test_randomised_shapes_on_the_tensor_path(engine):
    # seeded sweep over what the fixed cases above do not pin: ragged dimensions (K blocks partly out of bounds), row
    # counts around tile boundaries, query counts around tile boundaries, every k class, a row mask (integer L2: folded
    # into the accumulator seeds; other kinds: live-column words), every kind x metric -- each against the oracle
    rng = np.random.default_rng(20261018)
    engine.set_path(2)
    try:
        for case in range(36):
            dtype = [oracle.F16, oracle.U8, oracle.I8, oracle.F32][case % 4]
            d = int(rng.choice([8, 24, 40, 64, 100, 128, 136, 200, 256, 260, 384]))
            n = int(rng.choice([1, 255, 256, 257, 511, 1000, 4097, 9999, 20000]))
            nq = int(rng.choice([1, 5, 127, 128, 129, 200, 256, 257]))
            k = int(rng.choice([1, 7, 10, 32, 33, 56]))
            db = mv.synth_fill_host(dtype, 1000 + case, 0, n, d)
            if case % 3 == 0 and n > 40:                                  # duplicate rows: exact ties across tiles
                db[n // 2:n // 2 + 20] = db[:20]
            gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
            excluded = None
            if case % 2 == 1 and n > 10:
                excluded = rng.random(n) < 0.3
                gs.set_row_mask(excluded)
            q = mv.synth_fill_host(dtype, 2000 + case, 0, nq, d)
            if n > 3:
                q[0] = db[3]
            for metric in (oracle.L2, oracle.IP, oracle.COS):
                ids, sc, isc = gs.search(q, k, metric=metric)
                assert engine.profile_read()["last_path"] == 2, (case, dtype, n, d, nq, k)
                qq = q.astype(np.float32) if dtype == oracle.F16 else q
                sel = sorted(set([0, nq // 2, nq - 1]))
                if excluded is None:
                    check_batch(db, dtype, metric, qq[sel], k, ids[sel], sc[sel], isc[sel])
                else:
                    keep = np.flatnonzero(~excluded)
                    for qi in sel:
                        # the masked search must equal the oracle on the kept rows, reported with their original positions
                        o_ids, o_sc, o_isc = oracle.topk(db[keep], dtype, metric, qq[qi], min(k + 8, len(keep)))
                        got = ids[qi][ids[qi] != np.uint64(0xFFFFFFFFFFFFFFFF)].astype(np.int64)
                        assert not excluded[got].any(), (case, metric, qi)
                        kk = min(k, len(keep))
                        assert len(got) == kk
                        if dtype in (oracle.I8, oracle.U8):
                            assert np.array_equal(got, keep[o_ids[:kk].astype(np.int64)]), (case, metric, qi, got, keep[o_ids[:kk].astype(np.int64)])
                            if metric != oracle.COS:
                                assert np.array_equal(isc[qi][:kk], o_isc[:kk])
                            else:
                                assert np.array_equal(sc[qi][:kk].view(np.uint32), o_sc[:kk].view(np.uint32))
                        else:
                            if metric == oracle.L2:          # the tolerance is stated on d^2 (tests/util.py); near d = 0 the root amplifies it
                                qn2 = float((qq[qi].astype(np.float64) ** 2).sum())
                                xn2 = (db[got].astype(np.float64) ** 2).sum(1)
                                assert (np.abs(sc[qi][:kk].astype(np.float64) ** 2 - o_sc[:kk].astype(np.float64) ** 2)
                                        <= 2e-5 * (qn2 + xn2) + 1e-12).all(), (case, metric, qi, sc[qi][:kk], o_sc[:kk])
                            else:
                                assert np.allclose(sc[qi][:kk], o_sc[:kk], rtol=2e-4, atol=3e-5), (case, metric, qi)
            gs.free()
    finally:
        engine.set_path(0)
