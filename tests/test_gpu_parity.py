"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Bar: bit-exact ids and scores for Int8/UInt8 (ties -> lowest index); floats within the tolerances
written in tests/util.py.  All run on cuda:0 of a B200.
"""
import json
import os

import numpy as np
import pytest

import metrovector_b200 as mv
import oracle
from util import INVALID, as_query, check_batch, check_one, f32_bits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {oracle.F32: mv.DataType.Float32, oracle.F16: mv.DataType.Float16, oracle.I8: mv.DataType.Int8, oracle.U8: mv.DataType.UInt8}


def test_native_path_is_loaded(engine):
    info = engine.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100
    assert os.path.exists(mv.LIB_PATH)


# ---- known answers from the reference's example programs, end to end through a file ------------
def test_simple_rs_through_file(engine, golden_dir):
    r = mv.MvfReader.open(os.path.join(golden_dir, "simple_f32.mvf"))
    gs = engine.upload(r.vector_space("embeddings"))
    r.close()                                   # the HBM copy outlives the mapping
    top = mv.find_top_k_similar(gs, np.array([2.5, 4.5, 6.5, 8.5], dtype=np.float32), 3)
    assert [t.index for t in top] == [3, 1, 4]  # ids 1 and 4 tie at 3.0 exactly -> lowest index first
    assert [f32_bits(t.score) for t in top] == [0x3f800000, 0x40400000, 0x40400000]
    top1 = mv.find_top_k_similar(gs, np.array([2.5, 4.5, 6.5, 8.5], dtype=np.float32), 1)
    assert top1[0].index == 3                   # examples/simple.rs:73-94


def test_similarity_search_rs_through_file(engine, golden_dir):
    with open(os.path.join(golden_dir, "known_answers.json")) as fh:
        known = json.load(fh)
    r = mv.MvfReader.open(os.path.join(golden_dir, "clustered_f32.mvf"))
    space = r.vector_space("clustered_data")
    gs = engine.upload(space)
    db = space.map_vector_range(0, 60)
    assert np.array_equal(gs.read_rows(0, 60), db)
    for c in known["clustered"]:
        q = np.array(c["query"], dtype=np.float32)
        ids, sc, isc = gs.search(q, c["k"])
        assert ids[0].tolist() == c["ids"]      # the documented intent: k nearest, not the as-written farthest
        assert ids[0].tolist() != c["as_written_ids"]
        want = np.array(c["score_bits"], dtype=np.uint32).view(np.float32)
        assert np.allclose(sc[0], want, rtol=0, atol=4e-7)    # d=4: FMA vs mul+add, at most ~1 ulp of d^2
        check_one(db, oracle.F32, oracle.L2, q, c["k"], ids[0], sc[0], isc[0])


def test_seeded_golden_cases(engine, golden_dir):
    with open(os.path.join(golden_dir, "known_answers.json")) as fh:
        known = json.load(fh)
    cache = {}
    for c in known["seeded"]:
        key = (c["dtype"], c["seed"], c["n"], c["d"])
        if key not in cache:
            db = mv.synth_fill_host(c["dtype"], c["seed"], 0, c["n"], c["d"])
            cache = {key: (db, engine.from_host(db, DT[c["dtype"]], mv.DistanceMetric.L2))}
        db, gs = cache[key]
        q = mv.synth_fill_host(c["dtype"], c["qseed"], 0, 2, c["d"])[c["qrow"]]
        if c["planted"] is not None:
            q = db[c["planted"]]
        q = as_query(c["dtype"], q)
        ids, sc, isc = gs.search(q, c["k"], metric=c["metric"])
        if c["dtype"] in (oracle.I8, oracle.U8):
            assert ids[0].tolist() == c["ids"]
            if c["metric"] != oracle.COS:
                assert isc[0].tolist() == c["iscores"]
            else:
                assert sc[0].view(np.uint32).tolist() == c["score_bits"]
        else:
            check_one(db, c["dtype"], c["metric"], q, c["k"], ids[0], sc[0], isc[0])
            if c["planted"] is not None and c["metric"] == oracle.L2:
                assert ids[0][0] == c["planted"] and sc[0][0] == 0.0


# ---- sweep: dtype x metric x shapes (narrow rows, wide rows, ragged dims, k > n, batches) --------
SHAPES = [  # (n, d)
    (1, 1), (5, 3), (31, 4), (257, 7), (1000, 16), (999, 33), (2048, 64), (1500, 100), (3000, 128),
    (1200, 200), (2500, 768), (700, 1000), (40, 2048),
]


@pytest.mark.parametrize("dtype", [oracle.F32, oracle.F16, oracle.I8, oracle.U8])
@pytest.mark.parametrize("metric", [oracle.L2, oracle.IP, oracle.COS])
def test_sweep_shapes(engine, dtype, metric):
    for si, (n, d) in enumerate(SHAPES):
        db = mv.synth_fill_host(dtype, 100 + si, 0, n, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric(metric))
        assert gs.distance_metric == metric and gs.local_n == n
        for nq, k in ((1, 10), (3, 1), (8, 5), (9, 40), (2, 100)):
            qs = mv.synth_fill_host(dtype, 200 + si, 0, nq, d)
            qs[0] = db[n // 2]                                   # a query equal to a stored row
            queries = np.stack([as_query(dtype, r) for r in qs])
            ids, sc, isc = gs.search(queries, k)                  # metric = the space's default
            assert ids.shape == (nq, k)
            check_batch(db, dtype, metric, queries, k, ids, sc, isc)
            if k > n:
                assert (ids[:, n:] == INVALID).all()
        gs.free()


@pytest.mark.parametrize("k", [128, 200, 300, 752, 800, 1024])
def test_large_k_exercises_every_merge_flavour(engine, k):
    # many per-CTA lists x large k: single tournament (k=128), segmented tournament (200..752),
    # block-bitonic rounds (k > 752)
    for dtype, n, d in ((oracle.U8, 120_000, 16), (oracle.F32, 60_000, 40)):
        db = mv.synth_fill_host(dtype, 61, 0, n, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
        q = np.stack([as_query(dtype, r) for r in mv.synth_fill_host(dtype, 62, 0, 3, d)])
        for metric in (oracle.L2, oracle.COS):
            ids, sc, isc = gs.search(q, k, metric=metric)
            check_batch(db, dtype, metric, q, k, ids, sc, isc)
        gs.free()


def test_shard_of_a_space_beyond_2_32_rows(engine):
    # candidates carry 32-bit LOCAL row ids; the shard's first global row is added in 64 bits when ids are emitted, so a
    # resident shard may sit anywhere in a space of more than 2^32 rows (both kernels, every merge flavour's emit)
    big = (1 << 33) + 12345
    for dtype, n, d, nq, k in ((oracle.U8, 30_000, 64, 3, 10), (oracle.F16, 20_000, 96, 40, 10), (oracle.F32, 9_000, 32, 2, 300),
                               (oracle.U8, 50_000, 128, 70, 33)):
        db = mv.synth_fill_host(dtype, 71, 0, n, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2, row_base=big, global_n=big + n + 7)
        assert (gs.row_base, gs.global_n) == (big, big + n + 7)
        q = mv.synth_fill_host(dtype, 72, 0, nq, d)
        queries = q if (dtype == oracle.F16 and nq >= 4) else np.stack([as_query(dtype, r) for r in q])
        ids, sc, isc = gs.search(queries, k)
        assert (ids >= big).all() and (ids < big + n).all()
        check_batch(db, dtype, oracle.L2, queries.astype(np.float32) if dtype == oracle.F16 else queries, k,
                    ids - np.uint64(big), sc, isc)
        gs.free()


def test_ties_resolve_to_lowest_index(engine):
    # 3 distinct byte values -> thousands of exact ties in every metric
    rng = np.random.default_rng(5)
    for dtype, lo, hi in ((oracle.U8, 0, 3), (oracle.I8, -1, 2)):
        for d in (16, 128, 144, 512):
            db = rng.integers(lo, hi, (5000, d)).astype(oracle.np_dtype(dtype))
            gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
            q = rng.integers(lo, hi, (4, d)).astype(oracle.np_dtype(dtype))
            for metric in (oracle.L2, oracle.IP, oracle.COS):
                ids, sc, isc = gs.search(q, 64, metric=metric)
                check_batch(db, dtype, metric, q, 64, ids, sc, isc)
    # float ties: duplicated rows
    base = mv.synth_fill_host(oracle.F32, 1, 0, 50, 128)
    db = np.concatenate([base] * 8)
    gs = engine.from_host(db, mv.DataType.Float32, mv.DistanceMetric.L2)
    ids, sc, _ = gs.search(base[7], 8)
    assert ids[0].tolist() == [7 + 50 * j for j in range(8)] and (sc[0] == 0).all()


def test_cosine_zero_norm_rows_and_query(engine):
    db = mv.synth_fill_host(oracle.F32, 3, 0, 100, 64)
    db[10] = 0
    db[20] = 0
    gs = engine.from_host(db, mv.DataType.Float32, mv.DistanceMetric.Cosine)
    q = mv.synth_fill_host(oracle.F32, 4, 0, 1, 64)[0]
    ids, sc, isc = gs.search(q, 100)
    check_one(db, oracle.F32, oracle.COS, q, 100, ids[0], sc[0], isc[0])
    pos = {int(i): float(s) for i, s in zip(ids[0], sc[0])}
    assert pos[10] == 1.0 and pos[20] == 1.0
    ids, sc, _ = gs.search(np.zeros(64, dtype=np.float32), 5)      # zero query: every distance is 1 -> ids 0..4
    assert ids[0].tolist() == [0, 1, 2, 3, 4] and (sc[0] == 1.0).all()


def test_f16_space_accepts_f16_and_f32_queries(engine):
    db = mv.synth_fill_host(oracle.F16, 8, 0, 2000, 256)
    gs = engine.from_host(db, mv.DataType.Float16, mv.DistanceMetric.Cosine)
    q16 = mv.synth_fill_host(oracle.F16, 9, 0, 3, 256)
    a = gs.search(q16, 10)
    b = gs.search(q16.astype(np.float32), 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    check_batch(db, oracle.F16, oracle.COS, q16.astype(np.float32), 10, *a)


def test_empty_space_and_argument_errors(engine):
    empty = engine.from_host(np.zeros((0, 8), dtype=np.float32), mv.DataType.Float32, mv.DistanceMetric.L2)
    ids, sc, _ = empty.search(np.ones(8, dtype=np.float32), 4)
    assert (ids == INVALID).all() and np.isnan(sc).all()
    db = mv.synth_fill_host(oracle.U8, 1, 0, 64, 32)
    gs = engine.from_host(db, mv.DataType.UInt8, mv.DistanceMetric.L2)
    q = db[:1]
    with pytest.raises(mv.Build):
        gs.search(q, 0)
    with pytest.raises(mv.Build):
        gs.search(q, 1025)
    with pytest.raises(mv.DimensionMismatch):
        gs.search(np.zeros((1, 31), dtype=np.uint8), 3)
    with pytest.raises(mv.Build):
        gs.search(q, 3, metric=7)
    with pytest.raises(mv.Build):                                  # Custom metric is not searchable
        engine.from_host(db, mv.DataType.UInt8, mv.DistanceMetric.Custom)
    ids, _, isc = gs.search(q, 3)                                  # still usable after errors
    assert ids[0][0] == 0 and isc[0][0] == 0


# ---- residency ---------------------------------------------------------------------------------
def test_upload_from_file_matches_mapping_and_oracle_norms(engine, golden_dir, tmp_path, monkeypatch):
    # a file with several spaces: later blocks start at odd offsets (only 4-byte alignment is guaranteed)
    b = mv.MvfBuilder()
    spaces = {}
    for name, dtype, d, n, metric in (("odd", oracle.U8, 13, 301, mv.DistanceMetric.L2),
                                      ("f", oracle.F32, 128, 4000, mv.DistanceMetric.Cosine),
                                      ("h", oracle.F16, 100, 3000, mv.DistanceMetric.Dot),
                                      ("i", oracle.I8, 64, 5000, mv.DistanceMetric.Cosine)):
        rows = mv.synth_fill_host(dtype, 77, 0, n, d)
        spaces[name] = (dtype, rows, metric)
        b.add_vector_space(name, d, mv.VectorType.Dense, metric, DT[dtype])
        b.add_vectors_raw(name, rows)
    path = str(tmp_path / "multi.mvf")
    b.save(path)
    for no_register in ("", "1"):
        if no_register:
            monkeypatch.setenv("MVF_NO_HOST_REGISTER", "1")
        r = mv.MvfReader.open(path)
        for name, (dtype, rows, metric) in spaces.items():
            gs = engine.upload(r.vector_space(name))
            assert gs.upload_stats()["path"] in ((2,) if no_register else (1, 2))
            assert gs.local_n == rows.shape[0] and gs.distance_metric == metric
            assert np.array_equal(gs.read_rows(0, rows.shape[0]).view(np.uint8), rows.view(np.uint8))
            norm, sq = gs.read_norms()
            o_norm, o_sqf, o_sqi = oracle.row_norms(rows, dtype)
            if dtype in (oracle.I8, oracle.U8):
                assert np.array_equal(sq.view(np.int32), o_sqi) and np.array_equal(norm, o_norm)
            else:
                assert np.allclose(sq, o_sqf, rtol=2e-6) and np.allclose(norm, o_norm, rtol=2e-6)
            q = as_query(dtype, rows[17])
            ids, sc, isc = gs.search(q, 10)
            check_one(rows, dtype, int(metric), q, 10, ids[0], sc[0], isc[0])
        r.close()


def test_reference_written_f16_file_scans_what_the_reference_would(engine, golden_dir):
    # total_vectors is halved by the reference's builder (src/builder.rs:476): its loop visits rows 0..3 only
    r = mv.MvfReader.open(os.path.join(golden_dir, "f16_refbug.mvf"))
    gs = engine.upload(r.vector_space("half_space"))
    assert gs.local_n == 4 and gs.global_n == 4
    ids, _, _ = gs.search(np.array([0.3, -0.6, 0.9, 1.0], dtype=np.float32), 8)
    assert set(ids[0][:4].tolist()) == {0, 1, 2, 3} and (ids[0][4:] == INVALID).all()


def test_device_generator_matches_host_generator(engine):
    for dtype in (oracle.F32, oracle.F16, oracle.I8, oracle.U8):
        for d in (4, 13, 128, 770):
            gs = engine.synthetic(DT[dtype], mv.DistanceMetric.L2, 1000, d, seed=0x4D564631)
            host = mv.synth_fill_host(dtype, 0x4D564631, 0, 1000, d)
            assert np.array_equal(gs.read_rows(0, 1000).view(np.uint8), host.view(np.uint8))
            assert np.array_equal(host.view(np.uint8), oracle.synth(dtype, 0x4D564631, 0, 1000, d).view(np.uint8))
            gs.free()


def test_get_vectors_returns_the_rows_behind_a_result(engine):
    # ScoredVector.vector (examples/similarity_search.rs:18): the rows of the top-k, from the HBM copy
    for dtype, d in ((oracle.F32, 100), (oracle.F16, 768), (oracle.U8, 13), (oracle.I8, 128)):
        db = mv.synth_fill_host(dtype, 71, 0, 3000, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
        ids, _, _ = gs.search(as_query(dtype, db[5]), 7)
        rows = gs.get_vectors(ids[0])
        assert np.array_equal(rows.view(np.uint8), db[ids[0].astype(np.int64)].view(np.uint8))
        if dtype in (oracle.F32, oracle.F16):
            assert np.array_equal(gs.get_vectors(ids[0], as_f32=True), db[ids[0].astype(np.int64)].astype(np.float32))
        else:
            with pytest.raises(mv.Build):
                gs.get_vectors(ids[0], as_f32=True)
        pad = gs.get_vectors(np.array([2999, 3000, 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
        assert np.array_equal(pad[0], db[2999]) and not pad[1:].any()
        gs.free()


def test_device_crc32_matches_zlib_and_the_block_manifest(engine, tmp_path):
    import zlib
    # the finished intent of validate_with_checksum (src/reader.rs:172-221), on the resident copy
    for dtype, n, d in ((oracle.U8, 1, 16), (oracle.U8, 7, 16), (oracle.F32, 1000, 4), (oracle.F16, 4097, 200), (oracle.I8, 300_001, 48),
                        (oracle.F32, 20_000, 768)):
        db = mv.synth_fill_host(dtype, 81, 0, n, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
        crc, ms = gs.checksum()
        assert crc == (zlib.crc32(db.tobytes()) & 0xFFFFFFFF), (dtype, n, d)
        gs.free()
    # ragged rows (row bytes not a multiple of 16) are padded in HBM: the kernel hashes one row per thread and skips the padding
    for dtype, n, d in ((oracle.U8, 10, 13), (oracle.F16, 5003, 100), (oracle.I8, 70_001, 129), (oracle.F32, 999, 5)):
        db = mv.synth_fill_host(dtype, 83, 0, n, d)
        gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
        assert gs.checksum()[0] == (zlib.crc32(db.tobytes()) & 0xFFFFFFFF), (dtype, n, d)
        gs.free()
    # a shard of a larger space cannot vouch for the whole block without a communicator
    part = engine.from_host(mv.synth_fill_host(oracle.U8, 1, 0, 10, 16), mv.DataType.UInt8, mv.DistanceMetric.L2, row_base=5, global_n=40)
    with pytest.raises(mv.Build):
        part.checksum()
    part.free()
    rows = mv.synth_fill_host(oracle.F32, 82, 0, 5000, 64)
    b = mv.MvfBuilder()
    b.add_vector_space("v", 64, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors_raw("v", rows)
    good = str(tmp_path / "good.mvf")
    b.save(good)
    r = mv.MvfReader.open(good)
    sp = r.vector_space("v")
    gs = engine.upload(sp)
    gs.verify_checksum(sp)
    raw = bytearray(open(good, "rb").read())
    raw[4 + 12345] ^= 0x40                                          # one flipped bit in the vector block
    bad = str(tmp_path / "bad.mvf")
    open(bad, "wb").write(bytes(raw))
    rb = mv.MvfReader.open(bad)
    spb = rb.vector_space("v")
    gb = engine.upload(spb)
    with pytest.raises(mv.CorruptedData) as e:
        gb.verify_checksum(spb)
    assert "Block checksum mismatch" in str(e.value)


def _masked_oracle(db, dtype, metric, q, k, excluded):
    keep = np.flatnonzero(~excluded)
    ids, sc, isc = oracle.topk(db[keep], dtype, metric, q, k)
    return keep[ids.astype(np.int64)], sc, isc


@pytest.mark.parametrize("dtype,d,n", [(oracle.F32, 768, 4000), (oracle.U8, 128, 30000), (oracle.F16, 96, 9000), (oracle.I8, 16, 5000)])
def test_row_mask_excludes_rows_on_both_paths(engine, dtype, d, n):
    # tombstones / pre-filters as a key mask: masked rows never appear
    rng = np.random.default_rng(3)
    db = mv.synth_fill_host(dtype, 95, 0, n, d)
    excluded = rng.random(n) < 0.3
    gs = engine.from_host(db, DT[dtype], mv.DistanceMetric.L2)
    k = 10
    q = mv.synth_fill_host(dtype, 96, 0, 70, d)
    q[0] = db[np.flatnonzero(excluded)[0]]                 # a query equal to a deleted row must not find it
    try:
        for path in (1, 2):
            if path == 2 and dtype == oracle.F32:
                continue
            engine.set_path(path)
            qq = q if path == 2 else np.stack([as_query(dtype, r) for r in q[:5]])
            gs.set_row_mask(excluded)
            ids, sc, isc = gs.search(qq, k)
            for qi in (0, 1, 2, len(qq) - 1):
                query = qq[qi].astype(np.float32) if dtype in (oracle.F32, oracle.F16) else qq[qi]
                o_ids, o_sc, o_isc = _masked_oracle(db, dtype, oracle.L2, query, k, excluded)
                assert not excluded[ids[qi].astype(np.int64)].any()
                if dtype in (oracle.I8, oracle.U8):
                    assert ids[qi].tolist() == o_ids.tolist() and isc[qi].tolist() == o_isc.tolist()
                else:
                    assert ids[qi].tolist() == o_ids.tolist()
                    assert np.allclose(sc[qi], o_sc, rtol=1e-4, atol=1e-4)
            gs.set_row_mask(None)                              # cleared: the deleted row is found again
            ids, _, _ = gs.search(qq[:1], 1)
            assert ids[0][0] == np.flatnonzero(excluded)[0]
    finally:
        engine.set_path(0)
        gs.free()


def test_upload_applies_the_files_tombstones(engine, tmp_path):
    n, d = 2000, 64
    rows = mv.synth_fill_host(oracle.F32, 97, 0, n, d)
    deleted = np.zeros(n, dtype=bool)
    deleted[::3] = True
    b = mv.MvfBuilder()
    b.add_vector_space("v", d, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors_raw("v", rows)
    b.set_tombstones("v", deleted)
    path = str(tmp_path / "t.mvf")
    b.save(path)
    r = mv.MvfReader.open(path)
    gs = engine.upload(r.vector_space("v"))
    ids, sc, _ = gs.search(rows[0], 5)                         # row 0 is deleted
    o_ids, o_sc, _ = _masked_oracle(rows, oracle.F32, oracle.L2, rows[0], 5, deleted)
    assert ids[0].tolist() == o_ids.tolist() and ids[0][0] != 0


def test_external_ids_are_reported_and_ties_still_break_by_position(engine):
    # vector_ids_block_index (schema/core.fbs:54): results carry ids[row]; duplicated rows tie, the lower POSITION wins
    base = mv.synth_fill_host(oracle.U8, 11, 0, 3000, 128)
    db = np.concatenate([base, base])                              # every row twice: position p and p + 3000
    ext = (np.arange(len(db), dtype=np.uint64)[::-1] * 3 + (1 << 33)).copy()   # decreasing, above 2^32
    gs = engine.from_host(db, mv.DataType.UInt8, mv.DistanceMetric.L2)
    q = mv.synth_fill_host(oracle.U8, 12, 0, 40, 128)
    try:
        for path in (1, 2):
            engine.set_path(path)
            gs.set_ids(ext)
            ids, sc, isc = gs.search(q, 8)
            gs.set_ids(None)
            pos, _, isc0 = gs.search(q, 8)
            for qi in range(len(q)):
                o_ids, _, o_isc = oracle.topk(db, oracle.U8, oracle.L2, q[qi], 8)
                assert pos[qi].tolist() == o_ids.tolist() and isc[qi].tolist() == o_isc.tolist()
                assert ids[qi].tolist() == ext[o_ids.astype(np.int64)].tolist()
        with pytest.raises(mv.DimensionMismatch):
            gs.set_ids(ext[:10])
    finally:
        engine.set_path(0)
        gs.free()


def test_upload_applies_ids_and_tombstone_list_and_filters_add_to_them(engine, tmp_path):
    n, d = 3000, 32
    rows = mv.synth_fill_host(oracle.F32, 98, 0, n, d)
    ext = np.arange(n, dtype=np.uint64) * 10 + 7
    gone = np.arange(0, n, 5, dtype=np.uint64)                     # TombstoneInfo{SortedList}
    year = (1990 + np.arange(n) % 40).astype(np.uint32)            # a metadata column to pre-filter on
    b = mv.MvfBuilder()
    b.add_vector_space("v", d, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors_raw("v", rows)
    b.set_vector_ids("v", ext)
    b.set_tombstone_list("v", gone)
    b.add_metadata_column("year", mv.DataType.UInt32, year.tobytes())
    path = str(tmp_path / "full.mvf")
    b.save(path)
    r = mv.MvfReader.open(path)
    gs = engine.upload(r.vector_space("v"))
    deleted = np.zeros(n, dtype=bool)
    deleted[gone.astype(np.int64)] = True
    raw, dt = r.metadata_column("year")
    filt = mv.filter_range_bitmap(raw, dt, 2000, 2010)
    assert np.array_equal(filt, ~((year >= 2000) & (year <= 2010)))
    q = rows[0]                                                    # row 0 is deleted
    ids, _, _ = gs.search(q, 6)
    o_ids, _, _ = _masked_oracle(rows, oracle.F32, oracle.L2, q, 6, deleted)
    assert ids[0].tolist() == ext[o_ids].tolist()
    gs.set_row_mask(filt)                                          # filter AND tombstones
    ids, _, _ = gs.search(q, 6)
    o_ids, _, _ = _masked_oracle(rows, oracle.F32, oracle.L2, q, 6, deleted | filt)
    assert ids[0].tolist() == ext[o_ids].tolist()
    gs.set_row_mask(None)                                          # the filter goes, the deletions stay
    ids, _, _ = gs.search(q, 6)
    o_ids, _, _ = _masked_oracle(rows, oracle.F32, oracle.L2, q, 6, deleted)
    assert ids[0].tolist() == ext[o_ids].tolist()
    gs.set_tombstones(None)                                        # explicit un-delete: row 0 finds itself
    ids, _, _ = gs.search(q, 1)
    assert ids[0][0] == ext[0]
    gs.free()


def test_searches_from_several_host_threads_are_serialised_not_raced(engine):
    # one engine = one stream + one set of scratch buffers; the library locks per engine (SURVEY.md 8b "Threading"),
    # so concurrent callers (ctypes drops the GIL during the call) must all get the serial answers
    import threading
    gs = engine.synthetic(mv.DataType.UInt8, mv.DistanceMetric.L2, 300000, 128, seed=5)
    q = mv.synth_fill_host(oracle.U8, 6, 0, 48, 128)
    serial = [gs.search(q[i], 10) for i in range(len(q))]
    batch = gs.search(q[:16], 10)                                  # a tensor-path batch in the mix
    got, errs = [None] * len(q), []

    def work(t):
        try:
            for i in range(t, len(q), 4):
                got[i] = gs.search(q[i], 10)
                if i % 8 == 0:
                    b = gs.search(q[:16], 10)
                    assert np.array_equal(b[0], batch[0]) and np.array_equal(b[2], batch[2])
        except Exception as e:                                     # surfaced below, not swallowed
            errs.append(e)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for i in range(len(q)):
        assert np.array_equal(got[i][0], serial[i][0]) and np.array_equal(got[i][2], serial[i][2]), i
    gs.free()


def test_profile_counters(engine):
    gs = engine.synthetic(mv.DataType.Float32, mv.DistanceMetric.L2, 20000, 128, seed=1)
    engine.profile(True)
    q = mv.synth_fill_host(oracle.F32, 2, 0, 1, 128)
    for _ in range(3):
        gs.search(q, 10)
    st = engine.profile_read()
    engine.profile(False)
    # one launch per search: the scan kernel's last CTA merges the per-CTA lists itself (K1 + K3 fused)
    assert st["scan_launches"] == 3 and st["all_launches"] == 3 and st["scan_ms_total"] > 0 and st["last_path"] == 1
