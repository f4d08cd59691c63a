"""Reader / writer parity with the reference's own unit tests (CPU only, through the C ABI).

Each test names the reference test it re-expresses; paths are relative to
the reference repository.  Fixtures are the committed files under tests/golden/ (written by the
independent pure-Python writer in oracle/mvf_format.py) plus files written by the product's own
MvfBuilder.
"""
import os
import struct
import threading

import numpy as np
import pytest

import metrovector_b200 as mv
from oracle import mvf_format as fmt


def test_open_valid_file(golden_dir):                      # src/reader.rs:305-318
    r = mv.MvfReader.open(os.path.join(golden_dir, "test_space.mvf"))
    assert r.num_vector_spaces() == 1
    assert r.vector_space_names() == ["test_space"]


def test_reader_properties(golden_dir):                    # src/reader.rs:390-403
    r = mv.MvfFile.open(os.path.join(golden_dir, "test_space.mvf"))
    assert r.version() == 1 and not r.has_metadata() and r.file_size() > 0
    assert r.metadata_column_names() == []


def test_get_vector_space(golden_dir):                     # src/reader.rs:406-419
    r = mv.MvfReader.open(os.path.join(golden_dir, "test_space.mvf"))
    assert r.vector_space("test_space").name() == "test_space"
    with pytest.raises(mv.VectorSpaceNotFound) as e:
        r.vector_space("nonexistent")
    assert str(e.value) == "Vector space 'nonexistent' not found"      # src/errors.rs:18


def test_vector_space_properties(golden_dir):              # src/reader.rs:501-516
    s = mv.MvfReader.open(os.path.join(golden_dir, "test_space.mvf")).vector_space("test_space")
    assert (s.name(), s.dimension(), s.total_vectors()) == ("test_space", 4, 3)
    assert s.vector_type() == mv.VectorType.Dense
    assert s.distance_metric() == mv.DistanceMetric.L2 == mv.DistanceMetric.Euclidean
    assert s.data_type() == mv.DataType.Float32


def test_validate(golden_dir):                             # src/reader.rs:422-432,611-625
    for name in ("test_space", "simple_f32", "clustered_f32", "multi", "f16_refbug"):
        r = mv.MvfReader.open(os.path.join(golden_dir, name + ".mvf"))
        r.validate()
        r.validate_with_checksum()


def test_open_nonexistent_and_directory(tmp_path):         # src/reader.rs:339-345,628-637
    with pytest.raises(mv.Io):
        mv.MvfReader.open(str(tmp_path / "missing.mvf"))
    with pytest.raises(mv.Io):
        mv.MvfReader.open(str(tmp_path))


def test_open_small_and_malformed(tmp_path):               # src/reader.rs:321-387 + commented intents :518-608
    def write(name, data):
        p = tmp_path / name
        p.write_bytes(data)
        return str(p)
    with pytest.raises(mv.InvalidFormat) as e:
        mv.MvfReader.open(write("empty.mvf", b""))
    assert "File too small" in str(e.value)
    with pytest.raises(mv.InvalidFormat):
        mv.MvfReader.open(write("eight.mvf", b"MVF1" + b"MVF1"))                 # 8 bytes, test_open_minimum_valid_size
    with pytest.raises(mv.InvalidFormat) as e:
        mv.MvfReader.open(write("zeros.mvf", b"\x00" * 16))                      # test_open_invalid_footer_size input
    assert "Invalid magic bytes at start of file" in str(e.value)
    good = open(os.path.join(os.path.dirname(__file__), "golden", "test_space.mvf"), "rb").read()
    with pytest.raises(mv.InvalidFormat) as e:
        mv.MvfReader.open(write("endmagic.mvf", good[:-4] + b"XXXX"))
    assert "Invalid magic bytes at end of file" in str(e.value)
    with pytest.raises(mv.InvalidFormat) as e:                                   # the branch reader.rs:236-238 never tested upstream
        mv.MvfReader.open(write("footerlen.mvf", good[:-8] + struct.pack("<I", len(good)) + b"MVF1"))
    assert "Invalid footer length" in str(e.value)
    with pytest.raises(mv.InvalidFormat) as e:                                   # zero-length footer
        mv.MvfReader.open(write("nofooter.mvf", b"MVF1" + struct.pack("<I", 0) + b"MVF1"))
    assert "Failed to parse footer" in str(e.value)
    # garbage footers of every length must be rejected, never crash
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 8, 17, 64, 200):
        junk = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        with pytest.raises(mv.MvfError):
            mv.MvfReader.open(write("junk%d.mvf" % n, b"MVF1" + junk + struct.pack("<I", n) + b"MVF1"))
    # every single-byte corruption of a valid footer either opens or fails cleanly
    flen = struct.unpack_from("<I", good, len(good) - 8)[0]
    start = len(good) - 8 - flen
    for i in range(start, len(good) - 8):
        bad = bytearray(good)
        bad[i] ^= 0xFF
        try:
            r = mv.MvfReader.open(write("flip.mvf", bytes(bad)))
            for nm in r.vector_space_names():
                s = r.vector_space(nm)
                try:
                    s.get_vector(0)
                except mv.MvfError:
                    pass
        except mv.MvfError:
            pass


def test_unsupported_version(tmp_path):                    # commented test_unsupported_version, src/reader.rs:518-560
    rows = np.arange(8, dtype=np.float32).reshape(2, 4)
    data = fmt.build_mvf([dict(name="s", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows.tobytes())],
                         format_version=99)
    p = tmp_path / "v99.mvf"
    p.write_bytes(data)
    with pytest.raises(mv.UnsupportedVersion) as e:
        mv.MvfReader.open(str(p))
    assert str(e.value) == "Unsupported version: got 99, expected 1"   # src/errors.rs:15


def test_metadata_functionality(golden_dir):               # src/reader.rs:460-498
    r = mv.MvfReader.open(os.path.join(golden_dir, "multi.mvf"))
    assert r.has_metadata() and r.metadata_column_names() == ["ids", "labels"]
    assert r.num_vector_spaces() == 4
    s = r.vector_space("alpha")
    assert s.total_vectors() == 16 and s.distance_metric() == mv.DistanceMetric.Cosine


def test_get_vector_bounds_and_rows(golden_dir):           # src/vectors/vector_space.rs:386-397, mem.rs:683-696,769-790
    s = mv.MvfReader.open(os.path.join(golden_dir, "test_space.mvf")).vector_space("test_space")
    with pytest.raises(mv.IndexOutOfBounds) as e:
        s.get_vector(999)
    assert str(e.value) == "Index out of bounds: 999 >= 3"              # src/errors.rs:21
    v = s.get_vector(0)
    assert v.dimension() == 4 and v.data_type() == mv.DataType.Float32 and len(v.as_bytes()) == 16
    assert v.as_f32().tolist() == [1.0, 2.0, 3.0, 4.0]
    assert s.get_vector(2).as_f32().tolist() == [9.0, 10.0, 11.0, 12.0]


def test_map_vector_range(golden_dir):                     # src/vectors/mem.rs:543-592
    s = mv.MvfReader.open(os.path.join(golden_dir, "test_space.mvf")).vector_space("test_space")
    assert s.map_vector_range(0, 2).tolist() == [[1, 2, 3, 4], [5, 6, 7, 8]]
    with pytest.raises(mv.IndexOutOfBounds):
        s.map_vector_range(0, s.total_vectors() + 1)
    assert s.map_vector_range(1, 0).shape == (0, 4)


def test_clone_concurrent_and_threads(golden_dir):         # mem.rs:594-612, reader.rs:435-441 (Send + Sync)
    r = mv.MvfReader.open(os.path.join(golden_dir, "clustered_f32.mvf"))
    s = r.vector_space("clustered_data")
    c = s.clone_concurrent()
    assert c.get_vector(7).as_bytes() == s.get_vector(7).as_bytes()
    want = [s.get_vector(i).as_bytes() for i in range(60)]
    errs = []

    def work():
        try:
            v = r.vector_space("clustered_data")
            for _ in range(20):
                for i in range(60):
                    assert v.get_vector(i).as_bytes() == want[i]
        except Exception as ex:      # noqa
            errs.append(ex)
    th = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_as_f32_float16_and_unsupported(golden_dir):       # src/vectors/vector.rs:228-250
    r = mv.MvfReader.open(os.path.join(golden_dir, "f16_refbug.mvf"))
    s = r.vector_space("half_space")
    assert s.get_vector(0).as_f32().tolist() == [3.140625, 2.71875, 1.0, -0.5]    # 0x4248, 0x4170 exactly
    u8 = mv.MvfReader.open(os.path.join(golden_dir, "multi.mvf")).vector_space("bytes_u8")
    assert len(u8.get_vector(0).as_bytes()) == 16
    with pytest.raises(mv.Build) as e:
        u8.get_vector(0).as_f32()
    assert "Cannot convert to f32" in str(e.value)


def test_float16_total_vectors_as_the_reference_writes_it(golden_dir):
    # src/builder.rs:476 halves total_vectors for Float16; the footer value is exposed unchanged and
    # the scan visits exactly the rows `for i in 0..total_vectors()` would
    s = mv.MvfReader.open(os.path.join(golden_dir, "f16_refbug.mvf")).vector_space("half_space")
    assert s.total_vectors() == 4 and s.scannable_vectors() == 4
    with pytest.raises(mv.IndexOutOfBounds):
        s.get_vector(4)


def test_python_parser_and_cpp_reader_agree(golden_dir):
    for name in ("test_space", "simple_f32", "clustered_f32", "multi", "f16_refbug"):
        data = open(os.path.join(golden_dir, name + ".mvf"), "rb").read()
        parsed = fmt.parse_mvf(data)
        r = mv.MvfReader.open(os.path.join(golden_dir, name + ".mvf"))
        assert r.vector_space_names() == [s["name"] for s in parsed["spaces"]]
        for si, sp in enumerate(parsed["spaces"]):
            s = r.vector_space(sp["name"])
            assert (s.dimension(), s.total_vectors(), int(s.data_type()), int(s.distance_metric())) == \
                (sp["dimension"], sp["total_vectors"], sp["data_type"], sp["distance_metric"])
            for i in range(sp["total_vectors"]):
                assert s.get_vector(i).as_bytes() == fmt.get_vector_bytes(data, parsed, si, i)


def test_builder_round_trip_all_dtypes(tmp_path):          # src/builder.rs tests, as byte-layout documentation
    rng = np.random.default_rng(3)
    b = mv.MvfBuilder()
    f32 = rng.standard_normal((7, 5)).astype(np.float32)
    b.add_vector_space("a", 5, mv.VectorType.Dense, mv.DistanceMetric.Cosine, mv.DataType.Float32)
    b.add_vectors("a", f32)
    b.add_vector_space("h", 6, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float16)
    h_in = rng.standard_normal((9, 6)).astype(np.float32)
    b.add_vectors("h", h_in)
    b.add_vector_space("u", 3, mv.VectorType.Dense, mv.DistanceMetric.Dot, mv.DataType.UInt8)
    u = rng.integers(0, 256, (4, 3), dtype=np.uint8)
    with pytest.raises(mv.Build):                          # src/builder.rs:192
        b.add_vectors("u", u.astype(np.float32))
    b.add_vectors_raw("u", u)
    with pytest.raises(mv.DimensionMismatch):              # src/builder.rs:171-176
        b.add_vectors("a", np.zeros((1, 4), dtype=np.float32))
    with pytest.raises(mv.VectorSpaceNotFound):
        b.add_vectors("zzz", f32)
    b.add_metadata_column("ids", mv.DataType.UInt32, struct.pack("<7I", *range(7)))
    path = str(tmp_path / "rt.mvf")
    b.save(path)
    r = mv.MvfReader.open(path)
    r.validate_with_checksum()
    assert r.vector_space_names() == ["a", "h", "u"] and r.metadata_column_names() == ["ids"]
    assert np.array_equal(r.vector_space("a").map_vector_range(0, 7), f32)
    hs = r.vector_space("h")
    assert hs.total_vectors() == 9                         # correct count (the reference would write 4)
    assert np.array_equal(hs.map_vector_range(0, 9), h_in.astype(np.float16))
    assert np.array_equal(r.vector_space("u").map_vector_range(0, 4), u)
    parsed = fmt.parse_mvf(open(path, "rb").read())        # the independent parser reads the C++ writer's bytes
    assert [s["total_vectors"] for s in parsed["spaces"]] == [7, 9, 4]
    # corrupt one payload byte -> checksum validation reports it, plain validate does not
    raw = bytearray(open(path, "rb").read())
    raw[10] ^= 1
    bad = str(tmp_path / "bad.mvf")
    open(bad, "wb").write(bytes(raw))
    rb = mv.MvfReader.open(bad)
    rb.validate()
    with pytest.raises(mv.CorruptedData):
        rb.validate_with_checksum()


def test_truncated_block_is_corrupted_data(tmp_path):
    # a footer whose block runs past the file: the reference panics in get_vector (vector_space.rs:120);
    # we return CorruptedData like validate() would (reader.rs:149-162)
    rows = np.arange(16, dtype=np.float32).reshape(4, 4)
    good = fmt.build_mvf([dict(name="s", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows.tobytes())])
    parsed = fmt.parse_mvf(good)
    flen = struct.unpack_from("<I", good, len(good) - 8)[0]
    footer = bytearray(good[len(good) - 8 - flen:len(good) - 8])
    # patch block.size (first u64 after offset in the only DataBlock) to something huge
    idx = footer.find(struct.pack("<QQ", 0, 64))
    assert idx >= 0 and parsed["blocks"][0]["size"] == 64
    footer[idx + 8:idx + 16] = struct.pack("<Q", 1 << 40)
    p = tmp_path / "trunc.mvf"
    p.write_bytes(good[:len(good) - 8 - flen] + bytes(footer) + good[-8:])
    r = mv.MvfReader.open(str(p))
    with pytest.raises(mv.CorruptedData):
        r.validate()
    with pytest.raises(mv.CorruptedData):
        r.vector_space("s").get_vector(0)


def test_tombstone_bitmap_round_trip(tmp_path):
    # TombstoneInfo{Bitmap} (schema/core.fbs:34-39, schema/types.fbs:35-39): the reference reserves it and never
    # writes it (src/builder.rs:486); our writer and reader implement the Bitmap format
    rows = np.arange(40 * 4, dtype=np.float32).reshape(40, 4)
    deleted = np.zeros(40, dtype=bool)
    deleted[[0, 7, 8, 31, 39]] = True
    b = mv.MvfBuilder()
    b.add_vector_space("live", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors("live", rows)
    b.add_vector_space("plain", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors("plain", rows[:3])
    b.set_tombstones("live", deleted)
    b.add_metadata_column("ids", mv.DataType.UInt32, struct.pack("<40I", *range(40)))
    path = str(tmp_path / "tomb.mvf")
    b.save(path)
    r = mv.MvfReader.open(path)
    r.validate_with_checksum()
    got = r.vector_space("live").tombstones()
    assert got is not None and np.array_equal(got[:40], deleted) and not got[40:].any()
    assert r.vector_space("plain").tombstones() is None
    assert np.array_equal(r.vector_space("live").map_vector_range(0, 40), rows)
    data = open(path, "rb").read()
    parsed = fmt.parse_mvf(data)
    assert parsed["spaces"][0]["tombstones"] == dict(format=1, data_block_index=3, deleted_count=5)
    assert parsed["spaces"][1]["tombstones"] is None
    # the independent Python writer produces the same bytes
    bits = np.packbits(deleted, bitorder="little").tobytes()
    py = fmt.build_mvf([dict(name="live", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows.tobytes(), tombstones=bits),
                        dict(name="plain", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows[:3].tobytes())],
                       metadata_columns=[dict(name="ids", data_type=4, data=struct.pack("<40I", *range(40)))])
    assert py == data


def test_vector_ids_tombstone_list_and_metadata_column_round_trip(tmp_path):
    # the rest of what the schema reserves around the scan: external ids
    # (vector_ids_block_index, schema/core.fbs:54; src/builder.rs:61,257-271), TombstoneInfo{SortedList}
    # (schema/types.fbs:35-39) and metadata column bytes (schema/core.fbs:16-25).  The reference writes none of
    # them (src/builder.rs:483-486) and reads back only column names (src/reader.rs:132-143).
    rows = np.arange(30 * 4, dtype=np.float32).reshape(30, 4)
    ids = (np.arange(30, dtype=np.uint64) * 1000003 + (1 << 40))
    gone = np.array([2, 3, 17, 29], dtype=np.uint64)
    price = np.linspace(0.0, 29.0, 30).astype(np.float32)
    b = mv.MvfBuilder()
    b.add_vector_space("a", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors("a", rows)
    b.set_vector_ids("a", ids)
    b.set_tombstone_list("a", gone)
    b.add_vector_space("b", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors("b", rows[:5])
    b.add_metadata_column("price", mv.DataType.Float32, price.tobytes())
    with pytest.raises(mv.DimensionMismatch):
        b.set_vector_ids("b", ids)                             # 30 ids for 5 rows
    with pytest.raises(mv.Build):
        b.set_tombstone_list("b", np.array([3, 3], dtype=np.uint64))   # not strictly ascending
    path = str(tmp_path / "ids.mvf")
    b.save(path)
    r = mv.MvfReader.open(path)
    r.validate_with_checksum()
    a = r.vector_space("a")
    assert np.array_equal(a.vector_ids(), ids) and r.vector_space("b").vector_ids() is None
    assert np.array_equal(a.tombstone_list(), gone) and a.tombstones() is None
    assert r.vector_space("b").tombstone_list() is None
    raw, dt = r.metadata_column("price")
    assert dt == mv.DataType.Float32 and np.array_equal(np.frombuffer(raw, dtype=np.float32), price)
    with pytest.raises(mv.Build):
        r.metadata_column("nope")
    excl = mv.filter_range_bitmap(raw, dt, 10.0, 19.5)
    assert np.array_equal(excl, ~((price >= 10.0) & (price <= 19.5)))
    # independent Python writer: same bytes; independent parser: same indices
    data = open(path, "rb").read()
    py = fmt.build_mvf([dict(name="a", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows.tobytes(),
                             vector_ids=[int(x) for x in ids], tombstone_list=[int(x) for x in gone]),
                        dict(name="b", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows[:5].tobytes())],
                       metadata_columns=[dict(name="price", data_type=0, data=price.tobytes())])
    assert py == data
    parsed = fmt.parse_mvf(data)
    assert parsed["spaces"][0]["tombstones"] == dict(format=2, data_block_index=3, deleted_count=4)
    assert parsed["spaces"][0]["vector_ids_block_index"] == 4 and parsed["spaces"][1]["vector_ids_block_index"] == 0


def test_side_tables_are_rechecked_against_the_rows_at_save_time(tmp_path):
    # ids / tombstones were only validated when they were set; rows added afterwards produced a file every reader
    # (ours included) rejects as CorruptedData
    rows = np.arange(12 * 4, dtype=np.float32).reshape(12, 4)
    b = mv.MvfBuilder()
    b.add_vector_space("a", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    b.add_vectors("a", rows[:8])
    b.set_vector_ids("a", np.arange(8, dtype=np.uint64))
    b.add_vectors("a", rows[8:])                                # 12 rows now, 8 ids
    with pytest.raises(mv.DimensionMismatch):
        b.save(str(tmp_path / "bad.mvf"))
    b.set_vector_ids("a", np.arange(12, dtype=np.uint64))
    b.save(str(tmp_path / "good.mvf"))
    assert mv.MvfReader.open(str(tmp_path / "good.mvf")).vector_space("a").vector_ids().tolist() == list(range(12))
    c = mv.MvfBuilder()
    c.add_vector_space("a", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    c.add_vectors("a", rows[:8])
    c.set_tombstones("a", np.zeros(8, dtype=bool))
    c.add_vectors("a", rows)                                    # 20 rows, an 8-bit bitmap
    with pytest.raises(mv.DimensionMismatch):
        c.save(str(tmp_path / "bad2.mvf"))
    with pytest.raises(mv.Io):
        b.save(str(tmp_path / "no_such_dir" / "x.mvf"))


@pytest.mark.parametrize("dt,np_dt", [(mv.DataType.UInt8, np.uint8), (mv.DataType.Int8, np.int8), (mv.DataType.UInt32, np.uint32),
                                      (mv.DataType.UInt64, np.uint64), (mv.DataType.Float16, np.float16)])
def test_filter_range_bitmap_every_dtype(dt, np_dt):
    rng = np.random.default_rng(int(dt))
    col = (rng.integers(-100, 100, 1003) if np_dt in (np.int8, np.float16) else rng.integers(0, 200, 1003)).astype(np_dt)
    excl = mv.filter_range_bitmap(col, dt, 20, 90)
    assert np.array_equal(excl, ~((col.astype(np.float64) >= 20) & (col.astype(np.float64) <= 90)))


def test_footer_corruption_never_crashes_the_auxiliary_accessors(tmp_path):
    # the same single-byte sweep as test_invalid_footer_length..., over a footer that names an ids block, a
    # tombstone list and a metadata column: every accessor either answers or raises an MvfError
    rows = np.arange(12 * 4, dtype=np.float32).reshape(12, 4)
    good = fmt.build_mvf([dict(name="a", dimension=4, vector_type=0, distance_metric=0, data_type=0, data=rows.tobytes(),
                               vector_ids=list(range(100, 112)), tombstone_list=[1, 5]),
                          dict(name="b", dimension=4, vector_type=0, distance_metric=2, data_type=0, data=rows[:3].tobytes(),
                               tombstones=b"\x05")],
                         metadata_columns=[dict(name="price", data_type=0, data=rows[:, 0].tobytes())])
    flen = struct.unpack_from("<I", good, len(good) - 8)[0]
    start = len(good) - 8 - flen
    path = str(tmp_path / "flip.mvf")
    opened = 0
    for i in range(start, len(good) - 8):
        for mask in (0xFF, 0x01, 0x80):
            bad = bytearray(good)
            bad[i] ^= mask
            open(path, "wb").write(bytes(bad))
            try:
                r = mv.MvfReader.open(path)
            except mv.MvfError:
                continue
            opened += 1
            for call in (r.validate, r.validate_with_checksum, lambda: r.metadata_column("price")):
                try:
                    call()
                except mv.MvfError:
                    pass
            for nm in r.vector_space_names():
                try:
                    s = r.vector_space(nm)
                except mv.MvfError:
                    continue
                for call in (s.vector_ids, s.tombstone_list, s.tombstones, s.scannable_vectors, lambda: s.get_vector(0)):
                    try:
                        call()
                    except mv.MvfError:
                        pass
    assert opened > 0


def test_random_files_round_trip_and_match_the_independent_writer(tmp_path):
    # property test (seeded, no external generator needed): random spaces x dtypes x sizes x optional tombstones / ids /
    # metadata; the product's C++ writer, the independent Python writer and both readers must agree on every byte
    rng = np.random.default_rng(20261018)
    np_dt = {0: np.float32, 1: np.float16, 2: np.int8, 3: np.uint8}
    for trial in range(40):
        b = mv.MvfBuilder()
        spec, cols = [], []
        for si in range(int(rng.integers(1, 4))):
            dt = int(rng.integers(0, 4))
            dim = int(rng.choice([1, 3, 4, 16, 33, 128]))
            n = int(rng.integers(0, 50))
            rows = (rng.standard_normal((n, dim)) * 3).astype(np_dt[dt]) if dt < 2 else rng.integers(0, 120, (n, dim)).astype(np_dt[dt])
            name = "s%d_%d" % (trial, si)
            metric = int(rng.integers(0, 3))
            b.add_vector_space(name, dim, mv.VectorType.Dense, mv.DistanceMetric(metric), mv.DataType(dt))
            b.add_vectors_raw(name, rows)
            d = dict(name=name, dimension=dim, vector_type=0, distance_metric=metric, data_type=dt, data=rows.tobytes())
            extra = int(rng.integers(0, 4)) if n else 0
            if extra == 1:
                dele = rng.random(n) < 0.3
                if dele.any():
                    b.set_tombstones(name, dele)
                    d["tombstones"] = np.packbits(dele, bitorder="little").tobytes()
            elif extra == 2:
                gone = np.flatnonzero(rng.random(n) < 0.3).astype(np.uint64)
                if len(gone):
                    b.set_tombstone_list(name, gone)
                    d["tombstone_list"] = [int(x) for x in gone]
            if n and rng.random() < 0.4:
                ids = rng.integers(0, 1 << 62, n, dtype=np.uint64)
                b.set_vector_ids(name, ids)
                d["vector_ids"] = [int(x) for x in ids]
            spec.append((name, rows, d))
        for ci in range(int(rng.integers(0, 3))):
            data = rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes()
            b.add_metadata_column("c%d" % ci, mv.DataType.UInt8, data)
            cols.append(dict(name="c%d" % ci, data_type=3, data=data))
        path = str(tmp_path / ("r%d.mvf" % trial))
        b.save(path)
        raw = open(path, "rb").read()
        assert raw == fmt.build_mvf([d for _, _, d in spec], metadata_columns=cols), trial
        r = mv.MvfReader.open(path)
        r.validate_with_checksum()
        parsed = fmt.parse_mvf(raw)
        assert [s["name"] for s in parsed["spaces"]] == r.vector_space_names() == [nm for nm, _, _ in spec]
        for i, (name, rows, d) in enumerate(spec):
            s = r.vector_space(name)
            assert s.total_vectors() == len(rows) == parsed["spaces"][i]["total_vectors"]
            if len(rows):
                assert np.array_equal(s.map_vector_range(0, len(rows)), rows)
                assert fmt.get_vector_bytes(raw, parsed, i, len(rows) - 1) == rows[-1].tobytes()
            ids = s.vector_ids()
            assert (ids is None) == ("vector_ids" not in d) and (ids is None or ids.tolist() == d["vector_ids"])
            lst = s.tombstone_list()
            assert (lst is None) == ("tombstone_list" not in d) and (lst is None or lst.tolist() == d["tombstone_list"])
            assert (s.tombstones() is None) == ("tombstones" not in d)
        for c in cols:
            assert r.metadata_column(c["name"])[0] == c["data"]


def test_structured_error_fields_cross_the_abi(golden_dir):
    # IndexOutOfBounds{index, len} / DimensionMismatch{expected, actual} (src/errors.rs:21-25) keep their numbers
    # through the C ABI (mvf_last_error_fields), so a binding rebuilds the exact variant
    r = mv.MvfReader.open(os.path.join(golden_dir, "simple_f32.mvf"))
    sp = r.vector_space("embeddings")
    with pytest.raises(mv.IndexOutOfBounds) as ei:
        sp.get_vector(999)                                        # src/vectors/vector_space.rs:385-397
    assert ei.value.fields == (999, sp.total_vectors())
    with pytest.raises(mv.VectorSpaceNotFound) as ei:
        r.vector_space("nope")
    assert ei.value.fields == (0, 0)
    b = mv.MvfBuilder()
    b.add_vector_space("a", 4, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
    with pytest.raises(mv.DimensionMismatch) as ei:
        b.add_vectors("a", np.zeros((2, 5), dtype=np.float32))
    assert ei.value.fields == (4, 5)
