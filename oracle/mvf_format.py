"""MVF wire format in pure Python -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

An independent restatement of the reference's on-disk format, used to (a) generate the golden
fixture files under tests/golden/ (tests/golden/make_golden.py) and (b) cross-check the product's
C++ reader and writer.  Nothing in metrovector_b200/ imports this module.

Reference (paths relative to /root/reference):
  file framing      src/builder.rs:417-557, schema/FORMAT.md:9-24, src/lib.rs:25-26
  block manifest    src/builder.rs:241-308   (offsets relative to the end of the start magic)
  footer tables     schema/mvf.fbs:12-30, schema/core.fbs:7-57, schema/index.fbs:6-16
  reader checks     src/reader.rs:225-278
The FlatBuffers encoder below is a from-scratch back-to-front builder that follows the public
FlatBuffers wire format (flatbuffers 25.2.10 is what Cargo.lock pins).  Field emission order
mimics flatc's Rust codegen convention (largest scalars first) but readers must not depend on it.
"""
import struct
import zlib

MAGIC = b"MVF1"

F32, F16, I8, U8 = 0, 1, 2, 3
ELEM_SIZE = {F32: 4, F16: 2, I8: 1, U8: 1}


class _FbBuilder:
    """Minimal back-to-front FlatBuffers builder (tables, strings, vectors, structs)."""

    def __init__(self):
        self.buf = bytearray()      # stored reversed-in-meaning: we prepend by inserting at 0
        self.minalign = 1
        self.vtables = []           # offsets (from end) of written vtables
        self._fields = None
        self._obj_start = 0

    def offset(self):
        return len(self.buf)

    def _pad(self, n):
        self.buf[0:0] = b"\x00" * n

    def prep(self, size, additional):
        if size > self.minalign:
            self.minalign = size
        align = (-(len(self.buf) + additional)) % size
        self._pad(align)

    def _prepend(self, data):
        self.buf[0:0] = data

    def prepend_scalar(self, fmt, v):
        size = struct.calcsize(fmt)
        self.prep(size, 0)
        self._prepend(struct.pack("<" + fmt, v))

    def prepend_uoffset(self, off):
        self.prep(4, 0)
        assert off <= self.offset()
        self._prepend(struct.pack("<I", self.offset() - off + 4))

    def create_string(self, s):
        b = s.encode("utf-8")
        self.prep(4, len(b) + 1)
        self._prepend(b"\x00")
        self._prepend(b)
        self._prepend(struct.pack("<I", len(b)))
        return self.offset()

    def create_offset_vector(self, offs):
        self.prep(4, 4 * len(offs))
        for o in reversed(offs):
            self.prepend_uoffset(o)
        self._prepend(struct.pack("<I", len(offs)))
        return self.offset()

    def create_struct_vector(self, structs, elem_size, align):
        self.prep(4, elem_size * len(structs))
        self.prep(align, elem_size * len(structs))
        for s in reversed(structs):
            assert len(s) == elem_size
            self._prepend(s)
        self._prepend(struct.pack("<I", len(structs)))
        return self.offset()

    def start_table(self, nslots):
        self._fields = [0] * nslots
        self._obj_start = self.offset()

    def slot_scalar(self, slot, fmt, v, default):
        if v == default:
            return
        self.prepend_scalar(fmt, v)
        self._fields[slot] = self.offset()

    def slot_offset(self, slot, off):
        if off is None:
            return
        self.prepend_uoffset(off)
        self._fields[slot] = self.offset()

    def end_table(self):
        self.prepend_scalar("i", 0)            # soffset placeholder
        table_off = self.offset()
        fields = self._fields
        while fields and fields[-1] == 0:
            fields = fields[:-1]
        vt = struct.pack("<HH", 4 + 2 * len(fields), table_off - self._obj_start)
        for f in fields:
            vt += struct.pack("<H", (table_off - f) if f else 0)
        # dedup
        existing = None
        for vo in self.vtables:
            start = len(self.buf) - vo
            if bytes(self.buf[start:start + len(vt)]) == vt and \
                    struct.unpack_from("<H", self.buf, start)[0] == len(vt):
                existing = vo
                break
        if existing is None:
            self.prep(2, len(vt))
            # vtable sits directly before (lower address than) the table
            self._prepend(vt)
            existing = self.offset()
            self.vtables.append(existing)
        # patch soffset: table_pos - vtable_pos, both as absolute positions from buffer start
        tpos = len(self.buf) - table_off
        vpos = len(self.buf) - existing
        struct.pack_into("<i", self.buf, tpos, tpos - vpos)
        self._fields = None
        return table_off

    def finish_minimal(self, root):
        self.prep(self.minalign, 4)
        self.prepend_uoffset(root)
        return bytes(self.buf)


def _data_block(offset, size, compression, compressed_size, checksum):
    # struct DataBlock, 40 bytes, align 8 (schema/core.fbs:7-13)
    return struct.pack("<QQB7xQI4x", offset, size, compression, compressed_size, checksum)


def build_mvf(spaces, metadata_columns=(), string_heap=b"", format_version=1,
              reference_total_vectors_bug=False):
    """spaces: list of dict(name, dimension, vector_type, distance_metric, data_type, data=bytes).
    metadata_columns: list of dict(name, data_type, data=bytes, null_count).
    reference_total_vectors_bug=True reproduces src/builder.rs:476 (bytes / (dim*4) for every dtype).
    Returns the file bytes."""
    blocks, payload, off = [], b"", 0
    for sp in spaces:
        d = sp["data"]
        blocks.append(_data_block(off, len(d), 0, 0, zlib.crc32(d) & 0xFFFFFFFF))
        payload += d
        off += len(d)
    for col in metadata_columns:
        d = col["data"]
        blocks.append(_data_block(off, len(d), 0, 0, zlib.crc32(d) & 0xFFFFFFFF))
        payload += d
        off += len(d)
    tomb_block, ids_block = {}, {}
    for i, sp in enumerate(spaces):      # deletion bitmaps after the metadata columns (position is ours: the reference writes none)
        tb = sp.get("tombstones")
        if not tb and sp.get("tombstone_list") is not None and len(sp["tombstone_list"]):
            tb = struct.pack("<%dQ" % len(sp["tombstone_list"]), *sp["tombstone_list"])   # TombstoneFormat::SortedList, u64 LE
        if tb:
            blocks.append(_data_block(off, len(tb), 0, 0, zlib.crc32(tb) & 0xFFFFFFFF))
            payload += tb
            off += len(tb)
            tomb_block[i] = len(blocks) - 1
    for i, sp in enumerate(spaces):      # external ids, u64 LE per row (src/builder.rs:257-271), after the tombstones
        ids = sp.get("vector_ids")
        if ids is not None and len(ids):
            ib = struct.pack("<%dQ" % len(ids), *ids)
            blocks.append(_data_block(off, len(ib), 0, 0, zlib.crc32(ib) & 0xFFFFFFFF))
            payload += ib
            off += len(ib)
            ids_block[i] = len(blocks) - 1
    heap_index = 0
    if string_heap:
        blocks.append(_data_block(off, len(string_heap), 0, 0, zlib.crc32(string_heap) & 0xFFFFFFFF))
        payload += string_heap
        heap_index = len(blocks) - 1

    b = _FbBuilder()
    space_offs = []
    for i, sp in enumerate(spaces):
        name = b.create_string(sp["name"])
        b.start_table(0)
        flat = b.end_table()                       # FlatIndex {} (src/builder.rs:465-468)
        dim = sp["dimension"]
        es = 4 if reference_total_vectors_bug else ELEM_SIZE[sp["data_type"]]
        total = sp.get("total_vectors")
        if total is None:
            total = (len(sp["data"]) // (dim * es)) if dim else 0
        tomb = None
        if i in tomb_block:
            is_list = not sp.get("tombstones")
            deleted = len(sp["tombstone_list"]) if is_list else sum(bin(x).count("1") for x in sp["tombstones"])
            b.start_table(3)
            b.slot_scalar(2, "Q", deleted, 0)
            b.slot_scalar(1, "I", tomb_block[i], 0)
            b.slot_scalar(0, "B", 2 if is_list else 1, 0)   # TombstoneFormat::Bitmap / SortedList
            tomb = b.end_table()
        b.start_table(12)
        b.slot_scalar(2, "Q", total, 0)
        b.slot_offset(11, tomb)
        b.slot_scalar(9, "I", ids_block.get(i, 0), 0)   # vector_ids_block_index, 0 = positions
        b.slot_offset(8, flat)
        b.slot_scalar(6, "I", i, 0)
        b.slot_scalar(1, "I", dim, 0)
        b.slot_offset(0, name)
        b.slot_scalar(7, "B", 1, 0)                # Index::FlatIndex
        b.slot_scalar(5, "B", sp["data_type"], 0)
        b.slot_scalar(4, "B", sp["distance_metric"], 0)
        b.slot_scalar(3, "B", sp.get("vector_type", 0), 0)
        space_offs.append(b.end_table())
    vs_vec = b.create_offset_vector(space_offs)
    bm_vec = b.create_struct_vector(blocks, 40, 8)
    md_vec = None
    if metadata_columns:
        col_offs = []
        for i, col in enumerate(metadata_columns):
            name = b.create_string(col["name"])
            b.start_table(6)
            b.slot_scalar(3, "Q", col.get("null_count", 0), 0)
            b.slot_scalar(2, "I", len(spaces) + i, 0)
            b.slot_offset(0, name)
            b.slot_scalar(1, "B", col.get("data_type", 0), 0)
            col_offs.append(b.end_table())
        md_vec = b.create_offset_vector(col_offs)
    b.start_table(8)
    b.slot_scalar(4, "I", heap_index, 0)
    b.slot_offset(3, md_vec)
    b.slot_offset(2, bm_vec)
    b.slot_offset(1, vs_vec)
    b.slot_scalar(6, "H", 1, 3)                    # compatibility_version (default 3)
    b.slot_scalar(0, "H", format_version, 3)       # format_version (default 3)
    root = b.end_table()
    footer = b.finish_minimal(root)
    return MAGIC + payload + footer + struct.pack("<I", len(footer)) + MAGIC


# ---------------------------------------------------------------------------------------------
# Independent parser (restates src/reader.rs:225-278 + the FlatBuffers table walk)
# ---------------------------------------------------------------------------------------------
class MvfFormatError(Exception):
    pass


class _Table:
    def __init__(self, buf, pos):
        self.buf, self.pos = buf, pos
        soff = struct.unpack_from("<i", buf, pos)[0]
        self.vt = pos - soff
        self.vt_size = struct.unpack_from("<H", buf, self.vt)[0]

    def _field(self, slot):
        o = 4 + 2 * slot
        if o + 2 > self.vt_size:
            return 0
        return struct.unpack_from("<H", self.buf, self.vt + o)[0]

    def scalar(self, slot, fmt, default):
        f = self._field(slot)
        return struct.unpack_from("<" + fmt, self.buf, self.pos + f)[0] if f else default

    def indirect(self, slot):
        f = self._field(slot)
        if not f:
            return None
        p = self.pos + f
        return p + struct.unpack_from("<I", self.buf, p)[0]

    def string(self, slot):
        p = self.indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.buf, p)[0]
        return bytes(self.buf[p + 4:p + 4 + n]).decode("utf-8")


def parse_mvf(data):
    """Returns dict(version, spaces=[...], blocks=[...], metadata_columns=[names]).  Raises MvfFormatError."""
    if len(data) < 12:
        raise MvfFormatError("File too small")
    if data[:4] != MAGIC:
        raise MvfFormatError("Invalid magic bytes at start of file")
    if data[-4:] != MAGIC:
        raise MvfFormatError("Invalid magic bytes at end of file, it may be corrupted")
    flen = struct.unpack_from("<I", data, len(data) - 8)[0]
    if flen + 8 > len(data) - 4:
        raise MvfFormatError("Invalid footer length")
    footer = data[len(data) - 8 - flen:len(data) - 8]
    try:
        root = _Table(footer, struct.unpack_from("<I", footer, 0)[0])
        version = root.scalar(0, "H", 3)
        out = {"version": version, "spaces": [], "blocks": [], "metadata_columns": None,
               "string_heap_block_index": root.scalar(4, "I", 0)}
        bm = root.indirect(2)
        nb = struct.unpack_from("<I", footer, bm)[0]
        for i in range(nb):
            off, size, comp, csize, crc = struct.unpack_from("<QQB7xQI", footer, bm + 4 + 40 * i)
            out["blocks"].append(dict(offset=off, size=size, compression=comp,
                                      compressed_size=csize, checksum=crc))
        vs = root.indirect(1)
        ns = struct.unpack_from("<I", footer, vs)[0]
        for i in range(ns):
            p = vs + 4 + 4 * i
            t = _Table(footer, p + struct.unpack_from("<I", footer, p)[0])
            tomb = t.indirect(11)
            tinfo = None
            if tomb is not None:
                tt = _Table(footer, tomb)
                tinfo = dict(format=tt.scalar(0, "B", 0), data_block_index=tt.scalar(1, "I", 0), deleted_count=tt.scalar(2, "Q", 0))
            out["spaces"].append(dict(
                tombstones=tinfo,
                name=t.string(0), dimension=t.scalar(1, "I", 0), total_vectors=t.scalar(2, "Q", 0),
                vector_type=t.scalar(3, "B", 0), distance_metric=t.scalar(4, "B", 0),
                data_type=t.scalar(5, "B", 0), vectors_block_index=t.scalar(6, "I", 0),
                index_type_type=t.scalar(7, "B", 0), vector_ids_block_index=t.scalar(9, "I", 0)))
        md = root.indirect(3)
        if md is not None:
            nm = struct.unpack_from("<I", footer, md)[0]
            cols = []
            for i in range(nm):
                p = md + 4 + 4 * i
                t = _Table(footer, p + struct.unpack_from("<I", footer, p)[0])
                cols.append(t.string(0))
            out["metadata_columns"] = cols
    except (struct.error, IndexError, UnicodeDecodeError) as e:
        raise MvfFormatError("Failed to parse footer: %s" % e)
    return out


def get_vector_bytes(data, parsed, space_index, i):
    """src/vectors/vector_space.rs:101-142 row addressing."""
    sp = parsed["spaces"][space_index]
    blk = parsed["blocks"][sp["vectors_block_index"]]
    rb = sp["dimension"] * ELEM_SIZE[sp["data_type"]]
    start = 4 + blk["offset"] + i * rb
    return data[start:start + rb]
