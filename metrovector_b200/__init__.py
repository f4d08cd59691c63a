"""metrovector_b200 -- Python mirror of the reference's reader API plus the B200 search engine.

Everything here is a thin ctypes binding over the C ABI of ``libmvf_b200.so``
(``include/mvf_b200.h``); there is no Python or CPU implementation of the scan.  Names follow
thegenem0/metrovector: ``MvfReader`` (``MvfFile`` in its README), ``VectorSpace``, ``Vector``,
``DistanceMetric``, ``DataType``, ``MvfBuilder`` and the example's ``find_top_k_similar``
(examples/similarity_search.rs:140-176).  Importing fails loudly when the native library is
missing -- build it with ``make -C metrovector_b200/csrc`` or ``__graft_entry__.build()``.
"""
import ctypes
import enum
import os
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVF_B200_LIB selects another build of the same library (e.g. the -DMVF_TC_STATS diagnostic build, scripts/gpu_stats.sh)
LIB_PATH = os.environ.get("MVF_B200_LIB") or os.path.join(_HERE, "libmvf_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "metrovector_b200: native library %s is missing; build it with `make -C metrovector_b200/csrc` "
        "(there is no Python/CPU fallback)" % LIB_PATH)

_lib = ctypes.CDLL(LIB_PATH)
_c = ctypes


class DataType(enum.IntEnum):           # schema/types.fbs:3-11
    Float32 = 0
    Float16 = 1
    Int8 = 2
    UInt8 = 3
    UInt32 = 4
    UInt64 = 5
    StringRef = 6


class VectorType(enum.IntEnum):         # schema/types.fbs:14-17
    Dense = 0
    Sparse = 1


class DistanceMetric(enum.IntEnum):     # schema/types.fbs:19-25 (README spells them Euclidean / Dot)
    L2 = 0
    InnerProduct = 1
    Cosine = 2
    Custom = 255


DistanceMetric.Euclidean = DistanceMetric.L2
DistanceMetric.Dot = DistanceMetric.InnerProduct

_NP_DTYPE = {DataType.Float32: np.float32, DataType.Float16: np.float16, DataType.Int8: np.int8,
             DataType.UInt8: np.uint8}


class MvfError(Exception):
    """src/errors.rs:8-40.  ``code`` is the C status, ``kind`` the variant name."""
    code = 0
    kind = "MvfError"


def _mk(name, code):
    return type(name, (MvfError,), {"code": code, "kind": name})


Io = _mk("Io", 1)
InvalidFormat = _mk("InvalidFormat", 2)
UnsupportedVersion = _mk("UnsupportedVersion", 3)
VectorSpaceNotFound = _mk("VectorSpaceNotFound", 4)
IndexOutOfBounds = _mk("IndexOutOfBounds", 5)
DimensionMismatch = _mk("DimensionMismatch", 6)
InvalidVectorType = _mk("InvalidVectorType", 7)
CorruptedData = _mk("CorruptedData", 8)
Extension = _mk("Extension", 9)
Build = _mk("Build", 10)
Cuda = _mk("Cuda", 11)
Nccl = _mk("Nccl", 12)
_ERRORS = {e.code: e for e in (Io, InvalidFormat, UnsupportedVersion, VectorSpaceNotFound, IndexOutOfBounds,
                               DimensionMismatch, InvalidVectorType, CorruptedData, Extension, Build, Cuda, Nccl)}

_lib.mvf_last_error.restype = _c.c_char_p
_lib.mvf_version_string.restype = _c.c_char_p
_lib.mvf_ordered_key_f32.restype = _c.c_uint32
_lib.mvf_ordered_key_f32.argtypes = [_c.c_float]
_lib.mvf_ordered_key_to_f32.restype = _c.c_float
_lib.mvf_ordered_key_to_f32.argtypes = [_c.c_uint32]
for _fn in ("mvf_file_close", "mvf_space_release", "mvf_builder_free", "mvf_gpu_shutdown", "mvf_gpu_space_free"):
    getattr(_lib, _fn).restype = None
    getattr(_lib, _fn).argtypes = [_c.c_void_p]

_P, _U64, _U32, _INT, _SZ = _c.c_void_p, _c.c_uint64, _c.c_uint32, _c.c_int, _c.c_size_t
_SIGS = {
    "mvf_file_open": [_c.c_char_p, _c.POINTER(_P)],
    "mvf_file_version": [_P, _c.POINTER(_c.c_uint16)],
    "mvf_file_num_vector_spaces": [_P, _c.POINTER(_SZ)],
    "mvf_file_vector_space_name": [_P, _SZ, _c.POINTER(_c.c_char_p)],
    "mvf_file_size": [_P, _c.POINTER(_U64)],
    "mvf_file_has_metadata": [_P, _c.POINTER(_INT)],
    "mvf_file_num_metadata_columns": [_P, _c.POINTER(_SZ)],
    "mvf_file_metadata_column_name": [_P, _SZ, _c.POINTER(_c.c_char_p)],
    "mvf_file_validate": [_P],
    "mvf_file_validate_with_checksum": [_P],
    "mvf_file_vector_space": [_P, _c.c_char_p, _c.POINTER(_P)],
    "mvf_space_name": [_P, _c.POINTER(_c.c_char_p)],
    "mvf_space_dimension": [_P, _c.POINTER(_U32)],
    "mvf_space_total_vectors": [_P, _c.POINTER(_U64)],
    "mvf_space_vector_type": [_P, _c.POINTER(_INT)],
    "mvf_space_distance_metric": [_P, _c.POINTER(_INT)],
    "mvf_space_data_type": [_P, _c.POINTER(_INT)],
    "mvf_space_get_vector": [_P, _U64, _c.POINTER(_P), _c.POINTER(_SZ)],
    "mvf_space_get_vector_f32": [_P, _U64, _P, _SZ],
    "mvf_space_map_vector_range": [_P, _U64, _U64, _c.POINTER(_P), _c.POINTER(_SZ)],
    "mvf_space_scannable_vectors": [_P, _c.POINTER(_U64)],
    "mvf_builder_new": [_c.POINTER(_P)],
    "mvf_builder_add_vector_space": [_P, _c.c_char_p, _U32, _INT, _INT, _INT, _c.POINTER(_SZ)],
    "mvf_builder_add_vectors_f32": [_P, _c.c_char_p, _P, _SZ, _U32],
    "mvf_builder_add_vectors_raw": [_P, _c.c_char_p, _P, _SZ, _U32],
    "mvf_builder_add_metadata_column": [_P, _c.c_char_p, _INT, _P, _SZ],
    "mvf_builder_save": [_P, _c.c_char_p],
    "mvf_gpu_init": [_INT, _c.POINTER(_P)],
    "mvf_gpu_init_multi": [_c.POINTER(_INT), _INT, _c.POINTER(_P)],
    "mvf_gpu_device_count": [_P, _c.POINTER(_INT)],
    "mvf_gpu_space_path_facts": [_P, _c.POINTER(_U64), _c.POINTER(_INT), _c.POINTER(_U64)],
    "mvf_path_decision": [_U64, _INT, _INT, _INT, _U32, _U32, _U32, _c.POINTER(_U64), _c.POINTER(_INT)],
    "mvf_gpu_set_stream": [_P, _P],
    "mvf_gpu_device_info": [_P, _c.POINTER(_INT), _c.POINTER(_INT), _c.POINTER(_INT), _c.POINTER(_U64), _c.POINTER(_U64)],
    "mvf_gpu_comm_unique_id": [_P],
    "mvf_gpu_comm_init": [_P, _P, _INT, _INT],
    "mvf_shard_range": [_U64, _INT, _INT, _c.POINTER(_U64), _c.POINTER(_U64)],
    "mvf_gpu_space_upload": [_P, _P, _c.POINTER(_P)],
    "mvf_gpu_space_from_host": [_P, _P, _INT, _INT, _U64, _U32, _U64, _U64, _c.POINTER(_P)],
    "mvf_gpu_space_synthetic": [_P, _INT, _INT, _U64, _U32, _U64, _c.POINTER(_P)],
    "mvf_gpu_space_info": [_P, _c.POINTER(_U64), _c.POINTER(_U64), _c.POINTER(_U64), _c.POINTER(_U32),
                           _c.POINTER(_INT), _c.POINTER(_INT)],
    "mvf_gpu_space_upload_stats": [_P, _c.POINTER(_INT), _c.POINTER(_c.c_double), _c.POINTER(_U64)],
    "mvf_gpu_space_read_rows": [_P, _U64, _U64, _P],
    "mvf_gpu_space_read_norms": [_P, _P, _P],
    "mvf_gpu_search": [_P, _P, _INT, _U32, _U32, _INT, _P, _P, _P],
    "mvf_gpu_search_device": [_P, _P, _INT, _U32, _U32, _INT, _P, _P, _P],
    "mvf_gpu_space_set_row_mask": [_P, _P, _U64],
    "mvf_space_tombstones": [_P, _c.POINTER(_P), _c.POINTER(_SZ), _c.POINTER(_U64)],
    "mvf_builder_set_tombstones": [_P, _c.c_char_p, _P, _SZ],
    "mvf_builder_set_tombstone_list": [_P, _c.c_char_p, _P, _SZ],
    "mvf_builder_set_vector_ids": [_P, _c.c_char_p, _P, _SZ],
    "mvf_space_tombstone_list": [_P, _c.POINTER(_P), _c.POINTER(_SZ), _c.POINTER(_U64)],
    "mvf_space_vector_ids": [_P, _c.POINTER(_P), _c.POINTER(_SZ)],
    "mvf_file_metadata_column": [_P, _c.c_char_p, _c.POINTER(_P), _c.POINTER(_SZ), _c.POINTER(_INT)],
    "mvf_filter_range_bitmap": [_P, _INT, _U64, _c.c_double, _c.c_double, _P],
    "mvf_gpu_space_set_tombstones": [_P, _P, _U64],
    "mvf_gpu_space_set_ids": [_P, _P, _U64],
    "mvf_gpu_space_checksum": [_P, _c.POINTER(_U32), _c.POINTER(_c.c_double)],
    "mvf_gpu_space_verify_checksum": [_P, _P],
    "mvf_gpu_space_get_vectors": [_P, _P, _U32, _INT, _P],
    "mvf_gpu_tc_dump": [_P, _P, _INT, _U32, _U32, _INT, _P, _U64, _c.POINTER(_U32), _c.POINTER(_U32)],
    "mvf_gpu_set_path": [_P, _INT],
    "mvf_tc_plan": [_INT, _U32, _U32, _U64, _U32, _INT, _P],
    "mvf_gpu_profile": [_P, _INT],
    "mvf_gpu_profile_read": [_P, _c.POINTER(_c.c_double), _c.POINTER(_U64), _c.POINTER(_U64), _c.POINTER(_INT)],
    "mvf_gpu_synchronize": [_P],
    "mvf_gpu_exchange_info": [_P, _c.POINTER(_INT), _c.POINTER(_U64), _c.POINTER(_U64)],
    "mvf_synth_fill_host": [_P, _INT, _U64, _U64, _U64, _U32],
}
for _name, _args in _SIGS.items():
    getattr(_lib, _name).restype = _INT
    getattr(_lib, _name).argtypes = _args


TC_PLAN_FIELDS = ("mtiles", "ntiles", "chunks", "tiles_per_chunk", "cluster", "mgroups", "items", "grid", "kblocks", "cap",
                  "trigger", "sync_window", "epi_groups", "stages", "pair", "smem")


def tc_plan(data_type, nq, k, n, dimension, sm_count=148):
    """Launch shape of the tensor-core kernel for this problem (host arithmetic only, no GPU needed)."""
    out = (_c.c_uint32 * 16)()
    _check(_lib.mvf_tc_plan(int(data_type), nq, k, n, dimension, sm_count, out))
    return dict(zip(TC_PLAN_FIELDS, [int(v) for v in out]))


def path_decision(global_n, world, data_type, metric, nq, k, dimension):
    """(rows the scan-vs-tensor decision is taken on, would-take-the-tensor-path) for a space sharded over `world` ranks;
    host arithmetic only."""
    rows, tp = _U64(), _INT()
    _check(_lib.mvf_path_decision(global_n, world, int(data_type), int(metric), nq, k, dimension, _c.byref(rows), _c.byref(tp)))
    return rows.value, bool(tp.value)


def filter_range_bitmap(column, data_type, lo, hi):
    """Boolean array, True = row FAILS lo <= value <= hi (the exclusion convention of GpuSpace.set_row_mask).
    column: raw bytes or a numpy array of the column's values."""
    raw = column.tobytes() if isinstance(column, np.ndarray) else bytes(column)
    es = {0: 4, 1: 2, 2: 1, 3: 1, 4: 4, 5: 8}[int(data_type)]
    n = len(raw) // es
    out = np.zeros((n + 7) // 8, dtype=np.uint8)
    _check(_lib.mvf_filter_range_bitmap(raw, int(data_type), n, float(lo), float(hi), out.ctypes.data_as(_P)))
    return np.unpackbits(out, bitorder="little")[:n].astype(bool)


_lib.mvf_last_error_fields.restype = None
_lib.mvf_last_error_fields.argtypes = [_c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64)]


def _check(rc):
    if rc != 0:
        msg = (_lib.mvf_last_error() or b"").decode("utf-8", "replace")
        a, b = _c.c_uint64(), _c.c_uint64()
        _lib.mvf_last_error_fields(_c.byref(a), _c.byref(b))
        err = _ERRORS.get(rc, MvfError)(msg)
        err.fields = (a.value, b.value)        # (got, expected) / (index, len) / (expected, actual) of the structured variants
        raise err


def version_string():
    return _lib.mvf_version_string().decode()


def ordered_key_f32(v):
    return _lib.mvf_ordered_key_f32(float(v))


def ordered_key_to_f32(k):
    return _lib.mvf_ordered_key_to_f32(int(k))


def shard_range(n, world, rank):
    b, c = _U64(), _U64()
    _check(_lib.mvf_shard_range(n, world, rank, _c.byref(b), _c.byref(c)))
    return b.value, c.value


def synth_fill_host(dtype, seed, row0, nrows, d):
    """Product-side synthetic generator (host flavour), same definition as the device fill."""
    out = np.empty((nrows, d), dtype=_NP_DTYPE[DataType(dtype)])
    _check(_lib.mvf_synth_fill_host(out.ctypes.data_as(_P), int(dtype), seed, row0, nrows, d))
    return out


# ------------------------------------------------------------------------------------------------
class Vector:
    """src/vectors/vector.rs:28-33."""

    def __init__(self, owner, ptr, nbytes, dimension, data_type):
        self._owner, self._ptr, self._nbytes = owner, ptr, nbytes
        self._dimension, self._data_type = dimension, DataType(data_type)

    def dimension(self):
        return self._dimension

    def data_type(self):
        return self._data_type

    def as_bytes(self):                                   # vector.rs:61-63
        return _c.string_at(self._ptr, self._nbytes)

    def as_f32(self):                                     # vector.rs:71-92
        if self._data_type == DataType.Float32:
            return np.frombuffer(self.as_bytes(), dtype="<f4").copy()
        if self._data_type == DataType.Float16:
            return np.frombuffer(self.as_bytes(), dtype="<f2").astype(np.float32)
        raise Build("Build error: Cannot convert to f32")


class VectorSpace:
    """src/vectors/vector_space.rs:34-39.  Borrows from its reader."""

    def __init__(self, reader, handle):
        self._reader, self._h = reader, handle

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.mvf_space_release(self._h)
            self._h = None

    def _get(self, fn, ctype):
        v = ctype()
        _check(getattr(_lib, fn)(self._h, _c.byref(v)))
        return v.value

    def name(self):
        return self._get("mvf_space_name", _c.c_char_p).decode()

    def dimension(self):
        return self._get("mvf_space_dimension", _U32)

    def total_vectors(self):
        return self._get("mvf_space_total_vectors", _U64)

    def vector_type(self):
        return VectorType(self._get("mvf_space_vector_type", _INT))

    def distance_metric(self):
        return DistanceMetric(self._get("mvf_space_distance_metric", _INT))

    def data_type(self):
        return DataType(self._get("mvf_space_data_type", _INT))

    def scannable_vectors(self):
        return self._get("mvf_space_scannable_vectors", _U64)

    def tombstones(self):
        """Deletion bitmap of the footer's TombstoneInfo{Bitmap} as a boolean array, or None."""
        ptr, n, cnt = _P(), _SZ(), _U64()
        _check(_lib.mvf_space_tombstones(self._h, _c.byref(ptr), _c.byref(n), _c.byref(cnt)))
        if not ptr.value:
            return None
        raw = np.frombuffer(_c.string_at(ptr.value, n.value), dtype=np.uint8)
        return np.unpackbits(raw, bitorder="little").astype(bool)

    def tombstone_list(self):
        """Deleted row positions of the footer's TombstoneInfo{SortedList} (uint64 array), or None."""
        ptr, n, cnt = _P(), _SZ(), _U64()
        _check(_lib.mvf_space_tombstone_list(self._h, _c.byref(ptr), _c.byref(n), _c.byref(cnt)))
        if not ptr.value:
            return None
        return np.frombuffer(_c.string_at(ptr.value, n.value * 8), dtype="<u8").copy()

    def vector_ids(self):
        """External ids (vector_ids_block_index, schema/core.fbs:54) as a uint64 array, or None = positions."""
        ptr, n = _P(), _SZ()
        _check(_lib.mvf_space_vector_ids(self._h, _c.byref(ptr), _c.byref(n)))
        if not ptr.value:
            return None
        return np.frombuffer(_c.string_at(ptr.value, n.value * 8), dtype="<u8").copy()

    def get_vector(self, index):                          # vector_space.rs:101-142
        ptr, n = _P(), _SZ()
        _check(_lib.mvf_space_get_vector(self._h, index, _c.byref(ptr), _c.byref(n)))
        return Vector(self, ptr.value, n.value, self.dimension(), self.data_type())

    def map_vector_range(self, start, count):             # vector_space.rs:155-188
        ptr, n = _P(), _SZ()
        _check(_lib.mvf_space_map_vector_range(self._h, start, count, _c.byref(ptr), _c.byref(n)))
        raw = _c.string_at(ptr.value, n.value) if n.value else b""
        dt = _NP_DTYPE[self.data_type()]
        return np.frombuffer(raw, dtype=dt).reshape(count, self.dimension()) if count else np.zeros((0, self.dimension()), dt)

    def clone_concurrent(self):                           # vector_space.rs:194-201
        return self._reader.vector_space(self.name())


class MvfReader:
    """src/reader.rs:27-31."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def open(cls, path):                                  # reader.rs:45-79
        h = _P()
        _check(_lib.mvf_file_open(os.fsencode(path), _c.byref(h)))
        return cls(h)

    def close(self):
        if self._h:
            _lib.mvf_file_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def version(self):
        v = _c.c_uint16()
        _check(_lib.mvf_file_version(self._h, _c.byref(v)))
        return v.value

    def num_vector_spaces(self):
        v = _SZ()
        _check(_lib.mvf_file_num_vector_spaces(self._h, _c.byref(v)))
        return v.value

    def vector_space_names(self):
        out = []
        for i in range(self.num_vector_spaces()):
            s = _c.c_char_p()
            _check(_lib.mvf_file_vector_space_name(self._h, i, _c.byref(s)))
            out.append(s.value.decode())
        return out

    def vector_space(self, name):                         # reader.rs:104-119
        h = _P()
        _check(_lib.mvf_file_vector_space(self._h, name.encode(), _c.byref(h)))
        return VectorSpace(self, h)

    def file_size(self):
        v = _U64()
        _check(_lib.mvf_file_size(self._h, _c.byref(v)))
        return v.value

    def has_metadata(self):
        v = _INT()
        _check(_lib.mvf_file_has_metadata(self._h, _c.byref(v)))
        return bool(v.value)

    def metadata_column_names(self):
        n = _SZ()
        _check(_lib.mvf_file_num_metadata_columns(self._h, _c.byref(n)))
        out = []
        for i in range(n.value):
            s = _c.c_char_p()
            _check(_lib.mvf_file_metadata_column_name(self._h, i, _c.byref(s)))
            out.append(s.value.decode())
        return out

    def metadata_column(self, name):
        """(raw bytes, DataType) of a metadata column (schema/core.fbs:16-25)."""
        ptr, n, dt = _P(), _SZ(), _INT()
        _check(_lib.mvf_file_metadata_column(self._h, name.encode(), _c.byref(ptr), _c.byref(n), _c.byref(dt)))
        return _c.string_at(ptr.value, n.value), DataType(dt.value)

    def validate(self):
        _check(_lib.mvf_file_validate(self._h))

    def validate_with_checksum(self):
        _check(_lib.mvf_file_validate_with_checksum(self._h))


MvfFile = MvfReader     # README.md:97


class MvfBuilder:
    """src/builder.rs:93-98."""

    def __init__(self):
        self._h = _P()
        _check(_lib.mvf_builder_new(_c.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.mvf_builder_free(self._h)
            self._h = None

    def add_vector_space(self, name, dimension, vector_type, distance_metric, data_type):
        idx = _SZ()
        _check(_lib.mvf_builder_add_vector_space(self._h, name.encode(), dimension, int(vector_type),
                                                 int(distance_metric), int(data_type), _c.byref(idx)))
        return idx.value

    def add_vectors(self, space, vectors):
        """Rows of f32 (src/builder.rs:151-196)."""
        a = np.ascontiguousarray(vectors, dtype=np.float32)
        if a.size == 0:
            return
        _check(_lib.mvf_builder_add_vectors_f32(self._h, space.encode(), a.ctypes.data_as(_P), a.shape[0], a.shape[1]))

    def add_vectors_raw(self, space, rows):
        a = np.ascontiguousarray(rows)
        if a.size == 0:
            return
        _check(_lib.mvf_builder_add_vectors_raw(self._h, space.encode(), a.ctypes.data_as(_P), a.shape[0], a.shape[1]))

    def set_tombstones(self, space, deleted):
        """deleted: boolean array over the space's rows; written as TombstoneInfo{Bitmap}."""
        bits = np.packbits(np.asarray(deleted, dtype=bool), bitorder="little").tobytes()
        _check(_lib.mvf_builder_set_tombstones(self._h, space.encode(), bits, len(bits)))

    def set_tombstone_list(self, space, rows):
        """rows: strictly ascending deleted positions; written as TombstoneInfo{SortedList} (u64 LE)."""
        a = np.ascontiguousarray(rows, dtype=np.uint64)
        _check(_lib.mvf_builder_set_tombstone_list(self._h, space.encode(), a.ctypes.data_as(_P), a.size))

    def set_vector_ids(self, space, ids):
        """ids: one external uint64 id per row of the space (VectorSpaceBuilder.vector_ids, src/builder.rs:61)."""
        a = np.ascontiguousarray(ids, dtype=np.uint64)
        _check(_lib.mvf_builder_set_vector_ids(self._h, space.encode(), a.ctypes.data_as(_P), a.size))

    def add_metadata_column(self, name, data_type, data):
        b = bytes(data)
        _check(_lib.mvf_builder_add_metadata_column(self._h, name.encode(), int(data_type), b, len(b)))

    def save(self, path):
        _check(_lib.mvf_builder_save(self._h, os.fsencode(path)))


# ------------------------------------------------------------------------------------------------
ScoredVector = namedtuple("ScoredVector", ["index", "score"])    # similarity_search.rs:14-19


class GpuEngine:
    """The search engine on one B200 (``GpuEngine(0)``) or, from one process, on several (``GpuEngine([0, 1, ...])``:
    spaces are row-sharded over the devices and every search uses all of them)."""

    def __init__(self, device=0):
        self._h = _P()
        if isinstance(device, (list, tuple)):
            devs = (_INT * len(device))(*[int(x) for x in device])
            _check(_lib.mvf_gpu_init_multi(devs, len(device), _c.byref(self._h)))
            self.device = int(device[0]) if device else 0
        else:
            _check(_lib.mvf_gpu_init(device, _c.byref(self._h)))
            self.device = device
        self.rank, self.world = 0, 1

    def device_count(self):
        n = _INT()
        _check(_lib.mvf_gpu_device_count(self._h, _c.byref(n)))
        return n.value

    def shutdown(self):
        if getattr(self, "_h", None):
            _lib.mvf_gpu_shutdown(self._h)
            self._h = None

    def __del__(self):
        self.shutdown()

    def set_stream(self, cuda_stream_ptr):
        _check(_lib.mvf_gpu_set_stream(self._h, _P(cuda_stream_ptr)))

    def device_info(self):
        sm, ma, mi, fr, to = _INT(), _INT(), _INT(), _U64(), _U64()
        _check(_lib.mvf_gpu_device_info(self._h, _c.byref(sm), _c.byref(ma), _c.byref(mi), _c.byref(fr), _c.byref(to)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), free_bytes=fr.value, total_bytes=to.value)

    @staticmethod
    def comm_unique_id():
        buf = _c.create_string_buffer(128)
        _check(_lib.mvf_gpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        _check(_lib.mvf_gpu_comm_init(self._h, unique_id, rank, world))
        self.rank, self.world = rank, world

    def set_path(self, path):
        _check(_lib.mvf_gpu_set_path(self._h, path))

    def profile(self, enable):
        _check(_lib.mvf_gpu_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        ms, sl, al, lp = _c.c_double(), _U64(), _U64(), _INT()
        _check(_lib.mvf_gpu_profile_read(self._h, _c.byref(ms), _c.byref(sl), _c.byref(al), _c.byref(lp)))
        return dict(scan_ms_total=ms.value, scan_launches=sl.value, all_launches=al.value, last_path=lp.value)

    def synchronize(self):
        _check(_lib.mvf_gpu_synchronize(self._h))

    def exchange_info(self):
        """How sharded searches exchange their candidates: peer-memory kernel set up?, searches per route, note."""
        ready, ps, ns = _INT(), _U64(), _U64()
        _check(_lib.mvf_gpu_exchange_info(self._h, _c.byref(ready), _c.byref(ps), _c.byref(ns)))
        _lib.mvf_gpu_exchange_note.restype = _c.c_char_p
        _lib.mvf_gpu_exchange_note.argtypes = [_P]
        note = _lib.mvf_gpu_exchange_note(self._h)
        return dict(peer_ready=bool(ready.value), peer_searches=ps.value, nccl_searches=ns.value,
                    note=(note or b"").decode())

    def upload(self, space):
        h = _P()
        _check(_lib.mvf_gpu_space_upload(self._h, space._h, _c.byref(h)))
        return GpuSpace(self, h)

    def from_host(self, rows, data_type, metric, row_base=0, global_n=None):
        a = np.ascontiguousarray(rows, dtype=_NP_DTYPE[DataType(data_type)])
        n, d = a.shape
        h = _P()
        _check(_lib.mvf_gpu_space_from_host(self._h, a.ctypes.data_as(_P), int(data_type), int(metric), n, d, row_base,
                                            n if global_n is None else global_n, _c.byref(h)))
        return GpuSpace(self, h)

    def synthetic(self, data_type, metric, global_n, dimension, seed):
        h = _P()
        _check(_lib.mvf_gpu_space_synthetic(self._h, int(data_type), int(metric), global_n, dimension, seed, _c.byref(h)))
        return GpuSpace(self, h)


class GpuSpace:
    """HBM-resident shard of one vector space."""

    def __init__(self, engine, handle):
        self.engine, self._h = engine, handle
        ln, rb, gn, d, dt, m = _U64(), _U64(), _U64(), _U32(), _INT(), _INT()
        _check(_lib.mvf_gpu_space_info(self._h, _c.byref(ln), _c.byref(rb), _c.byref(gn), _c.byref(d), _c.byref(dt), _c.byref(m)))
        self.local_n, self.row_base, self.global_n = ln.value, rb.value, gn.value
        self.dimension, self.data_type, self.distance_metric = d.value, DataType(dt.value), DistanceMetric(m.value)

    def free(self):
        if getattr(self, "_h", None):
            _lib.mvf_gpu_space_free(self._h)
            self._h = None

    def __del__(self):
        self.free()

    def path_facts(self):
        """What the scan-vs-tensor choice depends on besides the shape (same on every rank of a sharded space)."""
        nf, sp, dr = _U64(), _INT(), _U64()
        _check(_lib.mvf_gpu_space_path_facts(self._h, _c.byref(nf), _c.byref(sp), _c.byref(dr)))
        return dict(nonfinite_rows=nf.value, tf32_planes=bool(sp.value), decision_rows=dr.value)

    def upload_stats(self):
        p, s, b = _INT(), _c.c_double(), _U64()
        _check(_lib.mvf_gpu_space_upload_stats(self._h, _c.byref(p), _c.byref(s), _c.byref(b)))
        _lib.mvf_gpu_space_upload_note.restype = _c.c_char_p
        _lib.mvf_gpu_space_upload_note.argtypes = [_P]
        note = (_lib.mvf_gpu_space_upload_note(self._h) or b"").decode()
        return dict(path=p.value, seconds=s.value, bytes=b.value, note=note)

    def read_rows(self, start, count):
        out = np.empty((count, self.dimension), dtype=_NP_DTYPE[self.data_type])
        _check(_lib.mvf_gpu_space_read_rows(self._h, start, count, out.ctypes.data_as(_P)))
        return out

    def read_norms(self):
        norm = np.empty(self.local_n, dtype=np.float32)
        sq = np.empty(self.local_n, dtype=np.float32)
        _check(_lib.mvf_gpu_space_read_norms(self._h, norm.ctypes.data_as(_P), sq.ctypes.data_as(_P)))
        return norm, sq

    def set_row_mask(self, excluded):
        """excluded: boolean array over GLOBAL rows (True = never returned), or None to clear."""
        if excluded is None:
            _check(_lib.mvf_gpu_space_set_row_mask(self._h, None, 0))
            return
        bits = np.packbits(np.asarray(excluded, dtype=bool), bitorder="little")
        _check(_lib.mvf_gpu_space_set_row_mask(self._h, bits.ctypes.data_as(_P), len(excluded)))

    def set_tombstones(self, deleted):
        """deleted: boolean array over GLOBAL rows, or None; kept apart from set_row_mask (searches exclude the union)."""
        if deleted is None:
            _check(_lib.mvf_gpu_space_set_tombstones(self._h, None, 0))
            return
        bits = np.packbits(np.asarray(deleted, dtype=bool), bitorder="little")
        _check(_lib.mvf_gpu_space_set_tombstones(self._h, bits.ctypes.data_as(_P), len(deleted)))

    def set_ids(self, ids):
        """ids: uint64 array, one external id per GLOBAL row (reported by searches instead of positions), or None."""
        if ids is None:
            _check(_lib.mvf_gpu_space_set_ids(self._h, None, 0))
            return
        a = np.ascontiguousarray(ids, dtype=np.uint64)
        _check(_lib.mvf_gpu_space_set_ids(self._h, a.ctypes.data_as(_P), a.size))

    def checksum(self):
        """CRC-32/IEEE of the resident rows, computed on the device.  Returns (crc, kernel_ms)."""
        crc, ms = _U32(), _c.c_double()
        _check(_lib.mvf_gpu_space_checksum(self._h, _c.byref(crc), _c.byref(ms)))
        return crc.value, ms.value

    def verify_checksum(self, space):
        """Device CRC of the resident block against the file's block manifest (raises CorruptedData)."""
        _check(_lib.mvf_gpu_space_verify_checksum(self._h, space._h))

    def get_vectors(self, ids, as_f32=False):
        """Rows of the HBM copy for global row ids (the `vector` field of the reference's ScoredVector)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64).ravel()
        dt = np.float32 if as_f32 else _NP_DTYPE[self.data_type]
        out = np.empty((len(ids), self.dimension), dtype=dt)
        _check(_lib.mvf_gpu_space_get_vectors(self._h, ids.ctypes.data_as(_P), len(ids), 1 if as_f32 else 0, out.ctypes.data_as(_P)))
        return out

    def _query_array(self, queries):
        q = np.asarray(queries)
        if q.ndim == 1:
            q = q[None, :]
        if self.data_type in (DataType.Float32, DataType.Float16):
            q = np.ascontiguousarray(q, dtype=np.float16 if (q.dtype == np.float16 and self.data_type == DataType.Float16)
                                     else np.float32)
            qd = DataType.Float16 if q.dtype == np.float16 else DataType.Float32
        else:
            q = np.ascontiguousarray(q, dtype=_NP_DTYPE[self.data_type])
            qd = self.data_type
        if q.shape[1] != self.dimension:
            raise DimensionMismatch("Dimension mismatch: expected %d, got %d" % (self.dimension, q.shape[1]))
        return q, qd

    def search(self, queries, k, metric=None):
        """Batched find_top_k_similar.  Returns (ids u64 [nq,k], scores f32 [nq,k], iscores i32 [nq,k])."""
        q, qd = self._query_array(queries)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        scores = np.empty((nq, k), dtype=np.float32)
        iscores = np.empty((nq, k), dtype=np.int32)
        _check(_lib.mvf_gpu_search(self._h, q.ctypes.data_as(_P), int(qd), nq, k, -1 if metric is None else int(metric),
                                   ids.ctypes.data_as(_P), scores.ctypes.data_as(_P), iscores.ctypes.data_as(_P)))
        return ids, scores, iscores

    def tc_dump(self, queries, k=10, metric=None):
        """Test hook: raw ranking values of the tensor-core path, shape (padded nq, padded n)."""
        q, qd = self._query_array(queries)
        nq = q.shape[0]
        rows = ((nq + 127) // 128) * 128
        cols = ((self.local_n + 255) // 256) * 256
        out = np.empty((rows, cols), dtype=np.float32)
        r, c = _U32(), _U32()
        _check(_lib.mvf_gpu_tc_dump(self._h, q.ctypes.data_as(_P), int(qd), nq, k, -1 if metric is None else int(metric),
                                    out.ctypes.data_as(_P), out.size, _c.byref(r), _c.byref(c)))
        assert (r.value, c.value) == (rows, cols)
        return out

    def search_raw(self, q_ptr, qdtype, nq, k, metric, ids_ptr, scores_ptr, iscores_ptr=None):
        """Host-pointer search without any numpy work (bench e2e loop)."""
        _check(_lib.mvf_gpu_search(self._h, _P(q_ptr), int(qdtype), nq, k, int(metric), _P(ids_ptr), _P(scores_ptr),
                                   _P(iscores_ptr) if iscores_ptr else None))

    def search_device(self, d_q_ptr, qdtype, nq, k, metric, d_ids_ptr, d_scores_ptr, d_iscores_ptr=None):
        """Device-pointer search, asynchronous on the engine stream."""
        _check(_lib.mvf_gpu_search_device(self._h, _P(d_q_ptr), int(qdtype), nq, k, int(metric), _P(d_ids_ptr),
                                          _P(d_scores_ptr), _P(d_iscores_ptr) if d_iscores_ptr else None))


def find_top_k_similar(gpu_space, query, k):
    """examples/similarity_search.rs:140-176 on the GPU: the k nearest rows, best first."""
    ids, scores, _ = gpu_space.search(query, k)
    return [ScoredVector(int(i), float(s)) for i, s in zip(ids[0], scores[0]) if i != np.uint64(0xFFFFFFFFFFFFFFFF)]
