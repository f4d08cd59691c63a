// c_api.cpp -- extern "C" boundary over the C++ host layer (reader, writer, host helpers).
// The GPU entry points live in engine.cu.  Every function catches everything: no exception
// crosses the ABI; failures become an mvf_status plus a thread-local message (include/mvf_b200.h).
#include "c_api_common.hpp"

#include <cstring>

#include "synth.h"

namespace mvf {
thread_local std::string g_last_error;
thread_local uint64_t g_last_fields[2] = {0, 0};
void set_last_error(const std::string& m, uint64_t first, uint64_t second) { g_last_error = m; g_last_fields[0] = first; g_last_fields[1] = second; }
}  // namespace mvf

using namespace mvf;

struct mvf_file { MvfReader* r; };
struct mvf_builder { MvfBuilder b; };

extern "C" {

const char* mvf_last_error(void) { return g_last_error.c_str(); }
void mvf_last_error_fields(uint64_t* first, uint64_t* second) {
    if (first) *first = g_last_fields[0];
    if (second) *second = g_last_fields[1];
}
const char* mvf_version_string(void) { return "mvf_b200 0.1 (sm_100a)"; }

int mvf_file_open(const char* path, mvf_file** out) {
    return guard([&] {
        if (!path || !out) throw MvfError::io("null argument");
        *out = nullptr;
        MvfReader* r = MvfReader::open(path);
        *out = new mvf_file{r};
    });
}
void mvf_file_close(mvf_file* f) { if (f) { delete f->r; delete f; } }
int mvf_file_version(const mvf_file* f, uint16_t* out) { return guard([&] { *out = f->r->version(); }); }
int mvf_file_num_vector_spaces(const mvf_file* f, size_t* out) { return guard([&] { *out = f->r->num_vector_spaces(); }); }
int mvf_file_vector_space_name(const mvf_file* f, size_t i, const char** out) {
    return guard([&] {
        if (i >= f->r->spaces().size()) throw MvfError::index_out_of_bounds(i, f->r->spaces().size());
        *out = f->r->spaces()[i].name.c_str();
    });
}
int mvf_file_size(const mvf_file* f, uint64_t* out) { return guard([&] { *out = f->r->file_size(); }); }
int mvf_file_has_metadata(const mvf_file* f, int* out) { return guard([&] { *out = f->r->has_metadata() ? 1 : 0; }); }
int mvf_file_num_metadata_columns(const mvf_file* f, size_t* out) {
    return guard([&] { *out = f->r->metadata_column_names().size(); });
}
int mvf_file_metadata_column_name(const mvf_file* f, size_t i, const char** out) {
    return guard([&] {
        const auto& n = f->r->metadata_column_names();
        if (i >= n.size()) throw MvfError::index_out_of_bounds(i, n.size());
        *out = n[i].c_str();
    });
}
int mvf_file_validate(const mvf_file* f) { return guard([&] { f->r->validate(); }); }
int mvf_file_validate_with_checksum(const mvf_file* f) { return guard([&] { f->r->validate_with_checksum(); }); }
int mvf_file_vector_space(const mvf_file* f, const char* name, mvf_space** out) {
    return guard([&] {
        *out = nullptr;
        VectorSpace vs = f->r->vector_space(name ? name : "");
        *out = new mvf_space{vs};
    });
}
void mvf_space_release(mvf_space* s) { delete s; }

int mvf_space_name(const mvf_space* s, const char** out) { return guard([&] { *out = s->vs.name().c_str(); }); }
int mvf_space_dimension(const mvf_space* s, uint32_t* out) { return guard([&] { *out = s->vs.dimension(); }); }
int mvf_space_total_vectors(const mvf_space* s, uint64_t* out) { return guard([&] { *out = s->vs.total_vectors(); }); }
int mvf_space_vector_type(const mvf_space* s, int* out) { return guard([&] { *out = (int)s->vs.vector_type(); }); }
int mvf_space_distance_metric(const mvf_space* s, int* out) { return guard([&] { *out = (int)s->vs.distance_metric(); }); }
int mvf_space_data_type(const mvf_space* s, int* out) { return guard([&] { *out = (int)s->vs.data_type(); }); }
int mvf_space_get_vector(const mvf_space* s, uint64_t index, const void** ptr, size_t* nbytes) {
    return guard([&] {
        Vector v = s->vs.get_vector(index);
        *ptr = v.as_bytes(); *nbytes = v.byte_len();
    });
}
int mvf_space_get_vector_f32(const mvf_space* s, uint64_t index, float* out, size_t cap) {
    return guard([&] {
        Vector v = s->vs.get_vector(index);
        if (cap < v.dimension()) throw MvfError::dimension_mismatch(v.dimension(), cap);
        v.as_f32_into(out, cap);
    });
}
int mvf_space_map_vector_range(const mvf_space* s, uint64_t start, uint64_t count, const void** ptr, size_t* nbytes) {
    return guard([&] { *ptr = s->vs.map_vector_range(start, count, nbytes); });
}
int mvf_space_tombstones(const mvf_space* s, const void** bitmap, size_t* nbytes, uint64_t* deleted_count) {
    return guard([&] {
        size_t n = 0;
        const uint8_t* p = s->vs.tombstone_bitmap(&n);
        if (bitmap) *bitmap = p;
        if (nbytes) *nbytes = n;
        if (deleted_count) *deleted_count = s->vs.deleted_count();
    });
}
int mvf_space_tombstone_list(const mvf_space* s, const uint64_t** rows, size_t* count, uint64_t* deleted_count) {
    return guard([&] {
        size_t n = 0;
        const uint8_t* p = s->vs.tombstone_list(&n);
        if (rows) *rows = (const uint64_t*)p;            // blocks of u64 start at arbitrary file offsets: read with memcpy if alignment matters
        if (count) *count = n;
        if (deleted_count) *deleted_count = s->vs.deleted_count();
    });
}
int mvf_space_vector_ids(const mvf_space* s, const void** ids, size_t* count) {
    return guard([&] {
        size_t n = 0;
        const uint8_t* p = s->vs.vector_ids(&n);
        if (ids) *ids = p;
        if (count) *count = n;
    });
}
int mvf_file_metadata_column(const mvf_file* f, const char* name, const void** bytes, size_t* nbytes, int* data_type) {
    return guard([&] {
        size_t n = 0; DataType dt = DataType::UInt8;
        const uint8_t* p = f->r->metadata_column(name ? name : "", &n, &dt);
        if (bytes) *bytes = p;
        if (nbytes) *nbytes = n;
        if (data_type) *data_type = (int)dt;
    });
}
// Pre-filter helper: bit i of out_bitmap (LSB first) is SET (= row excluded, the convention of
// mvf_gpu_space_set_row_mask) when value i of the column lies outside [lo, hi].
int mvf_filter_range_bitmap(const void* column, int data_type, uint64_t n, double lo, double hi, uint8_t* out_bitmap) {
    return guard([&] {
        if (!column || !out_bitmap) throw MvfError::build_error("null argument");
        const uint8_t* c = (const uint8_t*)column;
        std::memset(out_bitmap, 0, (size_t)((n + 7) / 8));
        auto run = [&](auto load) {
            for (uint64_t i = 0; i < n; i++) {
                const double v = load(i);
                if (!(v >= lo && v <= hi)) out_bitmap[i >> 3] |= (uint8_t)(1u << (i & 7));
            }
        };
        switch ((DataType)data_type) {
            case DataType::Float32: run([&](uint64_t i) { float v; std::memcpy(&v, c + 4 * i, 4); return (double)v; }); break;
            case DataType::Float16: run([&](uint64_t i) { uint16_t h; std::memcpy(&h, c + 2 * i, 2); return (double)f16_to_f32(h); }); break;
            case DataType::Int8: run([&](uint64_t i) { return (double)(int8_t)c[i]; }); break;
            case DataType::UInt8: run([&](uint64_t i) { return (double)c[i]; }); break;
            case DataType::UInt32: run([&](uint64_t i) { uint32_t v; std::memcpy(&v, c + 4 * i, 4); return (double)v; }); break;
            case DataType::UInt64: run([&](uint64_t i) { uint64_t v; std::memcpy(&v, c + 8 * i, 8); return (double)v; }); break;
            default: throw MvfError::build_error("Unsupported metadata data type for a range filter");
        }
    });
}
int mvf_space_scannable_vectors(const mvf_space* s, uint64_t* out) { return guard([&] { *out = s->vs.scannable_vectors(); }); }

int mvf_builder_new(mvf_builder** out) { return guard([&] { *out = new mvf_builder(); }); }
void mvf_builder_free(mvf_builder* b) { delete b; }
int mvf_builder_add_vector_space(mvf_builder* b, const char* name, uint32_t dimension, int vector_type, int distance_metric,
                                 int data_type, size_t* out_index) {
    return guard([&] {
        size_t i = b->b.add_vector_space(name, dimension, (VectorType)vector_type, (DistanceMetric)distance_metric,
                                         (DataType)data_type);
        if (out_index) *out_index = i;
    });
}
int mvf_builder_add_vectors_f32(mvf_builder* b, const char* space, const float* rows, size_t nrows, uint32_t dimension) {
    return guard([&] { b->b.add_vectors(space, rows, nrows, dimension); });
}
int mvf_builder_add_vectors_raw(mvf_builder* b, const char* space, const void* rows, size_t nrows, uint32_t dimension) {
    return guard([&] { b->b.add_vectors_raw(space, rows, nrows, dimension); });
}
int mvf_builder_add_metadata_column(mvf_builder* b, const char* name, int data_type, const void* bytes, size_t nbytes) {
    return guard([&] { b->b.add_metadata_column(name, (DataType)data_type, bytes, nbytes); });
}
int mvf_builder_set_tombstones(mvf_builder* b, const char* space, const void* bitmap, size_t nbytes) {
    return guard([&] { b->b.set_tombstones(space, bitmap, nbytes); });
}
int mvf_builder_set_tombstone_list(mvf_builder* b, const char* space, const uint64_t* rows, size_t count) {
    return guard([&] { b->b.set_tombstone_list(space, rows, count); });
}
int mvf_builder_set_vector_ids(mvf_builder* b, const char* space, const uint64_t* ids, size_t count) {
    return guard([&] { b->b.set_vector_ids(space, ids, count); });
}
int mvf_builder_save(mvf_builder* b, const char* path) { return guard([&] { b->b.save(path); }); }

int mvf_shard_range(uint64_t n, int world, int rank, uint64_t* begin, uint64_t* count) {
    return guard([&] {
        if (world < 1 || rank < 0 || rank >= world) throw MvfError::build_error("invalid rank/world");
        shard_range(n, world, rank, begin, count);
    });
}

uint32_t mvf_ordered_key_f32(float v) {
    if (v != v) return 0xffffffffu;
    if (v == 0.0f) v = 0.0f;
    uint32_t b; std::memcpy(&b, &v, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
float mvf_ordered_key_to_f32(uint32_t key) {
    uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    float f; std::memcpy(&f, &b, 4); return f;
}

int mvf_synth_fill_host(void* out, int data_type, uint64_t seed, uint64_t row0, uint64_t nrows, uint32_t dimension) {
    return guard([&] {
        if (data_type < 0 || data_type > 3) throw MvfError::build_error("Unsupported vector data type");
        const uint64_t sk = synth::sm64(seed);
        const size_t es = element_size((DataType)data_type);
        for (uint64_t r = 0; r < nrows; r++) {
            const uint64_t rk = synth::row_key(sk, row0 + r);
            uint8_t* row = (uint8_t*)out + r * (size_t)dimension * es;
            if (data_type <= 1) {
                for (uint32_t c = 0; c < dimension; c += 2) {
                    float v0, v1; synth::float_pair(rk, c >> 1, &v0, &v1);
                    if (data_type == 0) {
                        std::memcpy(row + 4 * (size_t)c, &v0, 4);
                        if (c + 1 < dimension) std::memcpy(row + 4 * (size_t)(c + 1), &v1, 4);
                    } else {
                        uint16_t h0 = f32_to_f16(v0), h1 = f32_to_f16(v1);
                        std::memcpy(row + 2 * (size_t)c, &h0, 2);
                        if (c + 1 < dimension) std::memcpy(row + 2 * (size_t)(c + 1), &h1, 2);
                    }
                }
            } else {
                for (uint32_t c = 0; c < dimension; c += 8) {
                    uint64_t o = synth::byte_octet(rk, c >> 3);
                    uint32_t m = dimension - c < 8 ? dimension - c : 8;
                    std::memcpy(row + c, &o, m);
                }
            }
        }
    });
}

}  // extern "C"
