// mvf_builder.cpp -- write side: MvfBuilder -> bytes -> file.
// Layout per src/builder.rs:241-308 (block manifest) and :417-558 (framing + footer tables) of the
// reference.  Differences, both deliberate: total_vectors is computed with the space's real
// element size (the reference divides by dimension*4 for every dtype, :476), and rows can be
// appended already encoded (add_vectors_raw) so Int8/UInt8 spaces can be written (:192 cannot).
// The footer is emitted by a small back-to-front FlatBuffers encoder written for this file.
#include "mvf_host.hpp"

#include <cerrno>
#include <cstdio>
#include <cstring>

namespace mvf {

namespace {
// Back-to-front FlatBuffers encoder: `buf` holds the finished tail of the buffer; offsets are
// measured from the END of the buffer, which stays fixed while the front grows.
class FbEncoder {
public:
    size_t off() const { return buf.size(); }
    void prep(size_t size, size_t additional) {
        if (size > minalign) minalign = size;
        size_t pad = (size - ((buf.size() + additional) % size)) % size;
        buf.insert(buf.begin(), pad, 0);
    }
    template <class T> void put(T v) {
        prep(sizeof(T), 0);
        uint8_t tmp[sizeof(T)]; std::memcpy(tmp, &v, sizeof(T));
        buf.insert(buf.begin(), tmp, tmp + sizeof(T));
    }
    void put_uoffset(size_t target) {
        prep(4, 0);
        uint32_t rel = (uint32_t)(buf.size() + 4 - target);
        uint8_t tmp[4]; std::memcpy(tmp, &rel, 4);
        buf.insert(buf.begin(), tmp, tmp + 4);
    }
    size_t string(const std::string& s) {
        prep(4, s.size() + 1);
        buf.insert(buf.begin(), (uint8_t)0);
        buf.insert(buf.begin(), s.begin(), s.end());
        uint32_t n = (uint32_t)s.size(); uint8_t tmp[4]; std::memcpy(tmp, &n, 4);
        buf.insert(buf.begin(), tmp, tmp + 4);
        return off();
    }
    size_t offset_vector(const std::vector<size_t>& offs) {
        prep(4, 4 * offs.size());
        for (size_t i = offs.size(); i-- > 0;) put_uoffset(offs[i]);
        raw_u32((uint32_t)offs.size());
        return off();
    }
    size_t struct_vector(const std::vector<uint8_t>& bytes, size_t elem, size_t align) {
        prep(4, bytes.size()); prep(align, bytes.size());
        buf.insert(buf.begin(), bytes.begin(), bytes.end());
        raw_u32((uint32_t)(bytes.size() / elem));
        return off();
    }
    void start(size_t nslots) { fields.assign(nslots, 0); obj_start = off(); }
    template <class T> void slot(size_t i, T v, T def) { if (v == def) return; put<T>(v); fields[i] = off(); }
    void slot_off(size_t i, size_t target) { if (!target) return; put_uoffset(target); fields[i] = off(); }
    size_t end() {
        put<int32_t>(0);
        size_t table = off();
        while (!fields.empty() && fields.back() == 0) fields.pop_back();
        std::vector<uint8_t> vt(4 + 2 * fields.size());
        uint16_t a = (uint16_t)vt.size(), b = (uint16_t)(table - obj_start);
        std::memcpy(&vt[0], &a, 2); std::memcpy(&vt[2], &b, 2);
        for (size_t i = 0; i < fields.size(); i++) {
            uint16_t f = fields[i] ? (uint16_t)(table - fields[i]) : 0;
            std::memcpy(&vt[4 + 2 * i], &f, 2);
        }
        size_t found = 0;
        for (size_t vo : vtables) {
            size_t at = buf.size() - vo;
            if (at + vt.size() <= buf.size() && std::memcmp(&buf[at], vt.data(), vt.size()) == 0) { found = vo; break; }
        }
        if (!found) {
            prep(2, vt.size());
            buf.insert(buf.begin(), vt.begin(), vt.end());
            found = off();
            vtables.push_back(found);
        }
        int32_t soff = (int32_t)((long)found - (long)table);   // table_pos - vtable_pos
        std::memcpy(&buf[buf.size() - table], &soff, 4);
        return table;
    }
    std::vector<uint8_t> finish_minimal(size_t root) { prep(minalign, 4); put_uoffset(root); return buf; }
private:
    void raw_u32(uint32_t v) { uint8_t tmp[4]; std::memcpy(tmp, &v, 4); buf.insert(buf.begin(), tmp, tmp + 4); }
    std::vector<uint8_t> buf;
    std::vector<size_t> fields, vtables;
    size_t obj_start = 0, minalign = 1;
};

void put_block(std::vector<uint8_t>& out, const DataBlock& b) {   // schema/core.fbs:7-13
    uint8_t s[40] = {0};
    std::memcpy(s + 0, &b.offset, 8); std::memcpy(s + 8, &b.size, 8); s[16] = b.compression;
    std::memcpy(s + 24, &b.compressed_size, 8); std::memcpy(s + 32, &b.checksum, 4);
    out.insert(out.end(), s, s + 40);
}
}  // namespace

VectorSpaceBuilder& MvfBuilder::find(const std::string& name) {
    for (auto& s : spaces_) if (s.name == name) return s;
    throw MvfError::space_not_found(name);
}

size_t MvfBuilder::add_vector_space(const std::string& name, uint32_t dimension, VectorType vt, DistanceMetric dm,
                                    DataType dt) {
    VectorSpaceBuilder s; s.name = name; s.dimension = dimension; s.vector_type = vt; s.distance_metric = dm; s.data_type = dt;
    spaces_.push_back(std::move(s));
    return spaces_.size() - 1;
}

void MvfBuilder::add_vectors(const std::string& space, const float* rows, size_t nrows, uint32_t dimension) {
    VectorSpaceBuilder& s = find(space);
    if (nrows == 0) return;
    if (s.dimension == 0) s.dimension = dimension;                         // builder.rs:168-170
    else if (s.dimension != dimension) throw MvfError::dimension_mismatch(s.dimension, dimension);
    size_t n = nrows * (size_t)dimension;
    if (s.data_type == DataType::Float32) {
        const uint8_t* p = (const uint8_t*)rows;
        s.vectors.insert(s.vectors.end(), p, p + n * 4);
    } else if (s.data_type == DataType::Float16) {
        size_t at = s.vectors.size(); s.vectors.resize(at + n * 2);
        for (size_t i = 0; i < n; i++) { uint16_t h = f32_to_f16(rows[i]); std::memcpy(&s.vectors[at + 2 * i], &h, 2); }
    } else {
        throw MvfError::build_error("Unsupported data type for vectors");   // builder.rs:192
    }
}

void MvfBuilder::add_vectors_raw(const std::string& space, const void* rows, size_t nrows, uint32_t dimension) {
    VectorSpaceBuilder& s = find(space);
    if (nrows == 0) return;
    size_t es = element_size(s.data_type);
    if (!es) throw MvfError::build_error("Unsupported data type for vectors");
    if (s.dimension == 0) s.dimension = dimension;
    else if (s.dimension != dimension) throw MvfError::dimension_mismatch(s.dimension, dimension);
    const uint8_t* p = (const uint8_t*)rows;
    s.vectors.insert(s.vectors.end(), p, p + nrows * (size_t)dimension * es);
}

void MvfBuilder::add_metadata_column(const std::string& name, DataType dt, const void* bytes, size_t n) {
    MetadataColumnBuilder c; c.name = name; c.data_type = dt;
    c.data.assign((const uint8_t*)bytes, (const uint8_t*)bytes + n);
    columns_.push_back(std::move(c));
}

void MvfBuilder::set_tombstones(const std::string& space, const void* bitmap, size_t nbytes) {
    VectorSpaceBuilder& s = find(space);
    s.tombstones.assign((const uint8_t*)bitmap, (const uint8_t*)bitmap + nbytes);
    s.tombstone_list.clear();
}

void MvfBuilder::set_tombstone_list(const std::string& space, const uint64_t* rows, size_t count) {
    VectorSpaceBuilder& s = find(space);
    for (size_t i = 1; i < count; i++)
        if (rows[i] <= rows[i - 1]) throw MvfError::build_error("Tombstone list must be strictly ascending");
    s.tombstone_list.assign((const uint8_t*)rows, (const uint8_t*)rows + count * 8);   // little-endian host
    s.tombstones.clear();
}

void MvfBuilder::set_vector_ids(const std::string& space, const uint64_t* ids, size_t count) {
    VectorSpaceBuilder& s = find(space);
    size_t es = element_size(s.data_type);
    uint64_t total = (s.dimension && es) ? s.vectors.size() / ((uint64_t)s.dimension * es) : 0;
    if (count != total) throw MvfError::dimension_mismatch(total, count);
    s.vector_ids.assign((const uint8_t*)ids, (const uint8_t*)ids + count * 8);
}

std::vector<uint8_t> MvfBuilder::to_bytes() const {
    // side tables are checked against the rows present NOW: rows may have been added after set_vector_ids /
    // set_tombstones, and a file whose id block is shorter than its space is rejected by every reader
    for (const auto& s : spaces_) {
        const size_t es = element_size(s.data_type);
        const uint64_t total = (s.dimension && es) ? s.vectors.size() / ((uint64_t)s.dimension * es) : 0;
        if (!s.vector_ids.empty() && s.vector_ids.size() / 8 != total) throw MvfError::dimension_mismatch(total, s.vector_ids.size() / 8);
        if (!s.tombstones.empty() && (uint64_t)s.tombstones.size() * 8 < total)
            throw MvfError::dimension_mismatch(total, (uint64_t)s.tombstones.size() * 8);
        if (!s.tombstone_list.empty()) {
            uint64_t last; std::memcpy(&last, s.tombstone_list.data() + s.tombstone_list.size() - 8, 8);
            if (last >= total) throw MvfError::index_out_of_bounds(last, total);
        }
    }
    std::vector<uint8_t> out = {'M', 'V', 'F', '1'};
    std::vector<uint8_t> manifest;
    uint64_t cur = 0;
    auto add_block = [&](const std::vector<uint8_t>& data) {
        DataBlock b; b.offset = cur; b.size = data.size(); b.checksum = crc32_ieee(data.data(), data.size());
        put_block(manifest, b);
        out.insert(out.end(), data.begin(), data.end());
        cur += data.size();
    };
    for (const auto& s : spaces_) add_block(s.vectors);
    for (const auto& c : columns_) add_block(c.data);
    // deletion bitmaps go after the metadata columns (the reference never emits any, so the position is ours)
    std::vector<uint32_t> tomb_block(spaces_.size(), 0), ids_block(spaces_.size(), 0);
    {
        uint32_t next = (uint32_t)(spaces_.size() + columns_.size());
        for (size_t i = 0; i < spaces_.size(); i++)
            if (!spaces_[i].tombstones.empty()) { add_block(spaces_[i].tombstones); tomb_block[i] = next++; }
            else if (!spaces_[i].tombstone_list.empty()) { add_block(spaces_[i].tombstone_list); tomb_block[i] = next++; }
        // external ids last: the reference's build() would interleave them after each vector block (builder.rs:257-271)
        // but then points vectors_block_index at the wrong block (:481); any position is legal through the index
        for (size_t i = 0; i < spaces_.size(); i++)
            if (!spaces_[i].vector_ids.empty()) { add_block(spaces_[i].vector_ids); ids_block[i] = next++; }
    }

    FbEncoder fb;
    std::vector<size_t> space_offs;
    for (size_t i = 0; i < spaces_.size(); i++) {
        const auto& s = spaces_[i];
        size_t name = fb.string(s.name);
        fb.start(0);
        size_t flat = fb.end();   // FlatIndex {} (builder.rs:465-468)
        size_t es = element_size(s.data_type);
        uint64_t total = (s.dimension && es) ? s.vectors.size() / ((uint64_t)s.dimension * es) : 0;
        size_t tomb = 0;
        if (tomb_block[i]) {
            const bool list = s.tombstones.empty();
            uint64_t deleted = list ? s.tombstone_list.size() / 8 : 0;
            for (uint8_t byte : s.tombstones) deleted += (uint64_t)__builtin_popcount(byte);
            fb.start(3);
            fb.slot<uint64_t>(2, deleted, 0);
            fb.slot<uint32_t>(1, tomb_block[i], 0);
            fb.slot<uint8_t>(0, list ? 2 : 1, 0);             // TombstoneFormat::Bitmap / SortedList (schema/types.fbs:35-39)
            tomb = fb.end();
        }
        fb.start(12);
        fb.slot<uint64_t>(2, total, 0);
        fb.slot_off(11, tomb);
        fb.slot<uint32_t>(9, ids_block[i], 0);               // vector_ids_block_index, 0 = positions
        fb.slot_off(8, flat);
        fb.slot<uint32_t>(6, (uint32_t)i, 0);
        fb.slot<uint32_t>(1, s.dimension, 0);
        fb.slot_off(0, name);
        fb.slot<uint8_t>(7, 1, 0);   // Index::FlatIndex
        fb.slot<uint8_t>(5, (uint8_t)s.data_type, 0);
        fb.slot<uint8_t>(4, (uint8_t)s.distance_metric, 0);
        fb.slot<uint8_t>(3, (uint8_t)s.vector_type, 0);
        space_offs.push_back(fb.end());
    }
    size_t vs = fb.offset_vector(space_offs);
    size_t bm = fb.struct_vector(manifest, 40, 8);
    size_t md = 0;
    if (!columns_.empty()) {
        std::vector<size_t> col_offs;
        for (size_t i = 0; i < columns_.size(); i++) {
            size_t name = fb.string(columns_[i].name);
            fb.start(6);
            fb.slot<uint64_t>(3, columns_[i].null_count, 0);
            fb.slot<uint32_t>(2, (uint32_t)(spaces_.size() + i), 0);
            fb.slot_off(0, name);
            fb.slot<uint8_t>(1, (uint8_t)columns_[i].data_type, 0);
            col_offs.push_back(fb.end());
        }
        md = fb.offset_vector(col_offs);
    }
    fb.start(8);
    fb.slot_off(3, md);
    fb.slot_off(2, bm);
    fb.slot_off(1, vs);
    fb.slot<uint16_t>(6, 1, 3);   // compatibility_version, default 3 (schema/mvf.fbs:28)
    fb.slot<uint16_t>(0, 1, 3);   // format_version = 1 (builder.rs:531), default 3
    std::vector<uint8_t> footer = fb.finish_minimal(fb.end());
    out.insert(out.end(), footer.begin(), footer.end());
    uint32_t flen = (uint32_t)footer.size(); uint8_t tmp[4]; std::memcpy(tmp, &flen, 4);
    out.insert(out.end(), tmp, tmp + 4);
    out.insert(out.end(), {'M', 'V', 'F', '1'});
    return out;
}

void MvfBuilder::save(const std::string& path) const {
    std::vector<uint8_t> bytes = to_bytes();
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw MvfError::io(std::string(std::strerror(errno)) + " (" + path + ")");
    errno = 0;
    const size_t w = std::fwrite(bytes.data(), 1, bytes.size(), f);
    if (w != bytes.size()) {
        const int e = errno;                         // of the failing fwrite, before fclose can overwrite it
        std::fclose(f);
        throw MvfError::io(std::string(e ? std::strerror(e) : "short write") + " (" + path + ")");
    }
    if (std::fclose(f) != 0) { const int e = errno; throw MvfError::io(std::string(std::strerror(e)) + " (" + path + ")"); }
}

}  // namespace mvf
