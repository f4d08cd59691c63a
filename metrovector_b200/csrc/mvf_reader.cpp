// mvf_reader.cpp -- mmap-backed MVF reader: MvfReader / VectorSpace / Vector.
// Behaviour follows src/reader.rs, src/vectors/vector_space.rs and src/vectors/vector.rs of the
// reference (line numbers beside each function).  The footer is decoded by a bounds-checked
// FlatBuffers table walker written for this file (the reference uses flatc-generated accessors;
// neither flatc nor a flatbuffers library exists in this image).
#include "mvf_host.hpp"

#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mvf {

static const uint8_t kMagic[4] = {'M', 'V', 'F', '1'};   // src/lib.rs:25
static const size_t kFooterSizeLen = 4;                  // src/lib.rs:26

// ---- conversions ------------------------------------------------------------------------------
float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int shift = __builtin_clz(man) - 21;   // bring the leading 1 to bit 10
            man = (man << shift) & 0x3ffu;
            bits = sign | ((uint32_t)(113 - shift) << 23) | (man << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

uint16_t f32_to_f16(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, man = x & 0x7fffffu;
    int exp = (int)((x >> 23) & 0xffu);
    if (exp == 255) return (uint16_t)(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0u));
    int e = exp - 112;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        unsigned shift = (unsigned)(14 - e);
        uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = man >> 13, rem = man & 0x1fffu;
    uint16_t out = (uint16_t)(sign | ((uint32_t)e << 10) | q);
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) out++;
    return out;
}

namespace {
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};
}  // namespace

uint32_t crc32_ieee(const void* data, size_t n, uint32_t crc) {
    // function-local static: initialised exactly once even when the first calls race (file handles are thread-safe)
    static const CrcTables tables;
    const uint32_t (*table)[256] = tables.t;
    const uint8_t* p = (const uint8_t*)data;
    crc = ~crc;
    while (n >= 8) {   // slicing-by-8
        uint32_t a, b; std::memcpy(&a, p, 4); std::memcpy(&b, p + 4, 4);
        a ^= crc;
        crc = table[7][a & 0xff] ^ table[6][(a >> 8) & 0xff] ^ table[5][(a >> 16) & 0xff] ^ table[4][a >> 24] ^
              table[3][b & 0xff] ^ table[2][(b >> 8) & 0xff] ^ table[1][(b >> 16) & 0xff] ^ table[0][b >> 24];
        p += 8; n -= 8;
    }
    while (n--) crc = table[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return ~crc;
}

// ---- bounds-checked FlatBuffers walker ---------------------------------------------------------
namespace {
struct Buf {
    const uint8_t* p; size_t n;
    void need(size_t off, size_t len) const {
        if (off > n || len > n - off) throw MvfError::invalid_format("Failed to parse footer: offset out of range");
    }
    template <class T> T rd(size_t off) const { need(off, sizeof(T)); T v; std::memcpy(&v, p + off, sizeof(T)); return v; }
};
bool valid_utf8(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        size_t extra;
        uint32_t cp;
        if (c < 0x80) { i++; continue; }
        else if ((c & 0xe0) == 0xc0) { extra = 1; cp = c & 0x1f; }
        else if ((c & 0xf0) == 0xe0) { extra = 2; cp = c & 0x0f; }
        else if ((c & 0xf8) == 0xf0) { extra = 3; cp = c & 0x07; }
        else return false;
        if (i + extra >= n) return false;
        for (size_t j = 1; j <= extra; j++) {
            if ((s[i + j] & 0xc0) != 0x80) return false;
            cp = (cp << 6) | (s[i + j] & 0x3f);
        }
        if ((extra == 1 && cp < 0x80) || (extra == 2 && cp < 0x800) || (extra == 3 && (cp < 0x10000 || cp > 0x10ffff)) ||
            (cp >= 0xd800 && cp <= 0xdfff)) return false;
        i += extra + 1;
    }
    return true;
}
struct Table {
    const Buf* b; size_t pos = 0, vt = 0; uint16_t vt_size = 0, tbl_size = 0;
    Table(const Buf* buf, size_t at) : b(buf), pos(at) {
        int32_t soff = b->rd<int32_t>(pos);
        int64_t v = (int64_t)pos - (int64_t)soff;
        if (v < 0 || (uint64_t)v > b->n) throw MvfError::invalid_format("Failed to parse footer: bad vtable offset");
        vt = (size_t)v;
        vt_size = b->rd<uint16_t>(vt);
        tbl_size = b->rd<uint16_t>(vt + 2);
        if (vt_size < 4 || (vt_size & 1)) throw MvfError::invalid_format("Failed to parse footer: bad vtable size");
        b->need(vt, vt_size);
        b->need(pos, tbl_size < 4 ? 4 : tbl_size);
    }
    size_t field(unsigned slot) const {   // absolute position of the field, 0 if absent
        size_t o = 4 + 2 * (size_t)slot;
        if (o + 2 > vt_size) return 0;
        uint16_t f = b->rd<uint16_t>(vt + o);
        return f ? pos + f : 0;
    }
    template <class T> T scalar(unsigned slot, T def) const { size_t f = field(slot); return f ? b->rd<T>(f) : def; }
    size_t indirect(unsigned slot) const {   // target of an offset field, 0 if absent
        size_t f = field(slot);
        if (!f) return 0;
        uint32_t rel = b->rd<uint32_t>(f);
        size_t t = f + rel;
        b->need(t, 4);
        return t;
    }
    std::string str(unsigned slot, bool required) const {
        size_t t = indirect(slot);
        if (!t) {
            if (required) throw MvfError::invalid_format("Failed to parse footer: missing required string");
            return std::string();
        }
        uint32_t len = b->rd<uint32_t>(t);
        b->need(t + 4, (size_t)len + 1);
        const uint8_t* s = b->p + t + 4;
        // like the flatbuffers verifier the reference runs (src/reader.rs:245): NUL-terminated, valid UTF-8
        if (s[len] != 0 || !valid_utf8(s, len)) throw MvfError::invalid_format("Failed to parse footer: invalid string");
        return std::string((const char*)s, len);
    }
};
}  // namespace

void MvfReader::parse_footer(const uint8_t* footer, size_t len) {
    Buf b{footer, len};
    uint32_t root = b.rd<uint32_t>(0);
    Table ft(&b, root);
    // schema/mvf.fbs:12-30
    version_ = ft.scalar<uint16_t>(0, 3);
    size_t vs = ft.indirect(1), bm = ft.indirect(2), md = ft.indirect(3);
    if (!vs || !bm) throw MvfError::invalid_format("Failed to parse footer: missing required field");
    uint32_t nb = b.rd<uint32_t>(bm);
    b.need(bm + 4, (size_t)nb * 40);
    blocks_.resize(nb);
    for (uint32_t i = 0; i < nb; i++) {   // schema/core.fbs:7-13, 40-byte struct
        size_t at = bm + 4 + (size_t)i * 40;
        blocks_[i].offset = b.rd<uint64_t>(at);
        blocks_[i].size = b.rd<uint64_t>(at + 8);
        blocks_[i].compression = b.rd<uint8_t>(at + 16);
        blocks_[i].compressed_size = b.rd<uint64_t>(at + 24);
        blocks_[i].checksum = b.rd<uint32_t>(at + 32);
    }
    uint32_t ns = b.rd<uint32_t>(vs);
    b.need(vs + 4, (size_t)ns * 4);
    spaces_.resize(ns);
    for (uint32_t i = 0; i < ns; i++) {   // schema/core.fbs:42-57
        size_t p = vs + 4 + (size_t)i * 4;
        Table t(&b, p + b.rd<uint32_t>(p));
        SpaceInfo& s = spaces_[i];
        s.name = t.str(0, true);
        s.dimension = t.scalar<uint32_t>(1, 0);
        s.total_vectors = t.scalar<uint64_t>(2, 0);
        s.vector_type = (VectorType)t.scalar<uint8_t>(3, 0);
        s.distance_metric = (DistanceMetric)t.scalar<uint8_t>(4, 0);
        s.data_type = (DataType)t.scalar<uint8_t>(5, 0);
        s.vectors_block_index = t.scalar<uint32_t>(6, 0);
        s.vector_ids_block_index = t.scalar<uint32_t>(9, 0);
        if (size_t tomb = t.indirect(11)) {                       // TombstoneInfo, schema/core.fbs:34-39
            Table tt(&b, tomb + 0);
            s.tombstone_format = tt.scalar<uint8_t>(0, 0);
            s.tombstone_block_index = tt.scalar<uint32_t>(1, 0);
            s.deleted_count = tt.scalar<uint64_t>(2, 0);
        }
    }
    has_metadata_ = md != 0;
    if (md) {
        uint32_t nm = b.rd<uint32_t>(md);
        b.need(md + 4, (size_t)nm * 4);
        for (uint32_t i = 0; i < nm; i++) {   // schema/core.fbs:16-25
            size_t p = md + 4 + (size_t)i * 4;
            Table t(&b, p + b.rd<uint32_t>(p));
            metadata_names_.push_back(t.str(0, true));
            metadata_types_.push_back(t.scalar<uint8_t>(1, 0));
            metadata_blocks_.push_back(t.scalar<uint32_t>(2, 0));
        }
    }
}

// src/reader.rs:45-79 with validate_file_structure (:259-278) and validate_footer_bounds (:225-256)
MvfReader* MvfReader::open(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) throw MvfError::io(std::string(std::strerror(errno)) + " (" + path + ")");
    struct stat st;
    if (fstat(fd, &st) != 0) { int e = errno; ::close(fd); throw MvfError::io(std::strerror(e)); }
    if (S_ISDIR(st.st_mode)) { ::close(fd); throw MvfError::io("Is a directory (" + path + ")"); }
    uint64_t size = (uint64_t)st.st_size;
    const uint8_t* map = nullptr;
    if (size > 0) {
        void* m = mmap(nullptr, size, PROT_READ, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) { int e = errno; ::close(fd); throw MvfError::io(std::strerror(e)); }
        map = (const uint8_t*)m;
    }
    MvfReader* r = new MvfReader();
    r->fd_ = fd; r->map_ = map; r->size_ = size;
    try {
        if (size < sizeof(kMagic) + kFooterSizeLen + sizeof(kMagic)) throw MvfError::invalid_format("File too small");
        if (std::memcmp(map, kMagic, 4) != 0) throw MvfError::invalid_format("Invalid magic bytes at start of file");
        if (std::memcmp(map + size - 4, kMagic, 4) != 0)
            throw MvfError::invalid_format("Invalid magic bytes at end of file, it may be corrupted");
        uint32_t footer_len; std::memcpy(&footer_len, map + size - 8, 4);   // little-endian host
        if ((uint64_t)footer_len + 8 > size - 4) throw MvfError::invalid_format("Invalid footer length");
        size_t footer_start = size - 8 - footer_len;
        r->parse_footer(map + footer_start, footer_len);
        if (r->version_ != 1) throw MvfError::unsupported_version(r->version_, 1);
    } catch (...) { delete r; throw; }
    return r;
}

MvfReader::~MvfReader() {
    if (map_) munmap((void*)map_, size_);
    if (fd_ >= 0) ::close(fd_);
}

std::vector<std::string> MvfReader::vector_space_names() const {
    std::vector<std::string> out;
    for (const auto& s : spaces_) out.push_back(s.name);
    return out;
}

VectorSpace MvfReader::vector_space(const std::string& name) const {   // :104-119, first match wins
    for (size_t i = 0; i < spaces_.size(); i++)
        if (spaces_[i].name == name) return VectorSpace(this, i);
    throw MvfError::space_not_found(name);
}

void MvfReader::validate() const {   // :149-162 (the reference compares offset+size with file length as is)
    for (const auto& blk : blocks_) {
        if (blk.offset + blk.size > size_ || blk.offset + blk.size < blk.offset)
            throw MvfError::corrupted_data("Block extends beyond file: offset=" + std::to_string(blk.offset) +
                                           ", size=" + std::to_string(blk.size) + ", file_size=" + std::to_string(size_));
    }
}

// The reference's validate_with_checksum (:172-221) ends in todo!() and subtracts the magic
// length where the layout adds it (:185-191).  This is the finished intent: CRC-32/IEEE of the
// block bytes at 4+offset, compared with the manifest (written by src/builder.rs:251).
void MvfReader::validate_with_checksum() const {
    validate();
    for (const auto& blk : blocks_) {
        if (blk.checksum == 0) continue;
        if (4 + blk.offset + blk.size > size_)
            throw MvfError::corrupted_data("Block extends beyond file: offset=" + std::to_string(blk.offset) +
                                           ", size=" + std::to_string(blk.size) + ", file_size=" + std::to_string(size_));
        uint32_t actual = crc32_ieee(map_ + 4 + blk.offset, blk.size);
        if (actual != blk.checksum)
            throw MvfError::corrupted_data("Block checksum mismatch: expected " + std::to_string(blk.checksum) +
                                           ", got " + std::to_string(actual));
    }
}

// ---- VectorSpace ----------------------------------------------------------------------------
const SpaceInfo& VectorSpace::info() const { return reader_->spaces()[index_]; }
const std::string& VectorSpace::name() const { return info().name; }
uint32_t VectorSpace::dimension() const { return info().dimension; }
uint64_t VectorSpace::total_vectors() const { return info().total_vectors; }
VectorType VectorSpace::vector_type() const { return info().vector_type; }
DistanceMetric VectorSpace::distance_metric() const { return info().distance_metric; }
DataType VectorSpace::data_type() const { return info().data_type; }

// block bytes = file[4+offset .. 4+offset+size]   (vector_space.rs:109-120).  The reference slices
// without checking the file length and panics on a truncated file; we return CorruptedData.
const uint8_t* VectorSpace::block_data(size_t* len) const {
    const auto& blocks = reader_->blocks();
    uint32_t bi = info().vectors_block_index;
    if (bi >= blocks.size()) throw MvfError::corrupted_data("Invalid vector block index");
    const DataBlock& blk = blocks[bi];
    uint64_t start = 4 + blk.offset;
    if (start < blk.offset || start > reader_->file_size() || blk.size > reader_->file_size() - start)
        throw MvfError::corrupted_data("Block extends beyond file: offset=" + std::to_string(blk.offset) +
                                       ", size=" + std::to_string(blk.size) +
                                       ", file_size=" + std::to_string(reader_->file_size()));
    *len = (size_t)blk.size;
    return reader_->data() + start;
}

Vector VectorSpace::get_vector(uint64_t index) const {   // vector_space.rs:101-142
    if (index >= total_vectors()) throw MvfError::index_out_of_bounds(index, total_vectors());
    size_t blen; const uint8_t* bd = block_data(&blen);
    size_t es = element_size(data_type());
    if (es == 0) throw MvfError::build_error("Unsupported vector data type");
    uint64_t vsize = (uint64_t)dimension() * es;
    // index * vsize may exceed 64 bits for hostile footers: compare by division instead
    if (vsize == 0 ? false : (index > (UINT64_MAX - vsize) / vsize)) throw MvfError::index_out_of_bounds(index, blen / vsize);
    uint64_t off = index * vsize;
    if (off + vsize > blen) throw MvfError::index_out_of_bounds(index, vsize ? blen / vsize : 0);
    return Vector(bd + off, (size_t)vsize, dimension(), data_type());
}

const uint8_t* VectorSpace::map_vector_range(uint64_t start, uint64_t count, size_t* nbytes) const {   // :155-188
    if (start + count > total_vectors() || start + count < start) throw MvfError::index_out_of_bounds(start + count, total_vectors());
    size_t blen; const uint8_t* bd = block_data(&blen);
    size_t es = element_size(data_type());
    if (es == 0) throw MvfError::build_error("Unsupported vector data type");
    uint64_t vsize = (uint64_t)dimension() * es;
    if (vsize && (start + count) > UINT64_MAX / vsize) throw MvfError::corrupted_data("Vector range out of bounds");
    uint64_t off = start * vsize, len = count * vsize;
    if (off + len > blen) throw MvfError::corrupted_data("Vector range out of bounds");
    *nbytes = (size_t)len;
    return bd + off;
}

uint64_t VectorSpace::scannable_vectors() const {
    size_t blen; block_data(&blen);
    size_t es = element_size(data_type());
    if (es == 0) throw MvfError::build_error("Unsupported vector data type");
    uint64_t vsize = (uint64_t)dimension() * es;
    if (vsize == 0) return 0;
    uint64_t fit = blen / vsize;
    return total_vectors() < fit ? total_vectors() : fit;
}

uint64_t VectorSpace::deleted_count() const { return info().deleted_count; }

// bounds-checked view of an auxiliary block (tombstones, ids) named by index in the manifest
static const uint8_t* aux_block(const MvfReader* r, uint32_t index, const char* what, size_t* nbytes) {
    const auto& blocks = r->blocks();
    if (index >= blocks.size()) throw MvfError::corrupted_data(std::string("Invalid ") + what + " block index");
    const DataBlock& blk = blocks[index];
    uint64_t start = 4 + blk.offset;
    if (start > r->file_size() || blk.size > r->file_size() - start)
        throw MvfError::corrupted_data("Block extends beyond file: offset=" + std::to_string(blk.offset) + ", size=" +
                                       std::to_string(blk.size) + ", file_size=" + std::to_string(r->file_size()));
    *nbytes = (size_t)blk.size;
    return r->data() + start;
}

const uint8_t* MvfReader::metadata_column(const std::string& name, size_t* nbytes, DataType* dt) const {
    for (size_t i = 0; i < metadata_names_.size(); i++) {
        if (metadata_names_[i] != name) continue;
        if (dt) *dt = (DataType)metadata_types_[i];
        return aux_block(this, metadata_blocks_[i], "metadata column", nbytes);
    }
    throw MvfError::build_error("Metadata column '" + name + "' not found");
}

const uint8_t* VectorSpace::tombstone_bitmap(size_t* nbytes) const {
    *nbytes = 0;
    const SpaceInfo& si = info();
    if (si.tombstone_format != 1 || si.tombstone_block_index == 0) return nullptr;
    return aux_block(reader_, si.tombstone_block_index, "tombstone", nbytes);
}

const uint8_t* VectorSpace::tombstone_list(size_t* count) const {
    *count = 0;
    const SpaceInfo& si = info();
    if (si.tombstone_format != 2 || si.tombstone_block_index == 0) return nullptr;
    size_t nb = 0;
    const uint8_t* p = aux_block(reader_, si.tombstone_block_index, "tombstone", &nb);
    if (nb % 8) throw MvfError::corrupted_data("Tombstone list is not a whole number of 64-bit positions");
    *count = nb / 8;
    return p;
}

const uint8_t* VectorSpace::vector_ids(size_t* count) const {
    *count = 0;
    const SpaceInfo& si = info();
    if (si.vector_ids_block_index == 0) return nullptr;              // 0 = use positions as IDs (schema/core.fbs:54)
    size_t nb = 0;
    const uint8_t* p = aux_block(reader_, si.vector_ids_block_index, "vector ids", &nb);
    if (nb / 8 < scannable_vectors()) throw MvfError::corrupted_data("Vector id block is shorter than the vector space");
    *count = nb / 8;
    return p;
}

// ---- Vector ------------------------------------------------------------------------------------
void Vector::as_f32_into(float* out, size_t cap) const {   // vector.rs:71-92
    if (data_type_ == DataType::Float32) {
        size_t n = nbytes_ / 4; if (n > cap) n = cap;
        std::memcpy(out, data_, n * 4);   // f32::from_le_bytes on a little-endian host
    } else if (data_type_ == DataType::Float16) {
        size_t n = nbytes_ / 2; if (n > cap) n = cap;
        for (size_t j = 0; j < n; j++) { uint16_t h; std::memcpy(&h, data_ + 2 * j, 2); out[j] = f16_to_f32(h); }
    } else {
        throw MvfError::build_error("Cannot convert to f32");
    }
}

std::vector<float> Vector::as_f32() const {
    size_t es = data_type_ == DataType::Float32 ? 4 : data_type_ == DataType::Float16 ? 2 : 0;
    if (!es) throw MvfError::build_error("Cannot convert to f32");
    std::vector<float> out(nbytes_ / es);
    as_f32_into(out.data(), out.size());
    return out;
}

}  // namespace mvf
