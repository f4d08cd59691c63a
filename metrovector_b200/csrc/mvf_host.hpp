// mvf_host.hpp -- C++ host side of libmvf_b200: the reader/writer API of thegenem0/metrovector
// under the reference's own names (the reference's toolchain, Rust, is absent from this image, so
// the host layer above the C ABI is C++).  Paths in comments are relative to the
// reference repository root.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace mvf {

// schema/types.fbs:3-25
enum class DataType : uint8_t { Float32 = 0, Float16 = 1, Int8 = 2, UInt8 = 3, UInt32 = 4, UInt64 = 5, StringRef = 6 };
enum class VectorType : uint8_t { Dense = 0, Sparse = 1 };
enum class DistanceMetric : uint8_t { L2 = 0, Euclidean = 0, InnerProduct = 1, Dot = 1, Cosine = 2, Custom = 255 };

// src/errors.rs:8-40 (declaration order == C status code), then ours
enum class ErrorKind : int {
    Ok = 0, Io, InvalidFormat, UnsupportedVersion, VectorSpaceNotFound, IndexOutOfBounds, DimensionMismatch,
    InvalidVectorType, CorruptedData, Extension, Build, Cuda, Nccl
};

class MvfError : public std::runtime_error {
public:
    MvfError(ErrorKind k, const std::string& msg, uint64_t a = 0, uint64_t b = 0) : std::runtime_error(msg), kind(k), first(a), second(b) {}
    ErrorKind kind;
    // the two fields of the reference's structured variants, in declaration order (src/errors.rs:15-31):
    // UnsupportedVersion{got, expected}, IndexOutOfBounds{index, len}, DimensionMismatch{expected, actual},
    // InvalidVectorType{expected, actual}; 0 for the string variants
    uint64_t first, second;
    // constructors producing the reference's Display strings (src/errors.rs:9-39)
    static MvfError io(const std::string& m) { return {ErrorKind::Io, "I/O error: " + m}; }
    static MvfError invalid_format(const std::string& m) { return {ErrorKind::InvalidFormat, "Invalid file format: " + m}; }
    static MvfError unsupported_version(unsigned got, unsigned expected) {
        return {ErrorKind::UnsupportedVersion,
                "Unsupported version: got " + std::to_string(got) + ", expected " + std::to_string(expected), got, expected};
    }
    static MvfError space_not_found(const std::string& n) { return {ErrorKind::VectorSpaceNotFound, "Vector space '" + n + "' not found"}; }
    static MvfError index_out_of_bounds(uint64_t index, uint64_t len) {
        return {ErrorKind::IndexOutOfBounds, "Index out of bounds: " + std::to_string(index) + " >= " + std::to_string(len), index, len};
    }
    static MvfError dimension_mismatch(uint64_t expected, uint64_t actual) {
        return {ErrorKind::DimensionMismatch,
                "Dimension mismatch: expected " + std::to_string(expected) + ", got " + std::to_string(actual), expected, actual};
    }
    static MvfError corrupted_data(const std::string& m) { return {ErrorKind::CorruptedData, "Corrupted data: " + m}; }
    static MvfError build_error(const std::string& m) { return {ErrorKind::Build, "Build error: " + m}; }
    static MvfError cuda(const std::string& m) { return {ErrorKind::Cuda, "CUDA error: " + m}; }
    static MvfError nccl(const std::string& m) { return {ErrorKind::Nccl, "NCCL error: " + m}; }
};

// bytes per component; 0 for dtypes that cannot hold vectors (src/vectors/vector_space.rs:122-127)
inline size_t element_size(DataType t) {
    switch (t) {
        case DataType::Float32: return 4;
        case DataType::Float16: return 2;
        case DataType::Int8: case DataType::UInt8: return 1;
        default: return 0;
    }
}

float f16_to_f32(uint16_t h);   // half::f16::to_f32 (exact)
uint16_t f32_to_f16(float f);   // half::f16::from_f32 (round-to-nearest-even)
uint32_t crc32_ieee(const void* data, size_t n, uint32_t crc = 0);   // crc32fast::hash

// struct DataBlock, schema/core.fbs:7-13
struct DataBlock {
    uint64_t offset = 0, size = 0;
    uint8_t compression = 0;
    uint64_t compressed_size = 0;
    uint32_t checksum = 0;
};

// table VectorSpace, schema/core.fbs:42-57 (fields the scan needs)
struct SpaceInfo {
    std::string name;
    uint32_t dimension = 0;
    uint64_t total_vectors = 0;
    VectorType vector_type = VectorType::Dense;
    DistanceMetric distance_metric = DistanceMetric::L2;
    DataType data_type = DataType::Float32;
    uint32_t vectors_block_index = 0;
    uint32_t vector_ids_block_index = 0;
    // table TombstoneInfo, schema/core.fbs:34-39 (never written by the reference's builder, src/builder.rs:486)
    uint8_t tombstone_format = 0;          // TombstoneFormat: None=0, Bitmap=1, SortedList=2 (schema/types.fbs:35-39)
    uint32_t tombstone_block_index = 0;    // 0 = no deletions
    uint64_t deleted_count = 0;
};

class MvfReader;

// Vector<'a>, src/vectors/vector.rs:28-33
class Vector {
public:
    Vector(const uint8_t* data, size_t nbytes, uint32_t dimension, DataType dt)
        : data_(data), nbytes_(nbytes), dimension_(dimension), data_type_(dt) {}
    uint32_t dimension() const { return dimension_; }                 // vector.rs:51
    DataType data_type() const { return data_type_; }                 // vector.rs:56
    const uint8_t* as_bytes() const { return data_; }                 // vector.rs:61
    size_t byte_len() const { return nbytes_; }
    std::vector<float> as_f32() const;                                // vector.rs:71-92
    void as_f32_into(float* out, size_t cap) const;
private:
    const uint8_t* data_;
    size_t nbytes_;
    uint32_t dimension_;
    DataType data_type_;
};

// VectorSpace<'a>, src/vectors/vector_space.rs:34-39
class VectorSpace {
public:
    VectorSpace(const MvfReader* r, size_t index) : reader_(r), index_(index) {}
    const std::string& name() const;            // :62
    uint32_t dimension() const;                 // :67
    uint64_t total_vectors() const;             // :72
    VectorType vector_type() const;             // :77
    DistanceMetric distance_metric() const;     // :82
    DataType data_type() const;                 // :87
    Vector get_vector(uint64_t index) const;    // :101-142
    // :155-188; returns pointer + byte length of `count` consecutive rows
    const uint8_t* map_vector_range(uint64_t start, uint64_t count, size_t* nbytes) const;
    VectorSpace clone_concurrent() const { return *this; }   // :194-201
    // rows the reference's `for i in 0..total_vectors()` loop can visit before get_vector errors
    uint64_t scannable_vectors() const;
    // deletion bitmap (bit i, LSB first, = row i deleted) if the footer carries TombstoneInfo{Bitmap}; nullptr otherwise
    const uint8_t* tombstone_bitmap(size_t* nbytes) const;
    // TombstoneInfo{SortedList}: ascending deleted row positions, u64 little-endian each (the schema leaves the element
    // type open and the reference never writes one; this is the definition both our writer and reader use)
    const uint8_t* tombstone_list(size_t* count) const;
    uint64_t deleted_count() const;
    // external ids (vector_ids_block_index, schema/core.fbs:54; u64 LE per row like src/builder.rs:257-271 serialises
    // them): nullptr when the footer says 0 = "use positions as IDs"
    const uint8_t* vector_ids(size_t* count) const;
    const MvfReader* reader() const { return reader_; }
private:
    const SpaceInfo& info() const;
    const uint8_t* block_data(size_t* len) const;
    const MvfReader* reader_;
    size_t index_;
};

// MvfReader, src/reader.rs:27-31.  README.md:97 calls it MvfFile.
class MvfReader {
public:
    static MvfReader* open(const std::string& path);                  // :45-79
    ~MvfReader();
    MvfReader(const MvfReader&) = delete;
    MvfReader& operator=(const MvfReader&) = delete;
    uint16_t version() const { return version_; }                     // :82
    size_t num_vector_spaces() const { return spaces_.size(); }       // :87
    std::vector<std::string> vector_space_names() const;              // :92-98
    VectorSpace vector_space(const std::string& name) const;          // :104-119
    uint64_t file_size() const { return size_; }                      // :122
    bool has_metadata() const { return has_metadata_; }               // :127
    const std::vector<std::string>& metadata_column_names() const { return metadata_names_; }   // :132-143
    // raw bytes of a metadata column (table MetadataColumn, schema/core.fbs:16-25; the reference reader stops at
    // the names): what a pre-filter is evaluated on.  Throws VectorSpaceNotFound-style Build error if absent.
    const uint8_t* metadata_column(const std::string& name, size_t* nbytes, DataType* dt) const;
    void validate() const;                                            // :149-162
    void validate_with_checksum() const;                              // :172-221 (intent)
    const uint8_t* data() const { return map_; }
    int fd() const { return fd_; }                                    // the open descriptor behind the mapping (residency re-maps ranges of it)
    const std::vector<SpaceInfo>& spaces() const { return spaces_; }
    const std::vector<DataBlock>& blocks() const { return blocks_; }
private:
    MvfReader() = default;
    void parse_footer(const uint8_t* footer, size_t len);
    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    uint64_t size_ = 0;
    uint16_t version_ = 0;
    bool has_metadata_ = false;
    std::vector<SpaceInfo> spaces_;
    std::vector<DataBlock> blocks_;
    std::vector<std::string> metadata_names_;
    std::vector<uint32_t> metadata_blocks_;
    std::vector<uint8_t> metadata_types_;
};
using MvfFile = MvfReader;

// ---- write side (src/builder.rs, src/io.rs) ------------------------------------------------
struct MetadataColumnBuilder {   // src/builder.rs:52-62
    std::string name;
    DataType data_type = DataType::UInt8;
    std::vector<uint8_t> data;
    uint64_t null_count = 0;
};
struct VectorSpaceBuilder {      // src/builder.rs:64-76
    std::string name;
    uint32_t dimension = 0;
    VectorType vector_type = VectorType::Dense;
    DistanceMetric distance_metric = DistanceMetric::L2;
    DataType data_type = DataType::Float32;
    std::vector<uint8_t> vectors;
    std::vector<uint8_t> tombstones;       // deletion bitmap, empty = none (extension: the reference never writes one)
    std::vector<uint8_t> tombstone_list;   // or: ascending deleted positions, u64 LE (TombstoneFormat::SortedList)
    std::vector<uint8_t> vector_ids;       // u64 LE external id per row (src/builder.rs:61,257-271), empty = positions
};

class MvfBuilder {               // src/builder.rs:93-98
public:
    size_t add_vector_space(const std::string& name, uint32_t dimension, VectorType vt, DistanceMetric dm,
                            DataType dt);                                                     // :112-135
    void add_vectors(const std::string& space, const float* rows, size_t nrows, uint32_t dimension);   // :151-196
    void add_vectors_raw(const std::string& space, const void* rows, size_t nrows, uint32_t dimension);
    void add_metadata_column(const std::string& name, DataType dt, const void* bytes, size_t n);       // :211-236
    void set_tombstones(const std::string& space, const void* bitmap, size_t nbytes);   // TombstoneInfo{Bitmap}, schema/core.fbs:34-39
    void set_tombstone_list(const std::string& space, const uint64_t* rows, size_t count);   // TombstoneInfo{SortedList}
    void set_vector_ids(const std::string& space, const uint64_t* ids, size_t count);   // VectorSpaceBuilder.vector_ids, src/builder.rs:61
    std::vector<uint8_t> to_bytes() const;      // build() + BuiltMvf::to_bytes, :241-308,417-558
    void save(const std::string& path) const;   // BuiltMvf::save + MvfWriter::write, io.rs:41-46
private:
    VectorSpaceBuilder& find(const std::string& name);
    std::vector<VectorSpaceBuilder> spaces_;
    std::vector<MetadataColumnBuilder> columns_;
};

}  // namespace mvf
