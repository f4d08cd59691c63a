/*
 * mvf_b200.h -- C ABI of libmvf_b200.so: the drop-in boundary for the brute-force scan that
 * thegenem0/metrovector's examples/similarity_search.rs performs on the CPU.
 *
 * The reference is a Rust library with no FFI of its own; its "API" for this path is the set of
 * public reader methods plus the example's free function `find_top_k_similar`.  Every entry
 * point below names the reference interface it replaces (paths relative to the reference
 * repository root).  The Rust `search` module a maintainer would add binds exactly these symbols
 * (see INTEGRATION.md).  Plain pointers and sizes only; no C++ or torch types cross this line.
 *
 * Conventions
 *   - every function returns an int status (mvf_status); 0 = OK.  1..10 are the variants of
 *     `MvfError` in declaration order (src/errors.rs:8-40); 11 = CUDA, 12 = NCCL.
 *   - mvf_last_error() returns the thread-local message of the last failing call on this
 *     thread, formatted like the reference's Display strings (src/errors.rs:9-39).
 *   - no exception or abort crosses the ABI.
 *   - handles are opaque.  mvf_space* and row pointers borrow from their mvf_file and die with
 *     it (src/vectors/vector_space.rs:34-39).  mvf_gpu_space* owns its HBM copy and outlives
 *     the file it was uploaded from.
 *   - file/space handles are immutable and thread-safe (src/reader.rs:284-289); every entry point that
 *     touches an engine's stream or scratch takes the engine's lock, so calls on one mvf_gpu from several
 *     host threads are serialised by the library (one stream + scratch per device).
 */
#ifndef MVF_B200_H
#define MVF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVF_API __attribute__((visibility("default")))

/* src/errors.rs:8-40, in declaration order, then ours */
typedef enum mvf_status {
    MVF_OK = 0,
    MVF_ERR_IO = 1,
    MVF_ERR_INVALID_FORMAT = 2,
    MVF_ERR_UNSUPPORTED_VERSION = 3,
    MVF_ERR_VECTOR_SPACE_NOT_FOUND = 4,
    MVF_ERR_INDEX_OUT_OF_BOUNDS = 5,
    MVF_ERR_DIMENSION_MISMATCH = 6,
    MVF_ERR_INVALID_VECTOR_TYPE = 7,
    MVF_ERR_CORRUPTED_DATA = 8,
    MVF_ERR_EXTENSION = 9,
    MVF_ERR_BUILD = 10,
    MVF_ERR_CUDA = 11,
    MVF_ERR_NCCL = 12
} mvf_status;

/* schema/types.fbs:3-11 */
typedef enum mvf_data_type {
    MVF_FLOAT32 = 0, MVF_FLOAT16 = 1, MVF_INT8 = 2, MVF_UINT8 = 3,
    MVF_UINT32 = 4, MVF_UINT64 = 5, MVF_STRING_REF = 6
} mvf_data_type;
/* schema/types.fbs:14-17 */
typedef enum mvf_vector_type { MVF_DENSE = 0, MVF_SPARSE = 1 } mvf_vector_type;
/* schema/types.fbs:19-25; README.md:160-164 spells InnerProduct "Dot" and L2 "Euclidean" */
typedef enum mvf_distance_metric {
    MVF_L2 = 0, MVF_INNER_PRODUCT = 1, MVF_DOT = 1, MVF_COSINE = 2, MVF_CUSTOM = 255,
    MVF_METRIC_SPACE_DEFAULT = -1
} mvf_distance_metric;

typedef struct mvf_file mvf_file;           /* MvfReader / "MvfFile"   src/reader.rs:27-31   */
typedef struct mvf_space mvf_space;         /* VectorSpace<'a>         src/vectors/vector_space.rs:34-39 */
typedef struct mvf_builder mvf_builder;     /* MvfBuilder              src/builder.rs:93-98  */
typedef struct mvf_gpu mvf_gpu;             /* the search engine on one B200, or on several from one process (new) */
typedef struct mvf_gpu_space mvf_gpu_space; /* HBM-resident shard of one vector space (new) */

MVF_API const char* mvf_last_error(void);
/* The two numbers of the reference's structured error variants for the last failing call on this thread, in
 * declaration order (src/errors.rs:15-31): UnsupportedVersion{got, expected}, IndexOutOfBounds{index, len},
 * DimensionMismatch{expected, actual}, InvalidVectorType{expected, actual} (VectorType values); 0, 0 otherwise.
 * With them a binding rebuilds every MvfError variant exactly. */
MVF_API void mvf_last_error_fields(uint64_t* first, uint64_t* second);
MVF_API const char* mvf_version_string(void);

/* ---------------------------------------------------------------- reader (src/reader.rs) -- */
/* MvfReader::open, src/reader.rs:45-79 (README.md:97 spells it MvfFile::open) */
MVF_API int mvf_file_open(const char* path, mvf_file** out);
MVF_API void mvf_file_close(mvf_file* f);
/* MvfReader::version, src/reader.rs:82 */
MVF_API int mvf_file_version(const mvf_file* f, uint16_t* out);
/* MvfReader::num_vector_spaces, src/reader.rs:87 */
MVF_API int mvf_file_num_vector_spaces(const mvf_file* f, size_t* out);
/* MvfReader::vector_space_names, src/reader.rs:92-98 (i-th name; pointer valid while file open) */
MVF_API int mvf_file_vector_space_name(const mvf_file* f, size_t i, const char** out);
/* MvfReader::file_size, src/reader.rs:122 */
MVF_API int mvf_file_size(const mvf_file* f, uint64_t* out);
/* MvfReader::has_metadata, src/reader.rs:127 */
MVF_API int mvf_file_has_metadata(const mvf_file* f, int* out);
/* MvfReader::metadata_column_names, src/reader.rs:132-143 */
MVF_API int mvf_file_num_metadata_columns(const mvf_file* f, size_t* out);
MVF_API int mvf_file_metadata_column_name(const mvf_file* f, size_t i, const char** out);
/* MvfReader::validate, src/reader.rs:149-162 */
MVF_API int mvf_file_validate(const mvf_file* f);
/* the finished intent of MvfReader::validate_with_checksum, src/reader.rs:172-221 (CRC-32/IEEE
 * of every block with a non-zero checksum, offsets taken relative to the start magic) */
MVF_API int mvf_file_validate_with_checksum(const mvf_file* f);
/* MvfReader::vector_space, src/reader.rs:104-119 */
MVF_API int mvf_file_vector_space(const mvf_file* f, const char* name, mvf_space** out);
MVF_API void mvf_space_release(mvf_space* s);

/* ---------------------------------------- space view (src/vectors/vector_space.rs:62-89) -- */
MVF_API int mvf_space_name(const mvf_space* s, const char** out);
MVF_API int mvf_space_dimension(const mvf_space* s, uint32_t* out);
MVF_API int mvf_space_total_vectors(const mvf_space* s, uint64_t* out);
MVF_API int mvf_space_vector_type(const mvf_space* s, int* out);
MVF_API int mvf_space_distance_metric(const mvf_space* s, int* out);
MVF_API int mvf_space_data_type(const mvf_space* s, int* out);
/* VectorSpace::get_vector + Vector::as_bytes: src/vectors/vector_space.rs:101-142,
 * src/vectors/vector.rs:61-63.  Zero-copy pointer into the mapping. */
MVF_API int mvf_space_get_vector(const mvf_space* s, uint64_t index, const void** ptr, size_t* nbytes);
/* get_vector + Vector::as_f32: src/vectors/vector.rs:71-92 (Int8/UInt8 -> MVF_ERR_BUILD) */
MVF_API int mvf_space_get_vector_f32(const mvf_space* s, uint64_t index, float* out, size_t cap);
/* VectorSpace::map_vector_range: src/vectors/vector_space.rs:155-188 */
MVF_API int mvf_space_map_vector_range(const mvf_space* s, uint64_t start, uint64_t count,
                                       const void** ptr, size_t* nbytes);
/* TombstoneInfo{Bitmap} of the space (schema/core.fbs:34-39, schema/types.fbs:35-39): bit i (LSB first)
 * set = row i deleted.  bitmap = NULL when the footer has none (the reference's builder never writes one,
 * src/builder.rs:486). */
MVF_API int mvf_space_tombstones(const mvf_space* s, const void** bitmap, size_t* nbytes, uint64_t* deleted_count);
/* TombstoneInfo{SortedList} (schema/types.fbs:35-39): ascending deleted row positions, u64 little-endian each (the
 * schema leaves the element type open and nothing in the reference writes one: this is the definition our writer
 * and reader share).  rows = NULL when the space has no list; the pointer is into the mmap and may be unaligned. */
MVF_API int mvf_space_tombstone_list(const mvf_space* s, const uint64_t** rows, size_t* count, uint64_t* deleted_count);
/* External ids: the block vector_ids_block_index names (schema/core.fbs:54; u64 LE per row, as
 * src/builder.rs:257-271 serialises VectorSpaceBuilder.vector_ids).  ids = NULL when the footer says 0 = "use
 * positions as IDs" (every file the reference's builder writes, src/builder.rs:483).  May be unaligned. */
MVF_API int mvf_space_vector_ids(const mvf_space* s, const void** ids, size_t* count);
/* Raw bytes of a metadata column (table MetadataColumn, schema/core.fbs:16-25; written by
 * MvfBuilder::add_metadata_column, src/builder.rs:211-236; the reference reader only lists the names,
 * src/reader.rs:132-143).  MVF_ERR_BUILD if there is no such column. */
MVF_API int mvf_file_metadata_column(const mvf_file* f, const char* name, const void** bytes, size_t* nbytes, int* data_type);
/* Pre-filter helper: evaluates lo <= column[i] <= hi for n values of `data_type`
 * (Float32, Float16, Int8, UInt8, UInt32, UInt64) and SETS bit i of out_bitmap ((n+7)/8 bytes, LSB first) for the
 * rows that FAIL, i.e. it produces exactly what mvf_gpu_space_set_row_mask expects. */
MVF_API int mvf_filter_range_bitmap(const void* column, int data_type, uint64_t n, double lo, double hi, uint8_t* out_bitmap);
/* rows the reference scan `for i in 0..space.total_vectors()` can actually visit:
 * min(total_vectors, block.size / (dimension*elem_size)) -- see DESIGN.md "f16 total_vectors" */
MVF_API int mvf_space_scannable_vectors(const mvf_space* s, uint64_t* out);

/* ------------------------------------------------ writer (src/builder.rs, src/io.rs) ----- */
/* MvfBuilder::new, src/builder.rs:101-107 */
MVF_API int mvf_builder_new(mvf_builder** out);
MVF_API void mvf_builder_free(mvf_builder* b);
/* MvfBuilder::add_vector_space, src/builder.rs:112-135 */
MVF_API int mvf_builder_add_vector_space(mvf_builder* b, const char* name, uint32_t dimension,
                                         int vector_type, int distance_metric, int data_type,
                                         size_t* out_index);
/* MvfBuilder::add_vectors, src/builder.rs:151-196: rows of f32, stored as the space's dtype
 * (Float32 verbatim, Float16 round-to-nearest-even).  Int8/UInt8 -> MVF_ERR_BUILD like :192. */
MVF_API int mvf_builder_add_vectors_f32(mvf_builder* b, const char* space, const float* rows,
                                        size_t nrows, uint32_t dimension);
/* extension the reference lacks: append rows already encoded in the space's dtype (the only way
 * to write Int8/UInt8) */
MVF_API int mvf_builder_add_vectors_raw(mvf_builder* b, const char* space, const void* rows,
                                        size_t nrows, uint32_t dimension);
/* MvfBuilder::add_metadata_column, src/builder.rs:211-236 (opaque bytes) */
MVF_API int mvf_builder_add_metadata_column(mvf_builder* b, const char* name, int data_type,
                                            const void* bytes, size_t nbytes);
/* extension: attach a deletion bitmap to a space (written as TombstoneInfo{Bitmap} + one data block) */
MVF_API int mvf_builder_set_tombstones(mvf_builder* b, const char* space, const void* bitmap, size_t nbytes);
/* same, as TombstoneInfo{SortedList}: strictly ascending row positions (replaces a bitmap set earlier) */
MVF_API int mvf_builder_set_tombstone_list(mvf_builder* b, const char* space, const uint64_t* rows, size_t count);
/* VectorSpaceBuilder.vector_ids (src/builder.rs:61): one external u64 id per row already added to the space; written
 * as its own data block and named by vector_ids_block_index.  MVF_ERR_DIMENSION_MISMATCH if count != rows. */
MVF_API int mvf_builder_set_vector_ids(mvf_builder* b, const char* space, const uint64_t* ids, size_t count);
/* MvfBuilder::build + BuiltMvf::save, src/builder.rs:241-308,404-411, src/io.rs:41-46.
 * Writes the correct total_vectors for every dtype (the reference halves it for Float16,
 * src/builder.rs:476). */
MVF_API int mvf_builder_save(mvf_builder* b, const char* path);

/* ---------------------------------------------------------------- GPU engine (new) ------- */
/* An engine bound to one device.  Fails with MVF_ERR_CUDA when no sm_100 GPU. */
MVF_API int mvf_gpu_init(int device, mvf_gpu** out);
/* ONE process, n devices: the engine behind a single call site of find_top_k_similar (examples/similarity_search.rs:117)
 * that uses every GPU of the box.  Spaces created on it are row-sharded over the devices (mvf_shard_range), a search
 * runs the shards' kernels concurrently, exchanges the local top-k (peer stores into devices[0]'s window for small
 * payloads, else one ncclAllGather per device on the communicator from ncclCommInitAll; see mvf_gpu_exchange_info)
 * and merges on devices[0].  n = 1 gives a plain single-device engine.  mvf_gpu_set_stream,
 * mvf_gpu_comm_init and mvf_gpu_tc_dump are single-device calls and answer MVF_ERR_BUILD on such an engine;
 * mvf_gpu_search_device takes pointers into devices[0]'s memory. */
MVF_API int mvf_gpu_init_multi(const int* devices, int n, mvf_gpu** out);
/* devices this engine spans (1 for mvf_gpu_init) */
MVF_API int mvf_gpu_device_count(const mvf_gpu* g, int* out);
MVF_API void mvf_gpu_shutdown(mvf_gpu* g);
/* run every search of this engine on the caller's CUDA stream (cudaStream_t as void*) */
MVF_API int mvf_gpu_set_stream(mvf_gpu* g, void* cuda_stream);
MVF_API int mvf_gpu_device_info(const mvf_gpu* g, int* sm_count, int* cc_major, int* cc_minor,
                                uint64_t* free_bytes, uint64_t* total_bytes);
/* multi-GPU: one process per GPU; rank 0 makes the id, the launcher broadcasts the 128 bytes */
MVF_API int mvf_gpu_comm_unique_id(void* out_128_bytes);
MVF_API int mvf_gpu_comm_init(mvf_gpu* g, const void* unique_id_128_bytes, int rank, int world);
/* contiguous row shard of rank `rank`: [begin, begin+count) with ceil(n/world) rows per rank */
MVF_API int mvf_shard_range(uint64_t n, int world, int rank, uint64_t* begin, uint64_t* count);

/* residency: upload this rank's shard of the space once (page-locked registration of the mapped
 * block, chunked async H2D, row norms computed on device).  Replaces the per-row
 * get_vector()+as_f32() of examples/similarity_search.rs:147-149. */
MVF_API int mvf_gpu_space_upload(mvf_gpu* g, const mvf_space* s, mvf_gpu_space** out);
/* same from caller memory: `rows` holds local_n rows of the shard starting at global row
 * row_base of a space with global_n rows.  Shards must be contiguous and in rank order (rank r's rows precede rank
 * r+1's): "lowest index on ties" is decided by (rank, local row).  Like every space creation on an engine with a
 * communicator this is a collective call (the ranks exchange their first rows and the facts the kernel choice
 * depends on). */
MVF_API int mvf_gpu_space_from_host(mvf_gpu* g, const void* rows, int data_type, int metric,
                                    uint64_t local_n, uint32_t dimension, uint64_t row_base,
                                    uint64_t global_n, mvf_gpu_space** out);
/* bench/test only: fill this rank's shard of a global_n x dimension space on the device from the
 * counter-based generator (DESIGN.md "synthetic data") */
MVF_API int mvf_gpu_space_synthetic(mvf_gpu* g, int data_type, int metric, uint64_t global_n,
                                    uint32_t dimension, uint64_t seed, mvf_gpu_space** out);
MVF_API void mvf_gpu_space_free(mvf_gpu_space* s);
/* Deletions and pre-filters (TombstoneInfo{Bitmap}, schema/core.fbs:34-39; metadata pre-filters): bit i (LSB
 * first, GLOBAL row index) of `bitmap` set = row i is excluded from every later search of this space; applied as
 * a key mask inside the scan / tensor-core epilogues.  NULL clears.  mvf_gpu_space_upload applies the file's own
 * tombstone bitmap automatically. */
MVF_API int mvf_gpu_space_set_row_mask(mvf_gpu_space* s, const uint8_t* bitmap, uint64_t nbits);
/* The space's own deletions, kept apart from the filter above: every search excludes (tombstones OR row mask).
 * mvf_gpu_space_upload sets this from the file's TombstoneInfo (bitmap or sorted list).  NULL clears. */
MVF_API int mvf_gpu_space_set_tombstones(mvf_gpu_space* s, const uint8_t* bitmap, uint64_t nbits);
/* External ids (vector_ids_block_index): a table of global_n u64 ids, one per GLOBAL row, kept in HBM on every
 * rank; searches then report ids[row] instead of the row position (ties are still broken by position, and
 * mvf_gpu_space_get_vectors still takes positions).  NULL clears.  mvf_gpu_space_upload applies the file's table. */
MVF_API int mvf_gpu_space_set_ids(mvf_gpu_space* s, const uint64_t* ids, uint64_t count);
MVF_API int mvf_gpu_space_info(const mvf_gpu_space* s, uint64_t* local_n, uint64_t* row_base,
                               uint64_t* global_n, uint32_t* dimension, int* data_type, int* metric);
/* What the scan-vs-tensor choice depends on besides the shape (identical on every rank of a sharded space):
 * rows holding Inf/NaN components over all shards (float spaces with any stay on the scan kernel, whose
 * element-wise arithmetic defines the result), whether the split-tf32 operand planes of a Float32 space are
 * resident on every shard, and the row count the choice is taken on (the smallest non-empty shard). */
MVF_API int mvf_gpu_space_path_facts(const mvf_gpu_space* s, uint64_t* nonfinite_rows, int* tf32_planes_resident,
                                     uint64_t* decision_rows);
/* how the rows got to HBM: path 1 = cudaHostRegister'd mapping, 2 = pinned staging, 0 = generated */
MVF_API int mvf_gpu_space_upload_stats(const mvf_gpu_space* s, int* path, double* seconds, uint64_t* bytes);
/* diagnostic: which in-place registration variants were tried ("ro" = cudaHostRegisterReadOnly on the reader's
 * mapping, "rw" = cudaHostRegisterDefault on it, "private" = a copy-on-write re-mapping of the file range), which
 * the driver refused and with what error, and which one was accepted (empty for generated spaces) */
MVF_API const char* mvf_gpu_space_upload_note(const mvf_gpu_space* s);
/* copy rows / norms back (tests: device bytes == mapped bytes) */
MVF_API int mvf_gpu_space_read_rows(const mvf_gpu_space* s, uint64_t local_start, uint64_t count, void* out);
MVF_API int mvf_gpu_space_read_norms(const mvf_gpu_space* s, float* out_norm, float* out_sq);

/* Integrity (the finished intent of MvfReader::validate_with_checksum, src/reader.rs:172-221, on the
 * device): CRC-32/IEEE of the resident rows computed by a CUDA kernel; _verify compares it with the
 * block manifest entry the builder wrote (src/builder.rs:251) -> MVF_ERR_CORRUPTED_DATA on mismatch. */
/* Sharded spaces: every shard hashes its rows, shard CRCs are combined with the crc32_combine identity (under a
 * communicator the call is a collective: every rank calls it and receives the CRC of the whole block). */
MVF_API int mvf_gpu_space_checksum(mvf_gpu_space* s, uint32_t* out_crc, double* out_kernel_ms);
MVF_API int mvf_gpu_space_verify_checksum(mvf_gpu_space* s, const mvf_space* file_space);

/* The rows behind a result list, from the HBM copy: the `vector` field of ScoredVector
 * (examples/similarity_search.rs:18) / VectorSpace::get_vectors_batch (src/vectors/vector_space.rs:230-243).
 * ids are global row ids; out receives count rows, raw (as_f32 = 0) or widened like Vector::as_f32
 * (as_f32 = 1; Int8/UInt8 -> MVF_ERR_BUILD like src/vectors/vector.rs:90).  Ids outside this rank's
 * shard (and the UINT64_MAX padding of short result lists) give zero rows. */
MVF_API int mvf_gpu_space_get_vectors(mvf_gpu_space* s, const uint64_t* ids, uint32_t count, int as_f32, void* out);

/* find_top_k_similar(space, query, k), examples/similarity_search.rs:140-176, for nq queries at
 * once.  queries: nq x dimension, HOST memory, dtype query_dtype (MVF_FLOAT32, or the space's own
 * dtype; integer spaces take only their own dtype).  metric: MVF_METRIC_SPACE_DEFAULT or an
 * override.  Outputs (HOST, caller-owned, nq*k each, best first; unused tail slots get id
 * UINT64_MAX): out_ids = ScoredVector.index; out_scores = ScoredVector.score (L2 distance for
 * float spaces, squared distance for integer spaces, dot for InnerProduct, 1-cos for Cosine);
 * out_iscores (may be NULL) = exact int32 score for integer spaces with L2/InnerProduct.
 * With a communicator every rank must call it with the same queries; every rank gets the result.
 * Limits of this build (each answered with MVF_ERR_BUILD, never with a wrong result): a resident SHARD holds
 * fewer than 2^32 rows (candidates carry 32-bit local row ids; the space itself, row_base and the ids returned are
 * 64-bit, so 8 GPUs take 8 x 2^32 rows); a row is at most 1 MiB; integer spaces have dimension <= 33025 (int32 accumulation,
 * SURVEY.md 8a); the per-CTA candidate lists of the scan kernel live in shared memory, so k is bounded by
 * 8 * pow2(8k) + dimension * 4 (or * 1 for integer spaces) <= ~227 KB, i.e. k <= 2048 for rows up to 24 KB;
 * the tensor-core path is taken for k <= 512 only (larger k run on the scan kernel).
 * Float spaces holding Inf/NaN rows, and batches holding Inf/NaN queries, always run on the scan kernel. */
MVF_API int mvf_gpu_search(mvf_gpu_space* s, const void* queries, int query_dtype, uint32_t nq,
                           uint32_t k, int metric, uint64_t* out_ids, float* out_scores,
                           int32_t* out_iscores);
/* same with DEVICE pointers, asynchronous on the engine stream.  One exception: a float batch that qualifies for
 * the tensor path is first checked for Inf/NaN on the device, which costs one stream synchronisation (~10 us). */
MVF_API int mvf_gpu_search_device(mvf_gpu_space* s, const void* d_queries, int query_dtype,
                                  uint32_t nq, uint32_t k, int metric, uint64_t* d_out_ids,
                                  float* d_out_scores, int32_t* d_out_iscores);
/* test hook for the tensor-core path: runs K2 and copies back the raw ranking values it selected
 * on (padded to 128 queries x 256 rows tiles); out holds out_elems floats */
MVF_API int mvf_gpu_tc_dump(mvf_gpu_space* s, const void* queries, int query_dtype, uint32_t nq, uint32_t k,
                            int metric, float* out, uint64_t out_elems, uint32_t* out_rows, uint32_t* out_cols);
/* Diagnostic, host arithmetic only (no GPU needed): the launch shape the tensor-core kernel would use for nq
 * queries, top-k, n resident rows of `dimension` elements (Float16, Int8 or UInt8) on a GPU with sm_count SMs.
 * out16 = { query tiles, database tiles, runs ("chunks"), tiles per run, cluster size, query-tile groups, work items,
 * grid, K blocks, candidate buffer entries, compaction trigger, pacing window, epilogue warpgroups, ring stages,
 * pair form (0/1), dynamic shared memory bytes }.  Lets the CPU test-suite check the planner's invariants. */
MVF_API int mvf_tc_plan(int data_type, uint32_t nq, uint32_t k, uint64_t n, uint32_t dimension, int sm_count, uint32_t* out16);
/* Host arithmetic only (no GPU needed): the row count every rank of a space sharded by mvf_shard_range(global_n, world)
 * takes the scan-vs-tensor decision on (the smallest non-empty shard: identical on every rank by construction, so the
 * ranks never exchange candidates in two key encodings) and whether a batch of nq queries of the space's dtype would
 * go to the tensor path in auto mode.  Lets the CPU test-suite pin the rule. */
MVF_API int mvf_path_decision(uint64_t global_n, int world, int data_type, int metric, uint32_t nq, uint32_t k,
                              uint32_t dimension, uint64_t* decision_rows, int* tensor_path);
/* path selection: 0 auto, 1 bandwidth scan (K1), 2 tensor-core contraction (K2) */
MVF_API int mvf_gpu_set_path(mvf_gpu* g, int path);
/* instrumentation: when enabled, searches bracket the dominant kernel with CUDA events on the
 * engine stream; stats are cumulative since the last reset */
MVF_API int mvf_gpu_profile(mvf_gpu* g, int enable);
MVF_API int mvf_gpu_profile_read(mvf_gpu* g, double* scan_ms_total, uint64_t* scan_launches,
                                 uint64_t* all_launches, int* last_path);
MVF_API int mvf_gpu_synchronize(mvf_gpu* g);
/* The one exchange step of a sharded search (every rank contributes its nq*k best, merged per query; replaces nothing
 * in the reference, which is single-threaded: examples/similarity_search.rs:147) runs in one of two ways, the same on
 * every rank: payloads of up to 512 candidates (every batch-1 search) are pushed into every rank's window with peer
 * stores over NVLink and merged by the same kernel (windows shared by cudaIpc* between processes, by peer access inside
 * a one-process engine); larger ones use ncclAllGather + the merge kernel.  *peer_ready = 1 when the peer path is set up
 * (decided collectively at mvf_gpu_comm_init / mvf_gpu_init_multi; MVF_PEER_EXCHANGE=0 turns it off), the counters say
 * which route the searches of this engine took, mvf_gpu_exchange_note why the peer path is off when it is. */
MVF_API int mvf_gpu_exchange_info(const mvf_gpu* g, int* peer_ready, uint64_t* peer_searches, uint64_t* nccl_searches);
MVF_API const char* mvf_gpu_exchange_note(const mvf_gpu* g);

/* host helpers shared by callers that move candidates themselves (and by the tests) */
MVF_API uint32_t mvf_ordered_key_f32(float v);
MVF_API float mvf_ordered_key_to_f32(uint32_t key);
/* product-side synthetic generator, host flavour (same definition as the device fill) */
MVF_API int mvf_synth_fill_host(void* out, int data_type, uint64_t seed, uint64_t row0,
                                uint64_t nrows, uint32_t dimension);

#ifdef __cplusplus
}
#endif
#endif /* MVF_B200_H */
