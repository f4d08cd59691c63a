// tc_path.hpp -- host side of K2, the batched (tensor-core) path: state, planning arithmetic, entry points.
// No device code here.  engine.cu includes this; tc_path.cu defines run()/prepare_split() and launches the
// kernels of tc_kernels.cuh, which tc_inst.cu instantiates (one object per operand kind x epilogue-group
// count, so the build parallelises).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mvf_host.hpp"

namespace mvf {
namespace tc {

typedef unsigned long long u64;

// tile shape (tc_kernels.cuh static_asserts the same numbers)
constexpr int BM = 128;            // queries per tile (TMEM lanes)
constexpr int BN = 256;            // database rows per tile (TMEM columns, MMA N)
constexpr int BKB = 128;           // K block in bytes (one 128-byte swizzle row)
constexpr int MAX_STAGES = 6;
constexpr int MAX_CAP = 1024;      // candidate buffer entries per query (power of two)
constexpr int MAX_K = 512;         // largest k the fused epilogue keeps (a buffer holds k + two tiles of appends)
constexpr int A_BYTES = BM * BKB;  // 16 KiB
constexpr int B_BYTES = BN * BKB;  // 32 KiB
// operand kinds.  KIND_TF32 is K2' (Float32 spaces): every operand is carried as two tf32-exact planes
// (hi = tf32(x), lo = tf32(x - hi)) and a product is three MMAs, q_hi.x_hi + q_lo.x_hi + q_hi.x_lo
enum { KIND_F16 = 0, KIND_I8 = 1, KIND_TF32 = 2 };
constexpr int planes_of(int kind) { return kind == KIND_TF32 ? 2 : 1; }
// bytes of one ring stage: the query tile and (the CTA's share of) the database tile, per plane
constexpr int stage_bytes_of(int kind, bool pair) { return planes_of(kind) * (A_BYTES + (pair ? B_BYTES / 2 : B_BYTES)); }
constexpr int threads_for(int eg) { return 64 + 128 * eg; }

#define TC_CUDA_CHECK(expr)                                                                                   \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            throw MvfError::cuda(std::string(cudaGetErrorName(_e)) + ": " + cudaGetErrorString(_e) + " at " #expr); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Buf {
    void* p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        TC_CUDA_CHECK(cudaMalloc(&p, n));
        cap = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct TcState {
    EncodeTiledFn encode = nullptr;
    Buf qprep, qconst, cand, lists, dump, gthr, progress, stats;
    int l2_prefetch = 0;           // (measured slower: the kernel is power-capped, extra L2 traffic costs clocks) database tiles prefetched into L2 ahead of the TMA loads (MVF_TC_PREFETCH=0..4)
    int epi_groups = 0;            // 0 = by row length (4 epilogue warpgroups for K <= 256 bytes, else 2); MVF_TC_EPI_GROUPS=2|4
    int cluster = 2;               // CTAs per cluster sharing database tiles by TMA multicast (MVF_TC_CLUSTER=1|2)
    int sync_window = -1;          // pacing of work items that share database tiles: -1 = auto (about 16 MB of tiles in flight, whenever two or more query-tile groups share rows), 0 = off, else tiles (MVF_TC_SYNC_WINDOW)
    int items_per_sm = 8;          // work items per SM the database is cut into (MVF_TC_ITEMS_PER_SM)
    int pair = 1;                  // clusters of two run cta_group::2 MMAs (256 queries x 256 rows, half the database tile per CTA) instead of multicasting whole tiles (MVF_TC_PAIR=0|1)
    // auto path: batches at least this large go to the tensor cores.  Measured crossover (scripts/crossover.py):
    // f16 768-d rows: the scan kernel pays a convert + FMA per element per query and loses from 4 queries on;
    // uint8 128-d rows: dp4a keeps the scan ahead up to ~7 queries.  f32 rows: the scan kernel takes 8 queries per
    // pass over HBM, the split-tf32 contraction one pass for the whole batch at a third of the tf32 rate: from 16 on.
    // MVF_TC_MIN_BATCH overrides all three.
    int min_batch_f16 = 4, min_batch_i8 = 8, min_batch_f32 = 16;
    bool env_read = false;
};

struct SearchArgs {
    const unsigned char* rows; const float* norm; const float* sq; const uint32_t* mask;
    const int* acc_seed;               // integer spaces, L2: per-row accumulator seeds padded to whole tiles (engine.cu build_acc_seed)
    const unsigned char* rows_split;   // KIND_TF32 only: [hi plane | lo plane], each n rows of `pitch` bytes (prepare_split)
    uint64_t n; uint32_t d, pitch; int dtype, metric;
    const void* queries; uint32_t nq, k;
    bool dump;                     // tests: also write the raw ranking values
};
struct SearchResult {
    const u64* lists; uint32_t chunks; const float* qconst; const float* dump; uint32_t dump_cols; uint64_t launches;
};

inline int kind_of(int dtype) { return dtype == 1 ? KIND_F16 : (dtype == 0 ? KIND_TF32 : KIND_I8); }
inline uint32_t elem_bytes(int kind) { return kind == KIND_F16 ? 2u : (kind == KIND_TF32 ? 4u : 1u); }

// tuning knobs from the environment (host only; read when the engine is created so that path selection sees them)
inline void read_env(TcState& st) {
    if (st.env_read) return;
    st.env_read = true;
    if (const char* c = std::getenv("MVF_TC_MIN_BATCH")) st.min_batch_f16 = st.min_batch_i8 = st.min_batch_f32 = std::max(1, std::atoi(c));
    if (const char* c = std::getenv("MVF_TC_CLUSTER")) st.cluster = std::atoi(c) == 1 ? 1 : 2;
    if (const char* c = std::getenv("MVF_TC_ITEMS_PER_SM")) st.items_per_sm = std::max(1, std::min(64, std::atoi(c)));
    if (const char* c = std::getenv("MVF_TC_PAIR")) st.pair = std::max(0, std::min(2, std::atoi(c)));
    if (const char* c = std::getenv("MVF_TC_SYNC_WINDOW")) st.sync_window = std::max(-1, std::atoi(c));
    if (const char* c = std::getenv("MVF_TC_EPI_GROUPS")) st.epi_groups = std::atoi(c) == 4 ? 4 : (std::atoi(c) == 2 ? 2 : 0);
    if (const char* c = std::getenv("MVF_TC_PREFETCH")) st.l2_prefetch = std::max(0, std::min(4, std::atoi(c)));
}

// shared memory: operand ring + per-warp sort scratch + per-warp column parameters + barriers
inline size_t smem_bytes(uint32_t cap, int eg, int stages, bool pair, int kind = KIND_F16) {
    const size_t warps = 4 * (size_t)eg, cols = BN / eg;
    // ... three column-parameter slots per epilogue warp (tc_kernels.cuh PSLOTS), barriers, live-column words
    return 1024 + (size_t)stages * (size_t)stage_bytes_of(kind, pair) + warps * (size_t)cap * 8 + warps * 3 * cols * 4 +
           (2 * MAX_STAGES + 4) * 8 + 16 + warps * 3 * (cols / 32) * 4 + 16;
}
// deepest operand ring that fits beside the sort scratch (large k / many epilogue warps cost stages):
// 3-4 stages of 48 KiB, or 5-6 of 32 KiB in the pair form; the two-plane tf32 stages are twice as large
inline int stages_for(uint32_t cap, int eg, bool pair, int kind = KIND_F16) {
    int s = pair ? MAX_STAGES : 4;
    while (s > 2 && smem_bytes(cap, eg, s, pair, kind) > 227 * 1024) s--;
    return s;
}
// Candidate buffers.  A buffer is compacted (bitonic sort of the next power of two >= its fill, keep k) between
// tiles once it holds more than `trigger` entries.  Sort cost per append ~ P log^2 P / (P - k) favours SMALL sorts
// for small k (k = 10: P = 64) and P ~ 2.5k for large k, so the trigger sits just under a power of two >= 2k + 16;
// the buffer itself (cap, power of two, <= MAX_CAP) also has to take one more tile of appends past the trigger
// (BN / eg columns per thread).
inline uint32_t trigger_for_k(uint32_t k) {
    uint32_t p2 = 32;
    while (p2 < 2 * k + 16) p2 <<= 1;
    return p2 - 16;
}
inline uint32_t cap_for_k(uint32_t k, int eg) {
    uint32_t want = trigger_for_k(k) + (uint32_t)(BN / eg);
    // k > 128: a 2.5k-entry sort scratch per epilogue warp would cost the operand ring its stages; k entries plus two
    // tiles of appends keep the scratch at 32-64 KiB per CTA (sorts of <= 1024 entries, >= one tile admitted per sort)
    if (k > 128) want = k + 2u * (uint32_t)(BN / eg);
    const uint32_t lim = k > 128 ? (uint32_t)MAX_CAP : 512u;
    uint32_t cap = 128;
    while (cap < want && cap < lim) cap <<= 1;
    return cap;
}

// K2 takes f16 spaces with f16 queries, f32 spaces with f32 queries (split tf32) and int8/uint8 spaces (every metric:
// integer cosine, whose exact definition needs a division per element, forms it for the candidates only), k <= MAX_K (512).  `n` must be a value every
// rank of a sharded space agrees on (engine.cu passes the smallest non-empty shard): ranks on different paths would
// exchange keys in different encodings.
struct TcState;
inline bool plan_fits(const TcState& st, int dtype, uint32_t nq, uint32_t k, uint64_t n, uint32_t d);   // below make_plan
inline bool eligible(const TcState& st, int path, int dtype, int qdtype, int metric, uint32_t nq, uint32_t k, uint64_t n,
                     uint32_t d, uint32_t pitch) {
    (void)pitch;
    if (path == 1) return false;
    const bool f16 = dtype == 1 && qdtype == 1;
    const bool f32 = dtype == 0 && qdtype == 0;
    const bool i8 = (dtype == 2 || dtype == 3) && qdtype == dtype;
    (void)metric;
    if (!(f16 || f32 || i8)) return false;
    if (k > (uint32_t)MAX_K || n == 0 || d == 0) return false;
    // split tf32: the tensor core's f32 accumulation over 3 * d / 8 MMA steps leaves ~4e-6 |q||x| at d = 768 (measured,
    // tests/test_gpu_tc.py); the stated tolerance is 1e-5 |q||x|, so longer rows stay on the scan kernel
    if (f32 && d > 1024) return false;
    // the single-CTA split-tf32 form (one query tile) has 96 KiB stages: two of them do not fit beside the 512-entry sort scratch
    if (f32 && (nq <= (uint32_t)BM || st.cluster == 1) && smem_bytes(cap_for_k(k, 2), 2, 2, false, KIND_TF32) > 227 * 1024) return false;
    // (large k: the sort scratch must leave room for the operand ring in the form make_plan picks)
    if (k > 128 && !plan_fits(st, dtype, nq, k, n, d)) return false;
    if (path == 2) return true;
    return nq >= (uint32_t)(f16 ? st.min_batch_f16 : (f32 ? st.min_batch_f32 : st.min_batch_i8)) && n >= 4096;
}

// Everything the launch is shaped by, as plain host arithmetic (no CUDA calls), so the planner can be checked on a
// machine without a GPU (mvf_tc_plan in the C ABI, tests/test_tc_plan.py).
struct Plan {
    uint32_t mtiles, ntiles, chunks, tiles_per_chunk, CL, mgroups, items, grid, kblocks, cap, trigger, sync_window, qpitch;
    int EG, stages, kind;
    bool pair;
    size_t smem;
};
inline Plan make_plan(const TcState& st, int dtype, uint32_t nq, uint32_t kk, uint64_t n, uint32_t d, uint32_t pitch, int sm_count) {
    Plan pl;
    pl.kind = kind_of(dtype);
    const uint32_t es = elem_bytes(pl.kind);
    pl.qpitch = (d * es + 15u) & ~15u;
    pl.mtiles = (nq + BM - 1) / BM;
    pl.ntiles = (uint32_t)((n + BN - 1) / BN);
    // work items = (query tile, contiguous run of database tiles); ~8 items per SM, each item keeps
    // its own running thresholds, the item lists are merged afterwards
    uint32_t chunks = std::max<uint32_t>(1, ((uint32_t)st.items_per_sm * (uint32_t)sm_count + pl.mtiles - 1) / pl.mtiles);
    chunks = std::max<uint32_t>(1, std::min(chunks, pl.ntiles));
    pl.tiles_per_chunk = (pl.ntiles + chunks - 1) / chunks;
    pl.chunks = (pl.ntiles + pl.tiles_per_chunk - 1) / pl.tiles_per_chunk;
    pl.CL = (st.cluster == 2 && pl.mtiles >= 2) ? 2u : 1u;
    pl.mgroups = (pl.mtiles + pl.CL - 1) / pl.CL;
    pl.items = pl.mgroups * pl.chunks;
    pl.grid = std::min<uint32_t>((uint32_t)sm_count / pl.CL, pl.items) * pl.CL;
    pl.kblocks = (d * es + BKB - 1) / BKB;
    pl.EG = st.epi_groups ? st.epi_groups : (pl.kblocks <= 2 ? 4 : 2);
    if (pl.kind == KIND_TF32) pl.EG = 2;          // three MMAs per K step: the tensor pipe, not the epilogue, is the bound
    pl.cap = cap_for_k(kk, pl.EG);
    // short rows are bound by the epilogue, and a pair's leader waits for the slower of two epilogues (measured on
    // config E batched: 840 -> 620 TOP/s), so the pair form is for rows of more than two K blocks (MVF_TC_PAIR=2 forces it)
    pl.pair = pl.CL == 2 && (st.pair == 2 || (st.pair == 1 && pl.kblocks > 2));
    if (pl.kind == KIND_TF32) pl.pair = pl.CL == 2;   // the two-plane stages only come as pair (64 KiB) or single-CTA (96 KiB)
    if (pl.EG == 4 && smem_bytes(pl.cap, pl.EG, 3, pl.pair, pl.kind) > 227 * 1024) { pl.EG = 2; pl.cap = cap_for_k(kk, pl.EG); }
    pl.stages = stages_for(pl.cap, pl.EG, pl.pair, pl.kind);
    pl.trigger = std::min<uint32_t>(trigger_for_k(kk), pl.cap - (uint32_t)(BN / pl.EG));
    pl.smem = smem_bytes(pl.cap, pl.EG, pl.stages, pl.pair, pl.kind);
    // pacing window (tiles a producer may lead the slowest item on the same rows): sized so that the tiles "in
    // flight" across all concurrently processed runs stay around 16 MB of L2 -- 9 tiles of 393 KB for config C
    // (4.6 runs at a time), 2 for D batch 1024 (18.5 runs; measured 1 105 vs 1 075 TFLOP/s at 8), dozens for 32 KB tiles
    pl.sync_window = 0;
    if (st.sync_window >= 0) pl.sync_window = (uint32_t)st.sync_window;
    else if (pl.mgroups >= 2) {
        const double runs = std::max(1.0, (double)(pl.grid / pl.CL) / (double)pl.mgroups);
        const double tile_bytes = (double)BN * (double)pitch * planes_of(pl.kind);
        pl.sync_window = (uint32_t)std::min(64.0, std::max(2.0, 16.0 * 1048576.0 / (runs * tile_bytes)));
    }
    return pl;
}

inline bool plan_fits(const TcState& st, int dtype, uint32_t nq, uint32_t k, uint64_t n, uint32_t d) {
    const uint32_t es = elem_bytes(kind_of(dtype));
    const Plan pl = make_plan(st, dtype, nq, k, n, d, (d * es + 15u) & ~15u, 148);
    return pl.smem <= 227 * 1024 && pl.trigger >= k && pl.trigger + (uint32_t)(BN / pl.EG) <= pl.cap;
}

// defined in tc_path.cu
void init(TcState& st);                                     // resolves cuTensorMapEncodeTiled (throws MvfError::cuda)
SearchResult run(TcState& st, const SearchArgs& a, cudaStream_t stream, int sm_count);
// K2' residency: writes the two tf32-exact planes of an f32 shard ([hi | lo], n rows of `pitch` bytes each) into `out`
void split_rows_tf32(const unsigned char* rows, uint64_t n, uint32_t pitch, unsigned char* out, cudaStream_t stream);
void release(TcState& st);

// defined in tc_inst.cu (compiled once per kind x epilogue-group count): kernel_k<KIND>_eg<EG>
const void* kernel_k0_eg2(int mode, int cl, bool pair);
const void* kernel_k0_eg4(int mode, int cl, bool pair);
const void* kernel_k1_eg2(int mode, int cl, bool pair);
const void* kernel_k1_eg4(int mode, int cl, bool pair);
const void* kernel_k2_eg2(int mode, int cl, bool pair);
// mode: 0 L2, 1 dot / float cosine, 3 integer cosine (tc_kernels.cuh)
inline const void* select_kernel(int kind, int eg, int mode, int cl, bool pair) {
    if (kind == KIND_TF32) return kernel_k2_eg2(mode, cl, pair);
    if (kind == KIND_I8) return eg == 4 ? kernel_k1_eg4(mode, cl, pair) : kernel_k1_eg2(mode, cl, pair);
    return eg == 4 ? kernel_k0_eg4(mode, cl, pair) : kernel_k0_eg2(mode, cl, pair);
}

}  // namespace tc
}  // namespace mvf
