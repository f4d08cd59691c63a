// tc_path.cuh -- host side of K2 (tc_kernels.cuh): TMA tensor maps, work partitioning, launch.
// Included by engine.cu only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mvf_host.hpp"
#include "tc_kernels.cuh"

namespace mvf {
namespace tc {

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
    Buf qprep, qconst, cand, lists, dump, gthr, progress;
    int l2_prefetch = 0;           // (measured slower: the kernel is power-capped, extra L2 traffic costs clocks) database tiles prefetched into L2 ahead of the TMA loads (MVF_TC_PREFETCH=0..4)
    int epi_groups = 0;            // 0 = by row length (4 epilogue warpgroups for K <= 256 bytes, else 2); MVF_TC_EPI_GROUPS=2|4
    int cluster = 2;               // CTAs per cluster sharing database tiles by TMA multicast (MVF_TC_CLUSTER=1|2)
    int sync_window = -1;          // pacing of work items that share database tiles: -1 = auto (about 16 MB of tiles in flight, whenever two or more query-tile groups share rows), 0 = off, else tiles (MVF_TC_SYNC_WINDOW)
    int items_per_sm = 8;          // work items per SM the database is cut into (MVF_TC_ITEMS_PER_SM)
    int pair = 1;                  // clusters of two run cta_group::2 MMAs (256 queries x 256 rows, half the database tile per CTA) instead of multicasting whole tiles (MVF_TC_PAIR=0|1)
    // auto path: batches at least this large go to the tensor cores.  Measured crossover (scripts/crossover.py):
    // f16 768-d rows: the scan kernel pays a convert + FMA per element per query and loses from 4 queries on;
    // uint8 128-d rows: dp4a keeps the scan ahead up to ~7 queries.  MVF_TC_MIN_BATCH overrides both.
    int min_batch_f16 = 4, min_batch_i8 = 8;
    bool attr_set = false;
    bool env_read = false;
};

struct SearchArgs {
    const unsigned char* rows; const float* norm; const float* sq; const uint32_t* mask;
    uint64_t n; uint32_t d, pitch; int dtype, metric;
    const void* queries; uint32_t nq, k;
    bool dump;                     // tests: also write the raw ranking values
};
struct SearchResult {
    const u64* lists; uint32_t chunks; const float* qconst; const float* dump; uint32_t dump_cols; uint64_t launches;
};

// tuning knobs from the environment (host only; read when the engine is created so that path selection sees them)
inline void read_env(TcState& st) {
    if (st.env_read) return;
    st.env_read = true;
    if (const char* c = std::getenv("MVF_TC_MIN_BATCH")) st.min_batch_f16 = st.min_batch_i8 = std::max(1, std::atoi(c));
    if (const char* c = std::getenv("MVF_TC_CLUSTER")) st.cluster = std::atoi(c) == 1 ? 1 : 2;
    if (const char* c = std::getenv("MVF_TC_ITEMS_PER_SM")) st.items_per_sm = std::max(1, std::min(64, std::atoi(c)));
    if (const char* c = std::getenv("MVF_TC_PAIR")) st.pair = std::max(0, std::min(2, std::atoi(c)));
    if (const char* c = std::getenv("MVF_TC_SYNC_WINDOW")) st.sync_window = std::max(-1, std::atoi(c));
    if (const char* c = std::getenv("MVF_TC_EPI_GROUPS")) st.epi_groups = std::atoi(c) == 4 ? 4 : (std::atoi(c) == 2 ? 2 : 0);
    if (const char* c = std::getenv("MVF_TC_PREFETCH")) st.l2_prefetch = std::max(0, std::min(4, std::atoi(c)));
}
inline void init(TcState& st) {
    read_env(st);
    if (st.encode) return;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) { cudaGetLastError(); throw MvfError::cuda("cuTensorMapEncodeTiled is not available from the driver"); }
    st.encode = (EncodeTiledFn)fn;
}

// shared memory: operand ring + per-warp sort scratch + per-warp column parameters + barriers
inline size_t smem_bytes(uint32_t cap, int eg, int stages, bool pair) {
    const size_t warps = 4 * (size_t)eg, cols = BN / eg;
    return 1024 + (size_t)stages * (pair ? PAIR_STAGE_BYTES : STAGE_BYTES) + warps * (size_t)cap * 8 + warps * 2 * cols * 4 +
           (2 * MAX_STAGES + 4) * 8 + 16 + warps * 2 * (cols / 32) * 4 + 16;
}
// deepest operand ring that fits beside the sort scratch (large k / many epilogue warps cost stages):
// 3-4 stages of 48 KiB, or 5-6 of 32 KiB in the pair form
inline int stages_for(uint32_t cap, int eg, bool pair) {
    int s = pair ? MAX_STAGES : 4;
    while (s > 2 && smem_bytes(cap, eg, s, pair) > 227 * 1024) s--;
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
    const uint32_t want = trigger_for_k(k) + (uint32_t)(BN / eg);
    uint32_t cap = 128;
    while (cap < want && cap < (uint32_t)MAX_CAP) cap <<= 1;
    return cap;
}

// K2 takes f16 spaces with f16 queries and int8/uint8 spaces (any metric except integer cosine,
// whose exact definition needs a division per element), k <= 128, rows of at least one K block.
inline bool eligible(const TcState& st, int path, int dtype, int qdtype, int metric, uint32_t nq, uint32_t k, uint64_t n,
                     uint32_t d, uint32_t pitch) {
    (void)pitch;
    if (path == 1) return false;
    const bool f16 = dtype == k::DT_F16 && qdtype == k::DT_F16;
    const bool i8 = (dtype == k::DT_I8 || dtype == k::DT_U8) && qdtype == dtype && metric != k::M_COS;
    if (!(f16 || i8)) return false;
    if (k > 128 || n == 0 || d == 0) return false;
    if (path == 2) return true;
    return nq >= (uint32_t)(f16 ? st.min_batch_f16 : st.min_batch_i8) && n >= 4096;
}

template <bool INT, int MODE, int EG>
inline const void* select_kernel_sc(int cl, bool pair) {
    if (pair) return (const void*)gemm_topk_tc<INT, MODE, 2, EG, true>;
    return cl == 2 ? (const void*)gemm_topk_tc<INT, MODE, 2, EG, false> : (const void*)gemm_topk_tc<INT, MODE, 1, EG, false>;
}
template <int EG>
inline const void* select_kernel_l(bool is_int, bool l2, int cl, bool pair) {
    if (is_int) return l2 ? select_kernel_sc<true, 0, EG>(cl, pair) : select_kernel_sc<true, 1, EG>(cl, pair);
    return l2 ? select_kernel_sc<false, 0, EG>(cl, pair) : select_kernel_sc<false, 1, EG>(cl, pair);
}
inline const void* select_kernel(bool is_int, bool l2, int cl, int eg, bool pair) {
    return eg == 4 ? select_kernel_l<4>(is_int, l2, cl, pair) : select_kernel_l<2>(is_int, l2, cl, pair);
}

inline void make_map(TcState& st, CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint32_t d, uint64_t pitch,
                     uint32_t box_rows) {
    const bool i8 = dtype != k::DT_F16;
    cuuint64_t gdim[2] = {d, rows};
    cuuint64_t gstride[1] = {pitch};
    cuuint32_t box[2] = {i8 ? 128u : 64u, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = st.encode(map, i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                           gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw MvfError::cuda("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

// Everything the launch is shaped by, as plain host arithmetic (no CUDA calls), so the planner can be checked on a
// machine without a GPU (mvf_tc_plan in the C ABI, tests/test_tc_plan.py).
struct Plan {
    uint32_t mtiles, ntiles, chunks, tiles_per_chunk, CL, mgroups, items, grid, kblocks, cap, trigger, sync_window, qpitch;
    int EG, stages;
    bool pair, is_int;
    size_t smem;
};
inline Plan make_plan(const TcState& st, int dtype, uint32_t nq, uint32_t kk, uint64_t n, uint32_t d, uint32_t pitch, int sm_count) {
    Plan pl;
    pl.is_int = dtype != k::DT_F16;
    const uint32_t es = pl.is_int ? 1 : 2;
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
    pl.cap = cap_for_k(kk, pl.EG);
    // short rows are bound by the epilogue, and a pair's leader waits for the slower of two epilogues (measured on
    // config E batched: 840 -> 620 TOP/s), so the pair form is for rows of more than two K blocks (MVF_TC_PAIR=2 forces it)
    pl.pair = pl.CL == 2 && (st.pair == 2 || (st.pair == 1 && pl.kblocks > 2));
    if (pl.EG == 4 && smem_bytes(pl.cap, pl.EG, 3, pl.pair) > 227 * 1024) { pl.EG = 2; pl.cap = cap_for_k(kk, pl.EG); }
    pl.stages = stages_for(pl.cap, pl.EG, pl.pair);
    pl.trigger = std::min<uint32_t>(trigger_for_k(kk), pl.cap - (uint32_t)(BN / pl.EG));
    pl.smem = smem_bytes(pl.cap, pl.EG, pl.stages, pl.pair);
    // pacing window (tiles a producer may lead the slowest item on the same rows): sized so that the tiles "in
    // flight" across all concurrently processed runs stay around 16 MB of L2 -- 9 tiles of 393 KB for config C
    // (4.6 runs at a time), 2 for D batch 1024 (18.5 runs; measured 1 105 vs 1 075 TFLOP/s at 8), dozens for 32 KB tiles
    pl.sync_window = 0;
    if (st.sync_window >= 0) pl.sync_window = (uint32_t)st.sync_window;
    else if (pl.mgroups >= 2) {
        const double runs = std::max(1.0, (double)(pl.grid / pl.CL) / (double)pl.mgroups);
        const double tile_bytes = (double)BN * (double)pitch;
        pl.sync_window = (uint32_t)std::min(64.0, std::max(2.0, 16.0 * 1048576.0 / (runs * tile_bytes)));
    }
    return pl;
}

inline SearchResult run(TcState& st, const SearchArgs& a, cudaStream_t stream, int sm_count) {
    init(st);
    const Plan pl = make_plan(st, a.dtype, a.nq, a.k, a.n, a.d, a.pitch, sm_count);
    const bool is_int = pl.is_int, pair = pl.pair;
    const uint32_t qpitch = pl.qpitch, mtiles = pl.mtiles, ntiles = pl.ntiles, chunks = pl.chunks, tiles_per_chunk = pl.tiles_per_chunk;
    const uint32_t CL = pl.CL, mgroups = pl.mgroups, items = pl.items, grid = pl.grid, kblocks = pl.kblocks, cap = pl.cap;
    const uint32_t sync_window = pl.sync_window;
    const int EG = pl.EG, stages = pl.stages;

    st.qprep.reserve((size_t)mtiles * BM * qpitch);
    st.qconst.reserve((size_t)a.nq * 4);
    st.cand.reserve((size_t)grid * EG * BM * cap * 8);
    st.lists.reserve((size_t)a.nq * chunks * EG * a.k * 8);
    st.gthr.reserve((size_t)mgroups * CL * BM * 4);
    TC_CUDA_CHECK(cudaMemsetAsync(st.gthr.p, 0xff, (size_t)mgroups * CL * BM * 4, stream));
    if (sync_window) {
        st.progress.reserve((size_t)items * 4);
        TC_CUDA_CHECK(cudaMemsetAsync(st.progress.p, 0, (size_t)items * 4, stream));
    }
    uint64_t launches = 0;
    TC_CUDA_CHECK(cudaMemsetAsync(st.qprep.p, 0, (size_t)mtiles * BM * qpitch, stream));
    if (is_int) {
        if (a.dtype == k::DT_I8)
            prep_queries<true, true><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
        else
            prep_queries<true, false><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
    } else {
        prep_queries<false, false><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
    }
    TC_CUDA_CHECK(cudaGetLastError());
    launches++;

    CUtensorMap map_q, map_x;
    make_map(st, &map_q, st.qprep.p, a.dtype, (uint64_t)mtiles * BM, a.d, qpitch, BM);
    make_map(st, &map_x, a.rows, a.dtype, a.n, a.d, a.pitch, BN / CL);

    TcParams p;
    p.norm = a.norm; p.sq = a.sq; p.mask = a.mask; p.cand = (u64*)st.cand.p; p.out_lists = (u64*)st.lists.p; p.dump = nullptr;
    p.gthr = (uint32_t*)st.gthr.p;
    p.n = a.n; p.nq = a.nq; p.k = a.k; p.cap = cap;
    p.trigger = pl.trigger;
    p.kblocks = kblocks;
    p.mtiles = mtiles; p.ntiles = ntiles; p.chunks = chunks; p.tiles_per_chunk = tiles_per_chunk;
    p.metric = a.metric; p.is_int = is_int ? 1 : 0;
    const uint32_t fmt = a.dtype == k::DT_I8 ? 1u : 0u;     // kind::i8: 0 = unsigned, 1 = signed; kind::f16: 0 = F16
    const uint32_t cfmt = is_int ? 2u : 1u;                  // accumulator: S32 / F32
    p.l2_prefetch_tiles = (uint32_t)st.l2_prefetch;
    p.stages = (uint32_t)stages;
    p.progress = (uint32_t*)st.progress.p; p.sync_window = sync_window;
    p.stats = nullptr;
#ifdef MVF_TC_STATS
    static Buf stats_buf;
    stats_buf.reserve(16 * 8);
    TC_CUDA_CHECK(cudaMemsetAsync(stats_buf.p, 0, 16 * 8, stream));
    p.stats = (unsigned long long*)stats_buf.p;
#endif
    // the pair form's instruction spans both CTAs: M = 256
    p.idesc = (cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
    if (a.dump) {
        st.dump.reserve((size_t)mtiles * BM * (size_t)ntiles * BN * 4);
        p.dump = (float*)st.dump.p;
    }
    const bool l2 = a.metric == k::M_L2;
    const void* fn = select_kernel(is_int, l2, (int)CL, EG, pair);
    const size_t smem = pl.smem;
    TC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&map_q, (void*)&map_x, (void*)&p};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads_for(EG)); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TC_CUDA_CHECK(cudaLaunchKernelExC(&cfg, fn, args));
    TC_CUDA_CHECK(cudaGetLastError());
    launches++;
#ifdef MVF_TC_STATS
    {   // diagnostic build: where each role's cycles went (sums over CTAs; one producer, one issuer, one epilogue warp each)
        unsigned long long h[16];
        TC_CUDA_CHECK(cudaStreamSynchronize(stream));
        TC_CUDA_CHECK(cudaMemcpy(h, stats_buf.p, sizeof(h), cudaMemcpyDeviceToHost));
        auto pct = [&](int part, int whole) { return h[whole] ? 100.0 * (double)h[part] / (double)h[whole] : 0.0; };
        std::fprintf(stderr, "[tc stats] producer: wait-empty %.1f%% pacing %.1f%% | issuer: wait-full %.1f%% wait-accumulator %.1f%% | "
                             "epilogue warp: wait-accumulator-ready %.1f%% compaction %.1f%% item-tail %.1f%%\n",
                     pct(1, 0), pct(2, 0), pct(4, 3), pct(5, 3), pct(7, 6), pct(8, 6), pct(9, 6));
    }
#endif
    SearchResult r;
    r.lists = (const u64*)st.lists.p; r.chunks = chunks * EG; r.qconst = (const float*)st.qconst.p;
    r.dump = (const float*)st.dump.p; r.dump_cols = ntiles * BN; r.launches = launches;
    return r;
}

inline void release(TcState& st) {
    st.qprep.release(); st.qconst.release(); st.cand.release(); st.lists.release(); st.dump.release(); st.gthr.release(); st.progress.release();
}

}  // namespace tc
}  // namespace mvf
