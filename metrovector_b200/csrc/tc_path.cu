// tc_path.cu -- host side of K2 (tc_kernels.cuh): TMA tensor maps, work partitioning, launch.
#define MVF_TC_PATH_TU
#include "tc_kernels.cuh"

namespace mvf {
namespace tc {

void init(TcState& st) {
    read_env(st);
    if (st.encode) return;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) { cudaGetLastError(); throw MvfError::cuda("cuTensorMapEncodeTiled is not available from the driver"); }
    st.encode = (EncodeTiledFn)fn;
}

static void make_map(TcState& st, CUtensorMap* map, const void* base, int kind, uint64_t rows, uint32_t d, uint64_t pitch,
                     uint32_t box_rows) {
    cuuint64_t gdim[2] = {d, rows};
    cuuint64_t gstride[1] = {pitch};
    cuuint32_t box[2] = {(cuuint32_t)(BKB / elem_bytes(kind)), box_rows};      // one 128-byte swizzle row of K
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = kind == KIND_I8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (kind == KIND_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    CUresult r = st.encode(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw MvfError::cuda("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

void split_rows_tf32(const unsigned char* rows, uint64_t n, uint32_t pitch, unsigned char* out, cudaStream_t stream) {
    if (!n) return;
    split_rows_tf32_kernel<<<148 * 16, 256, 0, stream>>>(rows, n * (uint64_t)pitch, out);
    TC_CUDA_CHECK(cudaGetLastError());
}

SearchResult run(TcState& st, const SearchArgs& a, cudaStream_t stream, int sm_count) {
    init(st);
    const Plan pl = make_plan(st, a.dtype, a.nq, a.k, std::max<uint64_t>(a.n, 1), a.d, a.pitch, sm_count);
    const int kind = pl.kind;
    const bool pair = pl.pair;
    const uint32_t qpitch = pl.qpitch, mtiles = pl.mtiles, ntiles = pl.ntiles, chunks = pl.chunks, tiles_per_chunk = pl.tiles_per_chunk;
    const uint32_t CL = pl.CL, mgroups = pl.mgroups, items = pl.items, grid = pl.grid, kblocks = pl.kblocks, cap = pl.cap;
    const uint32_t sync_window = pl.sync_window;
    const int EG = pl.EG, stages = pl.stages;
    if (kind == KIND_TF32 && a.n && !a.rows_split) throw MvfError::build_error("the split-tf32 planes of this Float32 space are not resident");
    if (kind == KIND_I8 && a.metric == k::M_L2 && a.n && !a.acc_seed) throw MvfError::build_error("the accumulator seeds of this integer space are not resident");
    if (pl.smem > 227 * 1024) throw MvfError::build_error("no tensor-core launch shape fits shared memory for this k");

    const size_t qplane = (size_t)mtiles * BM * qpitch;                  // one plane of the prepared queries
    st.qprep.reserve(qplane * planes_of(kind));
    st.qconst.reserve((size_t)a.nq * 4);
    st.cand.reserve((size_t)grid * EG * BM * cap * 8);
    st.lists.reserve((size_t)a.nq * chunks * EG * a.k * 8);
    // per-query bounds shared by all CTAs: gthr [rows] followed by gslot [rows][TC_SLOTS], all "nothing known yet"
    const size_t bound_rows = (size_t)mgroups * CL * BM;
    st.gthr.reserve(bound_rows * 4 * (1 + TC_SLOTS));
    TC_CUDA_CHECK(cudaMemsetAsync(st.gthr.p, 0xff, bound_rows * 4 * (1 + TC_SLOTS), stream));
    if (sync_window) {
        st.progress.reserve((size_t)items * 4);
        TC_CUDA_CHECK(cudaMemsetAsync(st.progress.p, 0, (size_t)items * 4, stream));
    }
    uint64_t launches = 0;
    TC_CUDA_CHECK(cudaMemsetAsync(st.qprep.p, 0, qplane * planes_of(kind), stream));
    if (kind == KIND_I8) {
        if (a.dtype == k::DT_I8)
            prep_queries<true, true><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
        else
            prep_queries<true, false><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
    } else if (kind == KIND_F16) {
        prep_queries<false, false><<<a.nq, 128, 0, stream>>>(a.queries, (unsigned char*)st.qprep.p, a.nq, a.d, qpitch, (float*)st.qconst.p, a.metric);
    } else {
        prep_queries_tf32<<<a.nq, 128, 0, stream>>>((const float*)a.queries, (unsigned char*)st.qprep.p, qplane, a.nq, a.d, qpitch,
                                                    (float*)st.qconst.p, a.metric);
    }
    TC_CUDA_CHECK(cudaGetLastError());
    launches++;
    if (a.n == 0) {
        // empty shard of a sharded space whose other ranks are on this path: only the per-query constants are needed
        // (the final merge decodes every rank's keys with them); the shard contributes one empty list per query
        st.lists.reserve((size_t)a.nq * a.k * 8);
        TC_CUDA_CHECK(cudaMemsetAsync(st.lists.p, 0xff, (size_t)a.nq * a.k * 8, stream));
        SearchResult r0;
        r0.lists = (const u64*)st.lists.p; r0.chunks = 1; r0.qconst = (const float*)st.qconst.p; r0.dump = nullptr; r0.dump_cols = 0;
        r0.launches = launches;
        return r0;
    }

    CUtensorMap map_q, map_x, map_q2, map_x2;
    make_map(st, &map_q, st.qprep.p, kind, (uint64_t)mtiles * BM, a.d, qpitch, BM);
    make_map(st, &map_x, kind == KIND_TF32 ? a.rows_split : a.rows, kind, a.n, a.d, a.pitch, BN / CL);
    if (kind == KIND_TF32) {
        make_map(st, &map_q2, (unsigned char*)st.qprep.p + qplane, kind, (uint64_t)mtiles * BM, a.d, qpitch, BM);
        make_map(st, &map_x2, a.rows_split + (size_t)a.n * a.pitch, kind, a.n, a.d, a.pitch, BN / CL);
    } else { map_q2 = map_q; map_x2 = map_x; }

    TcParams p;
    p.norm = a.norm; p.sq = a.sq; p.mask = a.mask; p.acc_seed = a.acc_seed; p.qconst = (const float*)st.qconst.p; p.cand = (u64*)st.cand.p; p.out_lists = (u64*)st.lists.p; p.dump = nullptr;
    p.gthr = (uint32_t*)st.gthr.p;
    p.gslot = p.gthr + bound_rows;
    p.n = a.n; p.nq = a.nq; p.k = a.k; p.cap = cap;
    p.trigger = pl.trigger;
    p.kblocks = kblocks;
    p.mtiles = mtiles; p.ntiles = ntiles; p.chunks = chunks; p.tiles_per_chunk = tiles_per_chunk;
    p.metric = a.metric; p.kind = kind;
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp layout): c_format [4,6) F32 = 1 / S32 = 2; a_format [7,10),
    // b_format [10,13): kind::f16 F16 = 0, kind::tf32 TF32 = 2, kind::i8 unsigned = 0 / signed = 1; both operands K-major;
    // N >> 3 at [17,23), M >> 4 at [24,29).  The pair form's instruction spans both CTAs: M = 256.
    const uint32_t fmt = kind == KIND_TF32 ? 2u : (a.dtype == k::DT_I8 ? 1u : 0u);
    const uint32_t cfmt = kind == KIND_I8 ? 2u : 1u;
    p.l2_prefetch_tiles = (uint32_t)st.l2_prefetch;
    p.stages = (uint32_t)stages;
    p.progress = (uint32_t*)st.progress.p; p.sync_window = sync_window;
    p.stats = nullptr;
#ifdef MVF_TC_STATS
    st.stats.reserve(16 * 8);
    TC_CUDA_CHECK(cudaMemsetAsync(st.stats.p, 0, 16 * 8, stream));
    p.stats = (unsigned long long*)st.stats.p;
#endif
    p.idesc = (cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
    if (a.dump) {
        st.dump.reserve((size_t)mtiles * BM * (size_t)ntiles * BN * 4);
        p.dump = (float*)st.dump.p;
    }
    const int mode = a.metric == k::M_L2 ? 0 : ((kind == KIND_I8 && a.metric == k::M_COS) ? 3 : 1);
    const void* fn = select_kernel(kind, EG, mode, (int)CL, pair);
    const size_t smem = pl.smem;
    // per device and cheap; set on every launch so a second engine on another GPU of this process gets it too
    TC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&map_q, (void*)&map_x, (void*)&map_q2, (void*)&map_x2, (void*)&p};
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
        TC_CUDA_CHECK(cudaMemcpy(h, st.stats.p, sizeof(h), cudaMemcpyDeviceToHost));
        auto pct = [&](int part, int whole) { return h[whole] ? 100.0 * (double)h[part] / (double)h[whole] : 0.0; };
        std::fprintf(stderr, "[tc stats] producer: wait-empty %.1f%% pacing %.1f%% | issuer: wait-full %.1f%% wait-accumulator %.1f%% | "
                             "epilogue warp: wait-accumulator-ready %.1f%% drain %.1f%% seed+handback %.1f%% process %.1f%% publish %.1f%% compaction %.1f%% item-tail %.1f%% | "
                             "cycles per tile-job: %.0f (grid %u, %u tiles x %u query tiles)\n",
                     pct(1, 0), pct(2, 0), pct(4, 3), pct(5, 3), pct(7, 6), pct(10, 6), pct(11, 6), pct(12, 6), pct(13, 6), pct(8, 6), pct(9, 6),
                     (double)h[6] / ((double)ntiles * (double)mtiles),      // every CTA's warp 2 adds its total; a CTA runs ntiles * mtiles / grid tile-jobs
                     grid, ntiles, mtiles);
    }
#endif
    SearchResult r;
    r.lists = (const u64*)st.lists.p; r.chunks = chunks * EG; r.qconst = (const float*)st.qconst.p;
    r.dump = (const float*)st.dump.p; r.dump_cols = ntiles * BN; r.launches = launches;
    return r;
}

void release(TcState& st) {
    st.qprep.release(); st.qconst.release(); st.cand.release(); st.lists.release(); st.dump.release(); st.gthr.release();
    st.progress.release(); st.stats.release();
}

}  // namespace tc
}  // namespace mvf
