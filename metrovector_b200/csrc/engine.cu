// engine.cu -- the GPU engine behind the C ABI: residency (mmap -> HBM once per space, with row
// norms), kernel dispatch for find_top_k_similar (examples/similarity_search.rs:140-176 of the
// reference), and the row-sharded multi-GPU merge (one exchange of nq*k 64-bit candidates per rank: a peer-memory
// kernel for small payloads, ncclAllGather above).
// An engine is bound to one B200 (mvf_gpu_init; several processes then form a communicator with
// mvf_gpu_comm_init) or, from ONE process, to several (mvf_gpu_init_multi: one member engine per device,
// ncclCommInitAll, the same search phases driven device by device).  There is no CPU compute fallback:
// without a device every entry point fails with MVF_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <memory>
#include <new>
#include <set>
#include <thread>
#include <vector>

#include "c_api_common.hpp"
#define MVF_ENGINE_TU
#include "kernels.cuh"
#include "tc_path.hpp"

using namespace mvf;
using mvf::k::u64;

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw MvfError::cuda(std::string(cudaGetErrorName(_e)) + ": " + cudaGetErrorString(_e) + " at " #expr); \
    } while (0)

// ---- NCCL through dlopen: no link-time dependency; inside a torch process this resolves to the
// libnccl.so.2 torch already loaded ------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (!api.handle) {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) throw MvfError::nccl(std::string("cannot load libnccl: ") + dlerror());
        api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
        api.CommInitAll = (int (*)(NcclComm*, int, const int*))dlsym(api.handle, "ncclCommInitAll");
        api.GroupStart = (int (*)())dlsym(api.handle, "ncclGroupStart");
        api.GroupEnd = (int (*)())dlsym(api.handle, "ncclGroupEnd");
        api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllGather");
        api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.GroupStart || !api.GroupEnd || !api.AllGather || !api.CommDestroy)
            throw MvfError::nccl("libnccl is missing required symbols");
    }
    return api;
}
void nccl_check(int rc, const char* what) {
    if (rc != 0) {
        const char* s = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
        throw MvfError::nccl(std::string(what) + ": " + s);
    }
}
constexpr int kNcclUint64 = 5;   // ncclUint64, nccl.h

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CUDA_CHECK(cudaMalloc(&p, n));
        cap = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
uint32_t pow2ceil(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
}  // namespace

struct mvf_gpu {
    // single-process multi-device engine (mvf_gpu_init_multi): the handle the caller holds is a GROUP whose members
    // are ordinary per-device engines (rank i of world n, communicator from ncclCommInitAll); a group has no device
    // state of its own
    std::vector<mvf_gpu*> members;
    bool is_group() const { return !members.empty(); }
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    NcclComm comm = nullptr;
    int rank = 0, world = 1;
    int path = 0;
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    double scan_ms_total = 0.0;
    uint64_t scan_launches = 0, all_launches = 0;
    int last_path = 0;
    int ctas_per_sm_cap = 0;   // MVF_SCAN_CTAS_PER_SM, 0 = occupancy
    DevBuf scan_counter;                   // "CTAs done" word of the fused K1 + K3 launch (zero between launches)
    bool no_fused_merge = false;           // MVF_NO_FUSED_MERGE: always launch K3 separately (A/B)
    bool block_sync = false;               // MVF_SYNC_MODE=block: wait for results with cudaStreamSynchronize instead of polling
    DevBuf cand, packed_local, gathered, queries, out_ids, out_scores, out_iscores, merge_tmp[2], gather_ids, gather_out, crc_tables, scalar_tmp;
    mvf_gpu* parent_group = nullptr;       // set on the members of a group engine
    int merge_tmp_flip = 0;
    unsigned char* pinned_out = nullptr;   // 64 KiB mapped pinned block small results are written into by the merge kernel
    unsigned char* pinned_out_dev = nullptr;
    std::vector<unsigned char*> stage_pool;   // pinned staging buffers (16 MiB each), kept across uploads
    int upload_threads = 8;                // MVF_UPLOAD_THREADS (default: host cores - 4, between 4 and 12)
    std::map<std::pair<const void*, size_t>, int> occ_cache;
    std::map<const void*, size_t> smem_optin_max;   // dynamic shared-memory limit raised so far per kernel, on THIS device
    uint32_t* pinned_flag = nullptr;       // mapped pinned word the device writes "this query batch holds Inf/NaN" into
    uint32_t* pinned_flag_dev = nullptr;
    cudaEvent_t xdev_event = nullptr;      // group engines: orders a member's stream against another member's
    // X1, the exchange over peer memory (kernels.cuh exchange_merge): this rank's window, every rank's window as mapped
    // into this device, and the search counter that selects buffer and flag.  `ready` is agreed by all ranks at set-up.
    struct PeerExchange {
        bool ready = false, ipc = false;
        unsigned char* window = nullptr;
        std::vector<unsigned char*> peer;      // [world]
        DevBuf peer_table;                     // the same pointers on the device
        uint64_t seq = 0;                      // peer-exchange searches so far (identical on every rank)
        uint32_t* error = nullptr;             // mapped pinned word the kernel raises when a peer's cells never arrive
        uint32_t* error_dev = nullptr;
        std::string note;                      // why the peer path is off, if it is
    } px;
    uint64_t peer_searches = 0, nccl_searches = 0;
    mvf::tc::TcState tc;
    // one stream + one set of scratch buffers per engine: every entry point that touches them takes this lock, so
    // calls from several host threads are serialised instead of racing (SURVEY.md 8b "Threading")
    std::recursive_mutex mu;
};
using EngineLock = std::lock_guard<std::recursive_mutex>;

struct mvf_gpu_space {
    mvf_gpu* g = nullptr;
    std::vector<mvf_gpu_space*> shards;    // space of a group engine: one resident shard per member, nothing else
    bool is_group() const { return !shards.empty(); }
    int dtype = 0, metric = 0;
    uint64_t local_n = 0, row_base = 0, global_n = 0;
    uint32_t d = 0, pitch = 0;
    unsigned char* rows = nullptr;
    float* norm = nullptr;
    float* sq = nullptr;
    uint32_t* mask = nullptr;  // optional row mask (bit set = row excluded): tombstones OR pre-filter, what the kernels read
    uint64_t masked_rows = 0;
    std::vector<uint32_t> mask_part[2];   // host copies of the two components, local rows: [0] tombstones, [1] filter
    uint64_t* ext_ids = nullptr;          // optional external id per GLOBAL row (vector_ids_block_index)
    int* acc_seed = nullptr;              // integer spaces: accumulator seeds of the tensor path's L2 form (built on first use, dropped when the mask changes)
    unsigned char* rows_split = nullptr;  // Float32 spaces: the two tf32-exact planes [hi | lo] the batched path multiplies (K2')
    // Facts every rank of a sharded space agrees on (exchanged at upload): ranks that chose different kernels would
    // exchange candidate keys in different encodings.
    uint64_t* rank_base = nullptr;        // [world] first global row of every rank's shard (device; cross-rank merges add it to the
                                          // 32-bit LOCAL row ids the candidates carry, so only a shard is limited to 2^32 rows)
    uint64_t min_shard_rows = 0;          // smallest non-empty shard over all ranks (spaces not cut by mvf_shard_range)
    bool ranged = false;                  // shard = mvf_shard_range(global_n, world, rank): the path is chosen from global_n
    uint64_t nonfinite_rows = 0;          // float spaces: rows with Inf/NaN components, over ALL shards
    bool split_everywhere = false;        // Float32 spaces: every rank holds its tf32 planes
    int upload_path = 0;       // 0 none/synthetic, 1 cudaHostRegister'd mapping, 2 pinned staging
    double upload_seconds = 0.0;
    std::string upload_note;   // why the in-place registration was not used, if it was not
};

namespace {

void flush_profile(mvf_gpu* g) {
    for (size_t i = 0; i < g->ev_used; i++) {
        float ms = 0.f;
        cudaEventSynchronize(g->ev_pool[i].second);
        if (cudaEventElapsedTime(&ms, g->ev_pool[i].first, g->ev_pool[i].second) == cudaSuccess) g->scan_ms_total += ms;
    }
    g->ev_used = 0;
}
// The host's wait for a search's result.  cudaStreamSynchronize backs off while it waits, and after a scan of tens of
// milliseconds its wake-up costs 50-100 us -- 3 % of a 2.8 ms search on 8 GPUs; polling cudaStreamQuery returns within
// a few microseconds of completion.  The poll gives the core away (sched_yield) once the wait is long, and
// MVF_SYNC_MODE=block restores the blocking call.
void wait_stream(mvf_gpu* g) {
    if (g->block_sync) { CUDA_CHECK(cudaStreamSynchronize(g->stream)); return; }
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    for (;;) {
        const cudaError_t e = cudaStreamQuery(g->stream);
        if (e == cudaSuccess) return;
        if (e != cudaErrorNotReady) CUDA_CHECK(e);
        if ((++spins & 63u) == 0) {
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (sec > 2.0) { CUDA_CHECK(cudaStreamSynchronize(g->stream)); return; }     // not a search any more: stop burning the core
            if (sec > 0.001) sched_yield();
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

struct ScanTimer {   // brackets the dominant kernel with events when profiling is on
    mvf_gpu* g; bool on = false; size_t idx = 0;
    explicit ScanTimer(mvf_gpu* gg) : g(gg) {
        if (!g->profile) return;
        if (g->ev_used == g->ev_pool.size()) {
            if (g->ev_pool.size() >= 8192) flush_profile(g);
            if (g->ev_used == g->ev_pool.size()) {
                cudaEvent_t a, b; CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
                g->ev_pool.push_back({a, b});
            }
        }
        idx = g->ev_used++; on = true;
        cudaEventRecord(g->ev_pool[idx].first, g->stream);
    }
    ~ScanTimer() { if (on) cudaEventRecord(g->ev_pool[idx].second, g->stream); }
};

// the dynamic shared-memory opt-in is a per-device function attribute: remembered per engine, not per process
// and only ever raised (a later, smaller request must not lower the limit under a launch shape cached earlier).
void allow_smem(mvf_gpu* g, const void* fn, size_t bytes) {
    size_t& have = g->smem_optin_max[fn];
    if (bytes <= have) return;
    CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
}

int occupancy(mvf_gpu* g, const void* fn, int threads, size_t smem) {
    auto key = std::make_pair(fn, smem);
    auto it = g->occ_cache.find(key);
    if (it != g->occ_cache.end()) return it->second;
    if (smem > 48 * 1024) allow_smem(g, fn, smem);
    int nb = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem));
    if (nb < 1) throw MvfError::cuda("kernel does not fit on an SM (shared memory " + std::to_string(smem) + " B)");
    g->occ_cache[key] = nb;
    return nb;
}

// ---- K1 dispatch ------------------------------------------------------------------------------
template <int DT> constexpr bool dt_is_float() { return DT == k::DT_F32 || DT == k::DT_F16; }

struct ScanShape { bool wide; int LPR; uint32_t dpad; };
template <int DT> ScanShape scan_shape(uint32_t pitch) {
    const uint32_t C = pitch >> 4;
    ScanShape s;
    if (C > 16) { s.wide = true; s.LPR = 32; s.dpad = ((C + 31) / 32) * 32 * k::Traits<DT>::EPC; }
    else { s.wide = false; s.LPR = (int)pow2ceil(C); s.dpad = (uint32_t)s.LPR * k::Traits<DT>::EPC; }
    return s;
}
template <int DT> size_t scan_smem(int NQ, uint32_t P, uint32_t dpad) {
    size_t q_bytes = (size_t)NQ * dpad * (dt_is_float<DT>() ? 4 : 1);
    return (size_t)NQ * P * 8 + ((q_bytes + 15) & ~(size_t)15) + (size_t)NQ * sizeof(k::QueryConst) + 64;
}

struct FinalOut { uint64_t* ids; float* scores; int32_t* iscores; };
bool launch_scan(mvf_gpu* g, const mvf_gpu_space* s, k::ScanParams p, int NQ, uint32_t* grid_out, k::MergeParams* fuse,
                 const k::ExchangeParams* xchg = nullptr, const FinalOut* fin = nullptr);
size_t scan_smem_any(int dtype, int NQ, uint32_t P, uint32_t pitch) {
    switch (dtype) {
        case k::DT_F32: return scan_smem<k::DT_F32>(NQ, P, scan_shape<k::DT_F32>(pitch).dpad);
        case k::DT_F16: return scan_smem<k::DT_F16>(NQ, P, scan_shape<k::DT_F16>(pitch).dpad);
        case k::DT_I8: return scan_smem<k::DT_I8>(NQ, P, scan_shape<k::DT_I8>(pitch).dpad);
        default: return scan_smem<k::DT_U8>(NQ, P, scan_shape<k::DT_U8>(pitch).dpad);
    }
}

// Launches K1.  With `fuse` (the K3 parameters of the same launch, `in`/`q_stride`/`nlists` filled in here) the
// per-CTA lists are merged by the last CTA of the SAME launch whenever they fit its shared memory (<= 1024 lists,
// <= 96 KiB): one launch per search instead of two -- what the latency-bound configs are made of.  Returns whether
// the merge was fused; if not the caller launches K3.
// With `xchg` (sharded search, small payload) a fused launch also carries the exchange: the last CTA pushes the merged list
// into every rank's window, waits for the other ranks' lists and writes the FINAL result (`fin`) -- one launch per search
// on every GPU.
bool launch_scan(mvf_gpu* g, const mvf_gpu_space* s, k::ScanParams p, int NQ, uint32_t* grid_out, k::MergeParams* fuse,
                 const k::ExchangeParams* xchg, const FinalOut* fin) {
    ScanShape sh;
    const void* fn = nullptr;
    const int op = p.metric == k::M_L2 ? k::OP_L2 : k::OP_DOT;
    switch (s->dtype) {
        case k::DT_F32: sh = scan_shape<k::DT_F32>(s->pitch); fn = k::scan_kernel_f32(op, NQ, sh.wide); break;
        case k::DT_F16: sh = scan_shape<k::DT_F16>(s->pitch); fn = k::scan_kernel_f16(op, NQ, sh.wide); break;
        case k::DT_I8: sh = scan_shape<k::DT_I8>(s->pitch); fn = k::scan_kernel_i8(op, NQ, sh.wide); break;
        default: sh = scan_shape<k::DT_U8>(s->pitch); fn = k::scan_kernel_u8(op, NQ, sh.wide); break;
    }
    size_t smem = scan_smem_any(s->dtype, NQ, p.P, s->pitch);
    int nb = occupancy(g, fn, k::kScanThreads, smem);
    if (g->ctas_per_sm_cap > 0) nb = std::min(nb, g->ctas_per_sm_cap);
    const int R = k::scan_rows_in_flight(NQ);
    const uint64_t rows_per_group = sh.wide ? (uint64_t)R : (uint64_t)R * (32 / sh.LPR);
    const uint64_t groups = (s->local_n + rows_per_group - 1) / rows_per_group;
    const uint64_t want = (groups + k::kScanWarps - 1) / k::kScanWarps;
    uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)g->sm_count * nb, want));
    k::ScanTail tail;
    tail.counter = nullptr;
    tail.exchange = 0;
    bool fused = false;
    // (the last CTA merges the NQ queries one after the other, k outputs each, while the rest of the chip is done: only
    // worth it while that tail is short -- otherwise K3 gets its own launch and one CTA per query)
    if (fuse && !g->no_fused_merge && (uint32_t)NQ * p.k <= 128) {
        auto merge_smem = [&](uint32_t lists) { return ((size_t)lists + k::kScanWarps) * p.k * 8 + 16; };
        size_t smem_f = std::max(smem, merge_smem(grid));
        if (xchg) smem_f = std::max(smem_f, ((size_t)xchg->world + 1 + k::kScanWarps) * p.k * 8 + 16);
        if (smem_f <= 96 * 1024) {
            int nb_f = occupancy(g, fn, k::kScanThreads, smem_f);          // a larger allocation may cost resident CTAs
            if (g->ctas_per_sm_cap > 0) nb_f = std::min(nb_f, g->ctas_per_sm_cap);
            const uint32_t grid_f = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)g->sm_count * nb_f, want));
            if (grid_f <= 1024 && grid_f <= grid) {
                grid = grid_f; smem = smem_f; fused = true;
                if (!g->scan_counter.p) {
                    g->scan_counter.reserve(16);
                    CUDA_CHECK(cudaMemsetAsync(g->scan_counter.p, 0, 16, g->stream));
                }
                fuse->in = p.cand; fuse->q_stride = (size_t)grid * p.k; fuse->list_stride = p.k; fuse->nlists = grid;
                tail.counter = (uint32_t*)g->scan_counter.p;
                tail.mp = *fuse;
                if (!xchg) tail.mp.list_base = nullptr;          // a local merge: ids count from id_base
                if (xchg) {
                    tail.exchange = 1; tail.x = *xchg;
                    tail.mp.out_packed = nullptr; tail.mp.out_ids = fin->ids; tail.mp.out_scores = fin->scores; tail.mp.out_iscores = fin->iscores;
                }
            }
        }
    }
    *grid_out = grid;
    ScanTimer t(g);
    int lpr = sh.LPR;
    void* args_wide[] = {(void*)&p, (void*)&tail};
    void* args_narrow[] = {(void*)&p, (void*)&lpr, (void*)&tail};
    CUDA_CHECK(cudaLaunchKernel(fn, dim3(grid), dim3(k::kScanThreads), sh.wide ? args_wide : args_narrow, smem, g->stream));
    g->scan_launches++; g->all_launches++;
    return fused;
}

void launch_merge(mvf_gpu* g, k::MergeParams mp, uint32_t nq) {
    // many queries: one warp per query straight from L2 (latency hidden by the other warps)
    if (nq >= 64 && mp.nlists > 0 && mp.nlists <= 32u * k::kMergeMaxHeads) {
        mp.seg_lists = mp.nlists; mp.P = 0;
        k::merge_topk_warp<<<(nq + 7) / 8, 256, 0, g->stream>>>(mp, nq);
        CUDA_CHECK(cudaGetLastError());
        g->all_launches++;
        return;
    }
    // fast flavour: the lists of one CTA staged in shared memory, two-level warp tournament; when
    // they do not fit, a segmented pass first reduces them to a few intermediate lists
    const size_t budget = std::min<size_t>(g->smem_optin, 200 * 1024);
    const size_t words = budget / 8;
    // (a cross-rank merge -- list_base set -- is never segmented: its words name (rank, position), which an intermediate
    // list would lose; with more ranks than fit one pass it takes the generic flavour below)
    const bool ranked = mp.list_base != nullptr;
    const bool one_pass = mp.nlists > 0 && (size_t)mp.k * (32 + 2) <= words &&
                          mp.nlists <= (uint32_t)std::min<size_t>((words - (size_t)(mp.nlists <= 32 ? 8u : 32u) * mp.k) / mp.k,
                                                                  (size_t)(mp.nlists <= 32 ? 8u : 32u) * 32 * k::kMergeMaxHeads);
    if (mp.nlists > 0 && (size_t)mp.k * (32 + 2) <= words && (!ranked || one_pass)) {
        allow_smem(g, (const void*)k::merge_topk_tournament, budget);
        for (;;) {
            const uint32_t nwarps = mp.nlists <= 32 ? 8u : 32u;
            const uint32_t fit = (uint32_t)std::min<size_t>((words - (size_t)nwarps * mp.k) / mp.k, (size_t)nwarps * 32 * k::kMergeMaxHeads);
            if (mp.nlists <= fit) {
                mp.seg_lists = mp.nlists; mp.P = 0;
                const size_t smem = ((size_t)mp.nlists + nwarps) * mp.k * 8;
                k::merge_topk_tournament<<<dim3(nq, 1), nwarps * 32, smem, g->stream>>>(mp);
                CUDA_CHECK(cudaGetLastError());
                g->all_launches++;
                return;
            }
            const uint32_t segs = (mp.nlists + fit - 1) / fit;
            const uint32_t seg_lists = (mp.nlists + segs - 1) / segs;
            g->merge_tmp[g->merge_tmp_flip].reserve((size_t)nq * segs * mp.k * 8);
            k::MergeParams sp = mp;
            sp.seg_lists = seg_lists; sp.P = 0; sp.id_base = 0;
            sp.out_packed = (u64*)g->merge_tmp[g->merge_tmp_flip].p; sp.out_ids = nullptr; sp.out_scores = nullptr; sp.out_iscores = nullptr;
            const size_t smem = ((size_t)seg_lists + nwarps) * mp.k * 8;
            k::merge_topk_tournament<<<dim3(nq, segs), nwarps * 32, smem, g->stream>>>(sp);
            CUDA_CHECK(cudaGetLastError());
            g->all_launches++;
            mp.in = sp.out_packed; mp.q_stride = (size_t)segs * mp.k; mp.list_stride = mp.k; mp.nlists = segs;
            g->merge_tmp_flip ^= 1;
        }
    }
    // generic flavour: rounds of block-wide bitonic sort
    const uint64_t total = (uint64_t)mp.nlists * mp.k + mp.k;
    uint32_t P = pow2ceil((uint32_t)std::min<uint64_t>(total, 16384));
    P = std::max<uint32_t>(P, std::max<uint32_t>(64, 2 * pow2ceil(mp.k)));
    P = std::min<uint32_t>(P, 16384);
    mp.P = P;
    const size_t smem = (size_t)P * 8;
    if (smem > 48 * 1024) allow_smem(g, (const void*)k::merge_topk, 16384 * 8);
    const int threads = (int)std::min<uint32_t>(1024, std::max<uint32_t>(32, P / 2));
    k::merge_topk<<<nq, threads, smem, g->stream>>>(mp);
    CUDA_CHECK(cudaGetLastError());
    g->all_launches++;
}

void validate_query(const mvf_gpu_space* s, int qdtype, uint32_t nq, uint32_t k, int metric) {
    if (nq == 0) throw MvfError::build_error("search needs at least one query");
    if (k == 0 || k > 1024) throw MvfError::build_error("k must be in 1..1024");
    if (metric < 0 || metric > 2) throw MvfError::build_error("unsupported distance metric (L2, InnerProduct, Cosine)");
    const bool fspace = s->dtype == k::DT_F32 || s->dtype == k::DT_F16;
    if (fspace) {
        if (!(qdtype == k::DT_F32 || (qdtype == k::DT_F16 && s->dtype == k::DT_F16)))
            throw MvfError::build_error("float spaces take Float32 queries (Float16 spaces also Float16)");
    } else if (qdtype != s->dtype) {
        throw MvfError::build_error("integer spaces take queries in their own data type");
    }
}

// ---- path choice ------------------------------------------------------------------------------
// Rows the scan-vs-tensor decision is taken on.  For a space sharded by mvf_shard_range every rank derives the same
// number -- the smallest non-empty shard -- from (global_n, world); a rank's own row count differs between ranks
// (the tail shard is short or empty) and must not decide: the ranks exchange packed candidates whose keys are the
// tensor kernel's raw ranking values on one path and final keys on the other.
uint64_t smallest_nonempty_shard(uint64_t global_n, int world) {
    if (world < 1) world = 1;
    const uint64_t per = (global_n + (uint64_t)world - 1) / (uint64_t)world;
    if (per == 0) return 0;
    const uint64_t tail = global_n % per;
    return tail ? tail : per;
}
uint64_t decision_rows(const mvf_gpu_space* s) {
    const mvf_gpu* g = s->g;
    if (g->world <= 1) return s->local_n;
    if (!s->ranged) return s->min_shard_rows ? s->min_shard_rows : s->local_n;   // caller-cut shards: exchanged at creation
    return smallest_nonempty_shard(s->global_n, g->world);
}
bool host_queries_finite(const void* q, int qdtype, size_t count) {
    if (qdtype == k::DT_F16) {
        const uint16_t* h = (const uint16_t*)q;
        uint32_t bad = 0;
        for (size_t i = 0; i < count; i++) bad |= (uint32_t)((h[i] & 0x7c00u) == 0x7c00u);
        return !bad;
    }
    const uint32_t* w = (const uint32_t*)q;
    uint32_t bad = 0;
    for (size_t i = 0; i < count; i++) bad |= (uint32_t)((w[i] & 0x7f800000u) == 0x7f800000u);
    return !bad;
}
// would this search take the tensor path if its queries are finite?  (everything except the queries' values)
bool tensor_candidate(const mvf_gpu_space* s, int qdtype, uint32_t nq, uint32_t kk, int metric) {
    const mvf_gpu* g = s->g;
    if (!mvf::tc::eligible(g->tc, g->path, s->dtype, qdtype, metric, nq, kk, decision_rows(s), s->d, s->pitch)) return false;
    const bool fspace = s->dtype == k::DT_F32 || s->dtype == k::DT_F16;
    if (fspace && s->nonfinite_rows != 0) return false;     // Inf/NaN rows: the scan kernel's element-wise arithmetic defines the result
    if (s->dtype == k::DT_F32 && !s->split_everywhere) return false;
    return true;
}
enum { QF_UNKNOWN = 0, QF_FINITE = 1, QF_NONFINITE = 2 };
// device-resident float queries of a tensor-path candidate: one tiny kernel raises a flag in mapped pinned memory and
// the host waits for it (the only host sync of mvf_gpu_search_device; mvf_gpu_search checks its host buffer instead)
int device_queries_finite(mvf_gpu* g, const void* d_queries, int qdtype, size_t count) {
    if (!g->pinned_flag) {
        CUDA_CHECK(cudaHostAlloc((void**)&g->pinned_flag, 64, cudaHostAllocMapped));
        CUDA_CHECK(cudaHostGetDevicePointer((void**)&g->pinned_flag_dev, g->pinned_flag, 0));
    }
    *g->pinned_flag = 0;
    const uint32_t grid = (uint32_t)std::min<size_t>((count + 2047) / 2048, (size_t)g->sm_count * 4);
    k::flag_nonfinite<<<std::max(1u, grid), 256, 0, g->stream>>>(d_queries, count, qdtype == k::DT_F16 ? 1 : 0, g->pinned_flag_dev);
    CUDA_CHECK(cudaGetLastError());
    g->all_launches++;
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    return *(volatile uint32_t*)g->pinned_flag ? QF_NONFINITE : QF_FINITE;
}

// X1 (exchange over peer memory, set-up further down): payloads of up to kPeerSlotCells candidates
constexpr uint32_t kPeerSlotCells = 512;                     // 16-byte cells per (buffer, rank) slot: 8 KiB
bool peer_exchange_fits(const mvf_gpu* g, uint32_t nq, uint32_t kk) {
    if (!g->px.ready) return false;
    return (size_t)nq * kk <= kPeerSlotCells && ((size_t)g->world + 1) * kk * 8 <= 48 * 1024;
}
// buffer and flag of the next peer exchange of this engine (one per search, the same sequence on every rank)
k::ExchangeParams next_exchange(mvf_gpu* g, bool do_merge) {
    auto& px = g->px;
    k::ExchangeParams x;
    x.mine = nullptr; x.peers = (uint4* const*)px.peer_table.p; x.window = (uint4*)px.window;
    x.world = (uint32_t)g->world; x.rank = (uint32_t)g->rank; x.slot_cells = kPeerSlotCells;
    x.buf = (uint32_t)(px.seq & 1); x.flag = (uint32_t)((px.seq >> 1) % 0xfffffffeull) + 1u;
    x.error = px.error_dev; x.do_merge = do_merge ? 1 : 0;
    px.seq++;
    return x;
}

// What one search leaves behind between its phases.
struct SearchCtx {
    uint32_t nq = 0, kk = 0; int metric = 0;
    bool multi = false;            // candidates go through the exchange (communicator present)
    bool tc_keys = false;          // packed keys are the tensor kernel's ranking values (decoded with qconst at the final merge)
    const float* tc_qconst = nullptr;
    u64* packed_local = nullptr;
    bool peer = false;             // the exchange goes over peer memory (X1); x holds this search's buffer and flag
    bool exchange_done = false;    // ... and already happened inside the scan launch (K1 + K3 + X1)
    k::ExchangeParams x;
};

// Phase 1: this shard's kernels and its local merge.  Without a communicator the local merge writes the final
// outputs; with one it leaves nq*k packed candidates (row ids local to the shard) for the exchange.
// Asynchronous on g->stream (except the finite check of device-resident float queries, see above).
void search_local(mvf_gpu_space* s, const void* d_queries, int qdtype, uint32_t nq, uint32_t kk, int metric, int qfinite,
                  uint64_t* d_ids, float* d_scores, int32_t* d_iscores, SearchCtx* ctx, bool final_here = true) {
    mvf_gpu* g = s->g;
    if (metric == MVF_METRIC_SPACE_DEFAULT) metric = s->metric;
    validate_query(s, qdtype, nq, kk, metric);
    CUDA_CHECK(cudaSetDevice(g->device));
    const bool multi = g->world > 1 && g->comm;
    u64* packed_local = nullptr;
    if (multi) { g->packed_local.reserve((size_t)nq * kk * 8); packed_local = (u64*)g->packed_local.p; }
    ctx->nq = nq; ctx->kk = kk; ctx->metric = metric; ctx->multi = multi; ctx->packed_local = packed_local;
    ctx->tc_keys = false; ctx->tc_qconst = nullptr;
    // (final_here = false: a group member whose merged result nobody reads -- it pushes and waits only)
    ctx->peer = multi && peer_exchange_fits(g, nq, kk);
    ctx->exchange_done = false;
    if (ctx->peer) ctx->x = next_exchange(g, final_here);

    bool use_tc = tensor_candidate(s, qdtype, nq, kk, metric);
    if (use_tc && (s->dtype == k::DT_F32 || s->dtype == k::DT_F16)) {
        if (qfinite == QF_UNKNOWN) qfinite = device_queries_finite(g, d_queries, qdtype, (size_t)nq * s->d);
        if (qfinite == QF_NONFINITE) use_tc = false;
    }
    if (use_tc) {
        if ((s->dtype == k::DT_I8 || s->dtype == k::DT_U8) && metric == k::M_L2 && !s->acc_seed && s->local_n) {
            const uint64_t padded = ((s->local_n + mvf::tc::BN - 1) / mvf::tc::BN) * mvf::tc::BN;
            CUDA_CHECK(cudaMalloc((void**)&s->acc_seed, padded * 4));
            k::build_acc_seed<<<(uint32_t)g->sm_count * 8, 256, 0, g->stream>>>(s->sq, s->mask, s->local_n, padded, (int)0x80000000 + (1 << 25), s->acc_seed);
            CUDA_CHECK(cudaGetLastError());
            g->all_launches++;
        }
        mvf::tc::SearchArgs a;
        a.acc_seed = s->acc_seed;
        a.rows = s->rows; a.rows_split = s->rows_split; a.norm = s->norm; a.sq = s->sq; a.mask = s->mask; a.n = s->local_n; a.d = s->d;
        a.pitch = s->pitch; a.dtype = s->dtype; a.metric = metric; a.queries = d_queries; a.nq = nq; a.k = kk; a.dump = false;
        mvf::tc::SearchResult r;
        {
            ScanTimer t(g);
            r = mvf::tc::run(g->tc, a, g->stream, g->sm_count);
        }
        g->scan_launches++; g->all_launches += r.launches;
        g->last_path = 2;
        // merge the per-item lists; keys are K2 ranking values, decoded with the per-query constant
        k::MergeParams mp;
        mp.in = r.lists; mp.q_stride = (size_t)r.chunks * kk; mp.list_stride = kk; mp.nlists = r.chunks; mp.k = kk;
        // (integer cosine: the tensor kernel's candidates already carry the exact final keys, nothing to decode)
        const bool final_keys = (s->dtype == k::DT_I8 || s->dtype == k::DT_U8) && metric == k::M_COS;
        mp.id_base = s->row_base; mp.metric = metric; mp.dtype = s->dtype; mp.tc_mode = final_keys ? 0 : 1; mp.qconst = r.qconst;
        mp.out_packed = multi ? packed_local : nullptr;
        mp.out_ids = multi ? nullptr : d_ids; mp.out_scores = multi ? nullptr : d_scores; mp.out_iscores = multi ? nullptr : d_iscores;
        mp.ext_ids = s->ext_ids;
        launch_merge(g, mp, nq);
        ctx->tc_keys = !final_keys; ctx->tc_qconst = r.qconst;
        return;
    }
    g->last_path = 1;
    const uint32_t P = std::max<uint32_t>(8, pow2ceil(k::kScanWarps * kk));
    // widest query group whose lists + staged queries fit comfortably in shared memory
    int NQmax = 8;
    while (NQmax > 1 && scan_smem_any(s->dtype, NQmax, P, s->pitch) > 96 * 1024) NQmax >>= 1;
    if (scan_smem_any(s->dtype, 1, P, s->pitch) > g->smem_optin)
        throw MvfError::build_error("dimension * k too large for the scan kernel's shared memory");
    const size_t grid_bound = (size_t)g->sm_count * 32;
    g->cand.reserve((size_t)NQmax * grid_bound * kk * 8);
    uint32_t q0 = 0;
    while (q0 < nq) {
        int NQ = NQmax;
        while ((uint32_t)NQ > nq - q0) NQ >>= 1;
        k::ScanParams p;
        p.rows = s->rows; p.norm = s->norm; p.mask = s->mask; p.queries = d_queries; p.cand = (u64*)g->cand.p;
        p.n = s->local_n; p.d = s->d; p.pitch = s->pitch; p.k = kk; p.P = P; p.metric = metric; p.qdtype = qdtype; p.q0 = q0;
        uint32_t grid = 0;
        k::MergeParams mp;
        mp.in = (const u64*)g->cand.p; mp.q_stride = 0; mp.list_stride = kk; mp.nlists = 0; mp.k = kk;
        mp.P = 0; mp.seg_lists = 0;
        mp.id_base = s->row_base; mp.metric = metric; mp.dtype = s->dtype; mp.tc_mode = 0; mp.qconst = nullptr;
        mp.list_base = s->rank_base;                             // read only by a fused exchange (K1 + K3 + X1)
        const size_t o = (size_t)q0 * kk;
        mp.out_packed = multi ? packed_local + o : nullptr;
        mp.out_ids = multi ? nullptr : d_ids + o;
        mp.out_scores = multi ? nullptr : (d_scores ? d_scores + o : nullptr);
        mp.out_iscores = multi ? nullptr : (d_iscores ? d_iscores + o : nullptr);
        mp.ext_ids = s->ext_ids;
        bool fused = false;
        // a sharded search that is ONE scan launch carries its exchange and the final merge in that launch
        const bool fuse_x = ctx->peer && (uint32_t)NQ == nq;
        const FinalOut fin = {d_ids, d_scores, d_iscores};
        if (s->local_n > 0) fused = launch_scan(g, s, p, NQ, &grid, &mp, fuse_x ? &ctx->x : nullptr, fuse_x ? &fin : nullptr);
        if (fused && fuse_x) { ctx->exchange_done = true; g->peer_searches++; }
        if (!fused) {
            mp.q_stride = (size_t)grid * kk; mp.nlists = grid; mp.list_base = nullptr;
            launch_merge(g, mp, (uint32_t)NQ);
        }
        q0 += (uint32_t)NQ;
    }
}

// Phase 2: the one exchange step (NCCL route) -- every rank contributes its nq*k best (shard-local row ids).  A group
// engine brackets the members' calls with ncclGroupStart/End.
void search_exchange(mvf_gpu_space* s, const SearchCtx& ctx) {
    mvf_gpu* g = s->g;
    if (!ctx.multi) return;
    CUDA_CHECK(cudaSetDevice(g->device));
    const size_t cnt = (size_t)ctx.nq * ctx.kk;
    g->gathered.reserve(cnt * 8 * (size_t)g->world);
    nccl_check(nccl().AllGather(ctx.packed_local, g->gathered.p, cnt, kNcclUint64, g->comm, g->stream), "ncclAllGather");
    g->all_launches++; g->nccl_searches++;
}

// Phase 3: merge of the W gathered lists per query; identical on every rank that runs it.
void search_final(mvf_gpu_space* s, const SearchCtx& ctx, uint64_t* d_ids, float* d_scores, int32_t* d_iscores) {
    mvf_gpu* g = s->g;
    if (!ctx.multi) return;
    CUDA_CHECK(cudaSetDevice(g->device));
    const size_t cnt = (size_t)ctx.nq * ctx.kk;
    k::MergeParams mp;
    mp.in = (const u64*)g->gathered.p; mp.q_stride = ctx.kk; mp.list_stride = cnt; mp.nlists = (uint32_t)g->world; mp.k = ctx.kk;
    mp.id_base = 0; mp.list_base = s->rank_base; mp.metric = ctx.metric; mp.dtype = s->dtype; mp.tc_mode = ctx.tc_keys ? 1 : 0; mp.qconst = ctx.tc_qconst;
    mp.out_packed = nullptr; mp.out_ids = d_ids; mp.out_scores = d_scores; mp.out_iscores = d_iscores;
    mp.ext_ids = s->ext_ids;
    launch_merge(g, mp, ctx.nq);
}


// ---- X1: exchange over peer memory ---------------------------------------------------------------------
// Small payloads (nq * k <= kPeerSlotCells words: every batch-1 search) skip NCCL: the local merge's nq*k words are
// pushed into every rank's window and merged there by ONE kernel (kernels.cuh exchange_merge) -- the collective and
// the merge that follows it are the same launch.  Large batches keep ncclAllGather (bandwidth, not latency).
size_t peer_window_bytes(int world) { return (size_t)2 * (size_t)world * kPeerSlotCells * 16; }

bool peer_exchange_wanted() {
    const char* e = std::getenv("MVF_PEER_EXCHANGE");
    return !(e && *e == '0');
}
void peer_exchange_release(mvf_gpu* g) {
    auto& px = g->px;
    if (px.ipc) for (int r = 0; r < (int)px.peer.size(); r++) if (r != g->rank && px.peer[(size_t)r]) cudaIpcCloseMemHandle(px.peer[(size_t)r]);
    px.peer.clear();
    if (px.window) cudaFree(px.window);
    px.window = nullptr;
    px.peer_table.release();
    if (px.error) cudaFreeHost(px.error);
    px.error = nullptr; px.error_dev = nullptr;
    px.ready = false;
}
// this rank's window (zeroed: flag 0 is never used) and the error word
void peer_exchange_alloc(mvf_gpu* g) {
    auto& px = g->px;
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaMalloc((void**)&px.window, peer_window_bytes(g->world)));
    CUDA_CHECK(cudaMemset(px.window, 0, peer_window_bytes(g->world)));
    CUDA_CHECK(cudaHostAlloc((void**)&px.error, 64, cudaHostAllocMapped));
    *px.error = 0;
    CUDA_CHECK(cudaHostGetDevicePointer((void**)&px.error_dev, px.error, 0));
    CUDA_CHECK(cudaDeviceSynchronize());
}
void peer_exchange_publish_table(mvf_gpu* g) {
    auto& px = g->px;
    CUDA_CHECK(cudaSetDevice(g->device));
    px.peer_table.reserve(sizeof(void*) * px.peer.size());
    CUDA_CHECK(cudaMemcpy(px.peer_table.p, px.peer.data(), sizeof(void*) * px.peer.size(), cudaMemcpyHostToDevice));
}
// One process per GPU: windows are shared with cudaIpc handles, carried to the other ranks by the communicator that
// was just created.  Collective; every rank ends with the same `ready` (all-gathered), so the ranks of a search always
// take the same exchange.
void peer_exchange_setup_ipc(mvf_gpu* g) {
    auto& px = g->px;
    const int W = g->world;
    uint64_t ok = 1;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    // (a rank that does not want the peer path still takes part in the two rounds below: set-up is collective)
    if (!peer_exchange_wanted()) { ok = 0; px.note = "MVF_PEER_EXCHANGE=0"; }
    else if (W > 128) { ok = 0; px.note = "more than 128 ranks"; }
    if (ok) { try { peer_exchange_alloc(g); } catch (const MvfError& e) { ok = 0; px.note = e.what(); cudaGetLastError(); } }
    if (ok && cudaIpcGetMemHandle(&mine, px.window) != cudaSuccess) { ok = 0; px.note = "cudaIpcGetMemHandle refused"; cudaGetLastError(); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    // round 1: handle + "I could export"; round 2: "I could map everybody"
    const size_t rec = 72;
    g->scalar_tmp.reserve(rec + rec * (size_t)W + 8 + 8 * (size_t)W);
    unsigned char* base = (unsigned char*)g->scalar_tmp.p;
    std::vector<unsigned char> sendrec(rec), all(rec * (size_t)W);
    std::memcpy(sendrec.data(), &mine, 64); std::memcpy(sendrec.data() + 64, &ok, 8);
    CUDA_CHECK(cudaMemcpyAsync(base, sendrec.data(), rec, cudaMemcpyHostToDevice, g->stream));
    nccl_check(nccl().AllGather(base, base + rec, rec / 8, kNcclUint64, g->comm, g->stream), "ncclAllGather");
    CUDA_CHECK(cudaMemcpyAsync(all.data(), base + rec, rec * (size_t)W, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    uint64_t everyone = 1;
    for (int r = 0; r < W; r++) { uint64_t o; std::memcpy(&o, all.data() + rec * (size_t)r + 64, 8); everyone &= o; }
    uint64_t mapped = everyone;
    if (everyone) {
        px.peer.assign((size_t)W, nullptr);
        px.peer[(size_t)g->rank] = px.window;
        px.ipc = true;
        for (int r = 0; r < W && mapped; r++) {
            if (r == g->rank) continue;
            cudaIpcMemHandle_t h; std::memcpy(&h, all.data() + rec * (size_t)r, 64);
            void* ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { mapped = 0; px.note = std::string("cudaIpcOpenMemHandle refused: ") + cudaGetErrorName(e); cudaGetLastError(); }
            else px.peer[(size_t)r] = (unsigned char*)ptr;
        }
    } else if (px.note.empty()) px.note = "a peer could not export its window";
    unsigned char* base2 = base + rec + rec * (size_t)W;
    std::vector<uint64_t> all2((size_t)W);
    CUDA_CHECK(cudaMemcpyAsync(base2, &mapped, 8, cudaMemcpyHostToDevice, g->stream));
    nccl_check(nccl().AllGather(base2, base2 + 8, 1, kNcclUint64, g->comm, g->stream), "ncclAllGather");
    CUDA_CHECK(cudaMemcpyAsync(all2.data(), base2 + 8, 8 * (size_t)W, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    uint64_t all_mapped = 1;
    for (uint64_t v : all2) all_mapped &= v;
    if (!all_mapped) { if (px.note.empty()) px.note = "a peer could not map the windows"; peer_exchange_release(g); return; }
    peer_exchange_publish_table(g);
    px.ready = true;
}
// One process, G devices: plain peer access between the members' devices.
void peer_exchange_setup_group(mvf_gpu* grp) {
    const size_t G = grp->members.size();
    auto off = [&](const std::string& why) { for (mvf_gpu* m : grp->members) { peer_exchange_release(m); m->px.note = why; } };
    if (!peer_exchange_wanted()) { off("MVF_PEER_EXCHANGE=0"); return; }
    if (G > 128) { off("more than 128 devices"); return; }
    for (size_t i = 0; i < G; i++)
        for (size_t j = 0; j < G; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, grp->members[i]->device, grp->members[j]->device) != cudaSuccess || !can) {
                cudaGetLastError(); off("devices without peer access"); return;
            }
        }
    try {
        for (size_t i = 0; i < G; i++) {
            CUDA_CHECK(cudaSetDevice(grp->members[i]->device));
            for (size_t j = 0; j < G; j++) {
                if (i == j) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(grp->members[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
                cudaGetLastError();
            }
            peer_exchange_alloc(grp->members[i]);
        }
        for (size_t i = 0; i < G; i++) {
            auto& px = grp->members[i]->px;
            px.peer.assign(G, nullptr);
            for (size_t j = 0; j < G; j++) px.peer[j] = grp->members[j]->px.window;
            peer_exchange_publish_table(grp->members[i]);
            px.ready = true;
        }
    } catch (const MvfError& e) { cudaGetLastError(); off(e.what()); }
}
// Phases 2 + 3 as one launch (when they did not already ride on the scan launch).
void search_exchange_peer(mvf_gpu_space* s, const SearchCtx& ctx, uint64_t* d_ids, float* d_scores, int32_t* d_iscores) {
    mvf_gpu* g = s->g;
    if (ctx.exchange_done) return;
    CUDA_CHECK(cudaSetDevice(g->device));
    k::ExchangeParams x = ctx.x;
    x.mine = ctx.packed_local;
    k::MergeParams mp;
    mp.in = nullptr; mp.q_stride = 0; mp.list_stride = 0; mp.nlists = (uint32_t)g->world; mp.k = ctx.kk; mp.P = 0; mp.seg_lists = 0;
    mp.id_base = 0; mp.list_base = s->rank_base; mp.metric = ctx.metric; mp.dtype = s->dtype; mp.tc_mode = ctx.tc_keys ? 1 : 0; mp.qconst = ctx.tc_qconst;
    mp.out_packed = nullptr; mp.out_ids = d_ids; mp.out_scores = d_scores; mp.out_iscores = d_iscores;
    mp.ext_ids = s->ext_ids;
    const size_t smem = ((size_t)g->world + 1) * ctx.kk * 8;
    k::exchange_merge<<<ctx.nq, 128, smem, g->stream>>>(x, mp);
    CUDA_CHECK(cudaGetLastError());
    g->all_launches++; g->peer_searches++;
}
void peer_exchange_check(mvf_gpu* g) {                       // after a stream sync: did a kernel give up on a peer?
    if (g->px.error && *(volatile uint32_t*)g->px.error) {
        *g->px.error = 0;
        throw MvfError::nccl("peer exchange timed out: a rank of the communicator did not deliver its candidates");
    }
}

// Full search of a single-device space with device-resident queries/outputs.
void search_device(mvf_gpu_space* s, const void* d_queries, int qdtype, uint32_t nq, uint32_t kk, int metric, int qfinite,
                   uint64_t* d_ids, float* d_scores, int32_t* d_iscores) {
    SearchCtx ctx;
    search_local(s, d_queries, qdtype, nq, kk, metric, qfinite, d_ids, d_scores, d_iscores, &ctx);
    if (ctx.peer) { search_exchange_peer(s, ctx, d_ids, d_scores, d_iscores); return; }
    search_exchange(s, ctx);
    search_final(s, ctx, d_ids, d_scores, d_iscores);
}

// ---- residency ----------------------------------------------------------------------------------
mvf_gpu_space* alloc_space(mvf_gpu* g, int dtype, int metric, uint64_t local_n, uint32_t d, uint64_t row_base,
                           uint64_t global_n) {
    if (dtype < 0 || dtype > 3) throw MvfError::build_error("Unsupported vector data type");
    if (metric == 255) throw MvfError::build_error("Custom distance metrics are not searchable");
    if (metric < 0 || metric > 2) throw MvfError::build_error("unsupported distance metric");
    if (d == 0) throw MvfError::build_error("vector space has dimension 0");
    if (local_n >= 0xffffffffull)             // candidates carry a 32-bit LOCAL row id; the space itself may be larger
        throw MvfError::build_error("a resident shard holds fewer than 2^32 rows (shard the space over more GPUs)");
    const size_t es = element_size((DataType)dtype);
    if ((dtype == k::DT_I8 || dtype == k::DT_U8) && d > 33025)
        throw MvfError::build_error("integer spaces support dimension <= 33025 (int32 accumulation)");
    const uint64_t pitch64 = ((uint64_t)d * es + 15) & ~15ull;
    if (pitch64 > (1u << 20)) throw MvfError::build_error("row too large (max 1 MiB)");
    CUDA_CHECK(cudaSetDevice(g->device));
    mvf_gpu_space* s = new mvf_gpu_space();
    s->g = g; s->dtype = dtype; s->metric = metric; s->local_n = local_n; s->row_base = row_base; s->global_n = global_n;
    s->d = d; s->pitch = (uint32_t)pitch64;
    try {
        // the scan kernels over-read up to one warp step past the last row instead of clamping: keep
        // zero rows (and zero norms) behind the data
        const size_t pad_rows = s->pitch <= 256 ? k::kPadRows : 8;
        const size_t bytes = ((size_t)local_n + pad_rows) * s->pitch;
        CUDA_CHECK(cudaMalloc((void**)&s->rows, bytes));
        CUDA_CHECK(cudaMalloc((void**)&s->norm, (local_n + k::kPadRows) * 4));
        CUDA_CHECK(cudaMalloc((void**)&s->sq, (local_n + k::kPadRows) * 4));
        CUDA_CHECK(cudaMemsetAsync(s->rows + (size_t)local_n * s->pitch, 0, pad_rows * s->pitch, g->stream));
        CUDA_CHECK(cudaMemsetAsync(s->norm + local_n, 0, k::kPadRows * 4, g->stream));
        CUDA_CHECK(cudaMemsetAsync(s->sq + local_n, 0, k::kPadRows * 4, g->stream));
    } catch (...) {
        cudaFree(s->rows); cudaFree(s->norm); cudaFree(s->sq); delete s; throw;
    }
    return s;
}

// K0 (+ for Float32 spaces the tf32 planes of K2').  Leaves in s->nonfinite_rows / s->split_everywhere what THIS
// shard knows; publish_facts() turns them into values all ranks agree on.
void compute_norms(mvf_gpu_space* s) {
    mvf_gpu* g = s->g;
    s->nonfinite_rows = 0; s->split_everywhere = false;
    if (s->local_n == 0) { s->split_everywhere = s->dtype == k::DT_F32; return; }
    g->scalar_tmp.reserve(16);
    unsigned long long* nf = (unsigned long long*)g->scalar_tmp.p;
    CUDA_CHECK(cudaMemsetAsync(nf, 0, 8, g->stream));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((s->local_n + 7) / 8, (uint64_t)g->sm_count * 8);
    switch (s->dtype) {
        case k::DT_F32: k::row_norms<k::DT_F32><<<grid, 256, 0, g->stream>>>(s->rows, s->local_n, s->pitch, s->norm, s->sq, nf); break;
        case k::DT_F16: k::row_norms<k::DT_F16><<<grid, 256, 0, g->stream>>>(s->rows, s->local_n, s->pitch, s->norm, s->sq, nf); break;
        case k::DT_I8: k::row_norms<k::DT_I8><<<grid, 256, 0, g->stream>>>(s->rows, s->local_n, s->pitch, s->norm, s->sq, nullptr); break;
        default: k::row_norms<k::DT_U8><<<grid, 256, 0, g->stream>>>(s->rows, s->local_n, s->pitch, s->norm, s->sq, nullptr); break;
    }
    CUDA_CHECK(cudaGetLastError());
    g->all_launches++;
    unsigned long long host_nf = 0;
    CUDA_CHECK(cudaMemcpyAsync(&host_nf, nf, 8, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    s->nonfinite_rows = host_nf;
    if (s->dtype == k::DT_F32 && !std::getenv("MVF_NO_TF32_PLANES")) {
        // K2' operand planes: twice the shard again.  Kept only while that leaves at least a third of the device's
        // memory free (a space that large stays on the scan kernel); a refused allocation is not an error.
        size_t fr = 0, to = 0;
        CUDA_CHECK(cudaMemGetInfo(&fr, &to));
        const size_t need = 2 * (size_t)s->local_n * s->pitch;
        if (need + to / 3 <= fr && cudaMalloc((void**)&s->rows_split, need) == cudaSuccess) {
            mvf::tc::split_rows_tf32(s->rows, s->local_n, s->pitch, s->rows_split, g->stream);
            g->all_launches++;
            CUDA_CHECK(cudaStreamSynchronize(g->stream));
            s->split_everywhere = true;
        } else {
            cudaGetLastError();
            s->rows_split = nullptr;
        }
    }
}

// first global row of every rank's shard, kept on the device for the cross-rank merge
void set_rank_bases(mvf_gpu_space* s, const std::vector<uint64_t>& bases) {
    CUDA_CHECK(cudaSetDevice(s->g->device));
    if (!s->rank_base) CUDA_CHECK(cudaMalloc((void**)&s->rank_base, bases.size() * 8));
    CUDA_CHECK(cudaMemcpy(s->rank_base, bases.data(), bases.size() * 8, cudaMemcpyHostToDevice));
}

// Makes the path-deciding facts of a sharded space identical on every rank: the count of non-finite rows is summed
// and "tf32 planes resident" is AND-ed over the ranks (one ncclAllGather of two words per rank; collective: every
// rank creates the same spaces in the same order, as it must for searching them anyway).
void publish_facts(mvf_gpu_space* s) {
    mvf_gpu* g = s->g;
    if (!(g->world > 1 && g->comm)) return;
    if (g->parent_group) return;                          // members of a group engine: the group sums on the host
    g->scalar_tmp.reserve(32 + 32 * (size_t)g->world);
    uint64_t mine[4] = {s->nonfinite_rows, s->split_everywhere ? 1ull : 0ull, s->row_base, s->local_n};
    unsigned char* base = (unsigned char*)g->scalar_tmp.p;
    CUDA_CHECK(cudaMemcpyAsync(base, mine, 32, cudaMemcpyHostToDevice, g->stream));
    nccl_check(nccl().AllGather(base, base + 32, 4, kNcclUint64, g->comm, g->stream), "ncclAllGather");
    std::vector<uint64_t> all(4 * (size_t)g->world);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), base + 32, 32 * (size_t)g->world, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    uint64_t nf = 0, smallest = 0; bool split = true;
    std::vector<uint64_t> bases((size_t)g->world);
    for (int r = 0; r < g->world; r++) {
        nf += all[4 * r]; split = split && all[4 * r + 1] != 0; bases[(size_t)r] = all[4 * r + 2];
        const uint64_t rows = all[4 * r + 3];
        if (rows && (!smallest || rows < smallest)) smallest = rows;
    }
    s->nonfinite_rows = nf; s->split_everywhere = split; s->min_shard_rows = smallest;
    set_rank_bases(s, bases);
}

// Where the source rows live in a file, when they are a file mapping (lets the residency path re-map the range).
struct FileRange { int fd = -1; uint64_t offset = 0; };

// Copy `nrows` rows of `rowbytes` from host memory (any alignment, possibly a read-only file
// mapping) into the pitched device buffer.  First choice: page-lock the source range in place with
// cudaHostRegister and DMA straight out of it, window by window (1 GiB), trying in turn
//   "ro"       the reader's own PROT_READ / MAP_SHARED mapping, cudaHostRegisterReadOnly (when the device reports
//              cudaDevAttrHostRegisterReadOnlySupported),
//   "rw"       the same mapping, cudaHostRegisterDefault,
//   "private"  a second, MAP_PRIVATE PROT_READ|PROT_WRITE mapping of the same file range (copy-on-write: the file is
//              never written), cudaHostRegisterDefault -- only when MVF_HOST_REGISTER names it, see DESIGN.md,
// and recording every refusal in upload_note.  If none is accepted the rows stream through pinned staging
// buffers.  All of these are residency paths, not compute fallbacks.  MVF_HOST_REGISTER=ro|rw|private|none forces one.
void upload_rows(mvf_gpu_space* s, const unsigned char* src, uint64_t nrows, size_t rowbytes, bool try_register,
                 FileRange file = FileRange()) {
    mvf_gpu* g = s->g;
    if (nrows == 0) return;
    auto t0 = std::chrono::steady_clock::now();
    const size_t pitch = s->pitch;
    if (pitch != rowbytes) CUDA_CHECK(cudaMemsetAsync(s->rows, 0, (size_t)nrows * pitch, g->stream));
    auto copy_rows = [&](const unsigned char* from, uint64_t r0, uint64_t cnt) {
        unsigned char* dst = s->rows + r0 * pitch;
        if (pitch == rowbytes) CUDA_CHECK(cudaMemcpyAsync(dst, from, cnt * rowbytes, cudaMemcpyHostToDevice, g->stream));
        else CUDA_CHECK(cudaMemcpy2DAsync(dst, pitch, from, rowbytes, rowbytes, cnt, cudaMemcpyHostToDevice, g->stream));
    };
    const uint64_t window_rows = std::max<uint64_t>(1, (1ull << 30) / rowbytes);
    uint64_t r = 0;
    const long page = 4096;
    std::vector<std::string> variants;
    if (try_register) {
        const char* force = std::getenv("MVF_HOST_REGISTER");
        if (force && *force) { if (std::string(force) != "none") variants.push_back(force); }
        else {
            int ro = 0;
            cudaDeviceGetAttribute(&ro, cudaDevAttrHostRegisterReadOnlySupported, g->device);
            if (ro) variants.push_back("ro"); else s->upload_note += "cudaDevAttrHostRegisterReadOnlySupported = 0; ";
            variants.push_back("rw");
        }
    }
    for (const std::string& var : variants) {
        if (r >= nrows) break;
        if (var == "private" && file.fd < 0) { s->upload_note += "private: source is not a file mapping; "; continue; }
        bool ok = true;
        while (ok && r < nrows) {
            const uint64_t cnt = std::min<uint64_t>(window_rows, nrows - r);
            const unsigned char* a = src + r * rowbytes;
            const unsigned char* from = a;
            void* reg_base = nullptr; size_t reg_len = 0;
            void* priv = nullptr; size_t priv_len = 0;
            if (var == "private") {
                const uint64_t fo = file.offset + r * rowbytes, fo_lo = fo & ~(uint64_t)(page - 1);
                priv_len = (size_t)(((fo + cnt * rowbytes + page - 1) & ~(uint64_t)(page - 1)) - fo_lo);
                priv = mmap(nullptr, priv_len, PROT_READ | PROT_WRITE, MAP_PRIVATE, file.fd, (off_t)fo_lo);
                if (priv == MAP_FAILED) { s->upload_note += "private: mmap failed; "; ok = false; break; }
                reg_base = priv; reg_len = priv_len; from = (const unsigned char*)priv + (fo - fo_lo);
            } else {
                const uintptr_t lo = (uintptr_t)a & ~(uintptr_t)(page - 1);
                const uintptr_t hi = ((uintptr_t)a + cnt * rowbytes + page - 1) & ~(uintptr_t)(page - 1);
                reg_base = (void*)lo; reg_len = hi - lo;
            }
            const cudaError_t e = cudaHostRegister(reg_base, reg_len, var == "ro" ? cudaHostRegisterReadOnly : cudaHostRegisterDefault);
            if (e != cudaSuccess) {
                s->upload_note += "cudaHostRegister(" + var + ") refused: " + cudaGetErrorName(e) + "; ";
                cudaGetLastError();
                if (priv) munmap(priv, priv_len);
                ok = false; break;
            }
            try {
                copy_rows(from, r, cnt);
                CUDA_CHECK(cudaStreamSynchronize(g->stream));
            } catch (...) { cudaHostUnregister(reg_base); if (priv) munmap(priv, priv_len); throw; }
            cudaHostUnregister(reg_base);
            if (priv) munmap(priv, priv_len);
            s->upload_path = 1;
            if (s->upload_note.find("accepted: " + var) == std::string::npos) s->upload_note += "accepted: " + var + "; ";
            r += cnt;
        }
    }
    if (r < nrows) {
        // pinned staging, several host threads: each owns two staging buffers and a stream, copies
        // its share of the row chunks out of the mapping and queues the H2D copy behind it
        s->upload_path = 2;
        CUDA_CHECK(cudaStreamSynchronize(g->stream));          // the padding memset above is ordered first
        const size_t stage_bytes = std::max<size_t>(16u << 20, rowbytes);
        const uint64_t stage_rows = stage_bytes / rowbytes;
        const uint64_t first = r, chunks = (nrows - first + stage_rows - 1) / stage_rows;
        const int nthreads = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)g->upload_threads, chunks));
        // the kernel should read ahead: the mapping is consumed front to back exactly once
        {
            const long page = 4096;
            uintptr_t lo = (uintptr_t)(src + first * rowbytes) & ~(uintptr_t)(page - 1);
            madvise((void*)lo, (uintptr_t)(src + nrows * rowbytes) - lo, MADV_SEQUENTIAL);
            madvise((void*)lo, (uintptr_t)(src + nrows * rowbytes) - lo, MADV_WILLNEED);
        }
        const bool pooled = stage_bytes == (16u << 20);
        if (pooled) {
            while (g->stage_pool.size() < (size_t)nthreads * 2) {
                unsigned char* b = nullptr;
                CUDA_CHECK(cudaMallocHost((void**)&b, 16u << 20));
                g->stage_pool.push_back(b);
            }
        }
        std::vector<std::thread> workers;
        std::mutex err_mu;
        std::string err;
        const int device = g->device;
        for (int t = 0; t < nthreads; t++) {
            workers.emplace_back([&, t] {
                unsigned char* stage[2] = {nullptr, nullptr};
                cudaEvent_t ev[2] = {nullptr, nullptr};
                cudaStream_t st = nullptr;
                auto fail = [&](cudaError_t ce, const char* what) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (err.empty()) err = std::string(cudaGetErrorName(ce)) + ": " + cudaGetErrorString(ce) + " at " + what;
                };
                cudaError_t ce = cudaSetDevice(device);
                if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
                for (int i = 0; i < 2 && ce == cudaSuccess; i++) {
                    if (pooled) stage[i] = g->stage_pool[(size_t)t * 2 + i];
                    else ce = cudaMallocHost((void**)&stage[i], stage_rows * rowbytes);
                    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
                }
                if (ce != cudaSuccess) fail(ce, "staging setup");
                int b = 0;
                for (uint64_t c = (uint64_t)t; c < chunks && ce == cudaSuccess; c += (uint64_t)nthreads) {
                    const uint64_t r0 = first + c * stage_rows;
                    const uint64_t cnt = std::min<uint64_t>(stage_rows, nrows - r0);
                    ce = cudaEventSynchronize(ev[b]);
                    if (ce != cudaSuccess) { fail(ce, "cudaEventSynchronize"); break; }
                    std::memcpy(stage[b], src + r0 * rowbytes, cnt * rowbytes);
                    unsigned char* dst = s->rows + r0 * pitch;
                    if (pitch == rowbytes) ce = cudaMemcpyAsync(dst, stage[b], cnt * rowbytes, cudaMemcpyHostToDevice, st);
                    else ce = cudaMemcpy2DAsync(dst, pitch, stage[b], rowbytes, rowbytes, cnt, cudaMemcpyHostToDevice, st);
                    if (ce == cudaSuccess) ce = cudaEventRecord(ev[b], st);
                    if (ce != cudaSuccess) { fail(ce, "staged copy"); break; }
                    b ^= 1;
                }
                if (st) cudaStreamSynchronize(st);
                for (int i = 0; i < 2; i++) { if (!pooled && stage[i]) cudaFreeHost(stage[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
                if (st) cudaStreamDestroy(st);
            });
        }
        for (auto& w : workers) w.join();
        if (!err.empty()) throw MvfError::cuda(err);
    }
    s->upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Row mask: bit i (LSB first) of `bitmap`, indexed by GLOBAL row, set = row i never appears in a result.
// The shard's bits are re-based to local rows on the host and kept as 32-bit words on the device.
// Two components are kept apart on the host -- the space's deletions (part 0) and the caller's filter (part 1) --
// and their union is what the kernels see, so a filter never resurrects a deleted row.
void set_mask_part(mvf_gpu_space* s, int part, const uint8_t* bitmap, uint64_t nbits) {
    mvf_gpu* g = s->g;
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    const size_t words = (size_t)((s->local_n + k::kPadRows + 31) / 32);
    std::vector<uint32_t>& mine = s->mask_part[part];
    if (!bitmap || nbits == 0) {
        mine.clear(); mine.shrink_to_fit();
    } else {
        mine.assign(words, 0);
        auto bit = [&](uint64_t gb) -> uint32_t { return gb < nbits ? (bitmap[gb >> 3] >> (gb & 7)) & 1u : 0u; };
        for (uint64_t w = 0; w * 32 < s->local_n; w++) {          // 32 local rows per word, any bit offset in the global bitmap
            const uint64_t g0 = s->row_base + w * 32;
            const uint64_t lim = std::min<uint64_t>(32, s->local_n - w * 32);
            uint32_t v = 0;
            if (g0 + 40 <= nbits) {                                // whole window inside the bitmap: one unaligned 5-byte read
                uint64_t raw = 0;
                std::memcpy(&raw, bitmap + (g0 >> 3), 5);            // bits g0 .. g0+39 exist, so these 5 bytes do
                v = (uint32_t)(raw >> (g0 & 7));
                if (lim < 32) v &= (1u << lim) - 1u;
            } else {
                for (uint64_t b = 0; b < lim; b++) v |= bit(g0 + b) << b;
            }
            mine[w] = v;
        }
    }
    if (s->acc_seed) { cudaFree(s->acc_seed); s->acc_seed = nullptr; }    // they encode the mask: rebuilt on the next tensor-path search
    const std::vector<uint32_t>& a = s->mask_part[0];
    const std::vector<uint32_t>& b = s->mask_part[1];
    if (a.empty() && b.empty()) { cudaFree(s->mask); s->mask = nullptr; s->masked_rows = 0; return; }
    std::vector<uint32_t> both(words, 0);
    uint64_t masked = 0;
    for (size_t w = 0; w < words; w++) {
        both[w] = (a.empty() ? 0u : a[w]) | (b.empty() ? 0u : b[w]);
        masked += (uint64_t)__builtin_popcount(both[w]);
    }
    if (!s->mask) CUDA_CHECK(cudaMalloc((void**)&s->mask, words * 4));
    CUDA_CHECK(cudaMemcpy(s->mask, both.data(), words * 4, cudaMemcpyHostToDevice));
    s->masked_rows = masked;
}
void set_row_mask(mvf_gpu_space* s, const uint8_t* bitmap, uint64_t nbits) { set_mask_part(s, 1, bitmap, nbits); }

// External ids, one per GLOBAL row, replicated on every rank: the merged candidates carry global positions.
void set_ids(mvf_gpu_space* s, const void* ids, uint64_t count) {
    mvf_gpu* g = s->g;
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    if (!ids || count == 0) { cudaFree(s->ext_ids); s->ext_ids = nullptr; return; }
    if (count < s->global_n) throw MvfError::dimension_mismatch(s->global_n, count);
    if (!s->ext_ids) CUDA_CHECK(cudaMalloc((void**)&s->ext_ids, (size_t)s->global_n * 8));
    CUDA_CHECK(cudaMemcpy(s->ext_ids, ids, (size_t)s->global_n * 8, cudaMemcpyHostToDevice));
}

void free_space(mvf_gpu_space* s) {
    if (!s) return;
    for (mvf_gpu_space* sh : s->shards) free_space(sh);
    if (!s->is_group()) {
        cudaSetDevice(s->g->device);
        cudaFree(s->rows); cudaFree(s->norm); cudaFree(s->sq); cudaFree(s->mask); cudaFree(s->ext_ids); cudaFree(s->rows_split); cudaFree(s->acc_seed); cudaFree(s->rank_base);
    }
    delete s;
}

// ---- engine life cycle ----------------------------------------------------------------------------
mvf_gpu* create_engine(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw MvfError::cuda("no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= count) throw MvfError::cuda("invalid device ordinal " + std::to_string(device));
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        throw MvfError::cuda("device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                             "; this library carries sm_100a code only");
    mvf_gpu* g = new mvf_gpu();
    g->device = device; g->sm_count = prop.multiProcessorCount; g->cc_major = prop.major; g->cc_minor = prop.minor;
    g->smem_optin = prop.sharedMemPerBlockOptin;
    cudaError_t se = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete g; CUDA_CHECK(se); }
    g->own_stream = true;
    if (const char* c = std::getenv("MVF_SCAN_CTAS_PER_SM")) g->ctas_per_sm_cap = std::atoi(c);
    if (std::getenv("MVF_NO_FUSED_MERGE")) g->no_fused_merge = true;
    if (const char* c = std::getenv("MVF_SYNC_MODE")) g->block_sync = std::string(c) == "block";
    // staging threads: measured on the 16-core hosts of this pool 8 threads move 34 GB/s, 12 move 45 GB/s; leave a few cores alone
    g->upload_threads = (int)std::max(4u, std::min(12u, std::thread::hardware_concurrency() > 4 ? std::thread::hardware_concurrency() - 4 : 4u));
    if (const char* c = std::getenv("MVF_UPLOAD_THREADS")) g->upload_threads = std::max(1, std::min(16, std::atoi(c)));
    try { mvf::tc::init(g->tc); } catch (...) { /* the tensor path reports its own error when it is first used */ }
    return g;
}

void destroy_engine(mvf_gpu* g) {
    if (!g) return;
    for (mvf_gpu* m : g->members) destroy_engine(m);
    if (!g->is_group()) {
        cudaSetDevice(g->device);
        cudaStreamSynchronize(g->stream);
        if (g->comm) { try { nccl().CommDestroy(g->comm); } catch (...) {} }
        for (auto& e : g->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        g->cand.release(); g->packed_local.release(); g->gathered.release(); g->queries.release();
        g->out_ids.release(); g->out_scores.release(); g->out_iscores.release(); g->merge_tmp[0].release(); g->merge_tmp[1].release();
        g->gather_ids.release(); g->gather_out.release(); g->crc_tables.release(); g->scalar_tmp.release(); g->scan_counter.release();
        mvf::tc::release(g->tc);
        peer_exchange_release(g);
        if (g->pinned_out) cudaFreeHost(g->pinned_out);
        if (g->pinned_flag) cudaFreeHost(g->pinned_flag);
        if (g->xdev_event) cudaEventDestroy(g->xdev_event);
        for (auto* b : g->stage_pool) cudaFreeHost(b);
        if (g->own_stream) cudaStreamDestroy(g->stream);
    }
    delete g;
}

// ---- building resident shards (one engine = one shard) ----------------------------------------------
// the shard [begin, begin+count) of a file's space on engine g; `n` = rows the reference scan visits
mvf_gpu_space* shard_from_file(mvf_gpu* g, const VectorSpace& vs, uint64_t n, uint64_t begin, uint64_t count) {
    const int dtype = (int)vs.data_type();
    const size_t es = element_size(vs.data_type());
    mvf_gpu_space* s = alloc_space(g, dtype, (int)vs.distance_metric(), count, vs.dimension(), begin, n);
    try {
        s->ranged = true;
        if (count) {
            // map_vector_range semantics (vector_space.rs:155-188) without the total_vectors cap
            // (scannable_vectors already applied both limits)
            size_t nbytes = 0;
            const uint8_t* base = vs.map_vector_range(begin, count, &nbytes);
            FileRange fr;
            fr.fd = vs.reader()->fd(); fr.offset = (uint64_t)(base - vs.reader()->data());
            upload_rows(s, base, count, (size_t)vs.dimension() * es, !std::getenv("MVF_NO_HOST_REGISTER"), fr);
        }
        compute_norms(s);
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        if (count) {
            // TombstoneInfo{Bitmap} in the footer (schema/core.fbs:34-39): deleted rows never compete
            size_t tb = 0;
            if (const uint8_t* tomb = vs.tombstone_bitmap(&tb)) {
                set_mask_part(s, 0, tomb, (uint64_t)tb * 8);
            } else if (const uint8_t* lst = vs.tombstone_list(&tb)) {   // TombstoneInfo{SortedList}: positions -> bitmap
                std::vector<uint8_t> bm((size_t)((n + 7) / 8), 0);
                for (size_t i = 0; i < tb; i++) {
                    uint64_t r; std::memcpy(&r, lst + 8 * i, 8);
                    if (r < n) bm[r >> 3] |= (uint8_t)(1u << (r & 7));
                }
                set_mask_part(s, 0, bm.data(), n);
            }
        }
        // external ids (vector_ids_block_index != 0): reported instead of positions
        size_t nid = 0;
        if (const uint8_t* ids = vs.vector_ids(&nid)) set_ids(s, ids, (uint64_t)nid);
        publish_facts(s);
    } catch (...) { free_space(s); throw; }
    return s;
}

mvf_gpu_space* shard_synthetic(mvf_gpu* g, int data_type, int metric, uint64_t global_n, uint32_t dimension, uint64_t seed,
                               uint64_t begin, uint64_t count) {
    mvf_gpu_space* s = alloc_space(g, data_type, metric, count, dimension, begin, global_n);
    try {
        s->ranged = true;
        if (count) {
            const uint64_t sk = synth::sm64(seed);
            const uint32_t grid = (uint32_t)g->sm_count * 16;
            switch (data_type) {
                case k::DT_F32: k::synth_fill<k::DT_F32><<<grid, 256, 0, g->stream>>>(s->rows, count, dimension, s->pitch, sk, begin); break;
                case k::DT_F16: k::synth_fill<k::DT_F16><<<grid, 256, 0, g->stream>>>(s->rows, count, dimension, s->pitch, sk, begin); break;
                case k::DT_I8: k::synth_fill<k::DT_I8><<<grid, 256, 0, g->stream>>>(s->rows, count, dimension, s->pitch, sk, begin); break;
                default: k::synth_fill<k::DT_U8><<<grid, 256, 0, g->stream>>>(s->rows, count, dimension, s->pitch, sk, begin); break;
            }
            CUDA_CHECK(cudaGetLastError());
            g->all_launches++;
        }
        compute_norms(s);
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        publish_facts(s);
    } catch (...) { free_space(s); throw; }
    return s;
}

// a group space: the members' shards behind one handle, with the facts the path choice needs summed over them
mvf_gpu_space* wrap_group_space(mvf_gpu* grp, std::vector<mvf_gpu_space*>& shards) {
    mvf_gpu_space* S = new mvf_gpu_space();
    S->g = grp; S->shards = shards;
    const mvf_gpu_space* s0 = shards[0];
    S->dtype = s0->dtype; S->metric = s0->metric; S->d = s0->d; S->pitch = s0->pitch; S->global_n = s0->global_n;
    S->row_base = s0->row_base; S->local_n = 0;
    uint64_t nf = 0; bool split = true, ranged = true;
    for (mvf_gpu_space* sh : shards) { S->local_n += sh->local_n; nf += sh->nonfinite_rows; split = split && sh->split_everywhere; ranged = ranged && sh->ranged; }
    std::vector<uint64_t> bases;
    uint64_t smallest = 0;
    for (mvf_gpu_space* sh : shards) { bases.push_back(sh->row_base); if (sh->local_n && (!smallest || sh->local_n < smallest)) smallest = sh->local_n; }
    for (mvf_gpu_space* sh : shards) { sh->nonfinite_rows = nf; sh->split_everywhere = split; sh->min_shard_rows = smallest; set_rank_bases(sh, bases); }
    S->nonfinite_rows = nf; S->split_everywhere = split; S->ranged = ranged;
    return S;
}

// ---- CRC-32 of resident rows --------------------------------------------------------------------------
uint32_t gf2_mul_host(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) { if (a & (1u << i)) p ^= b; b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1; }
    return p;
}
// crc(A || B) = shift(crc(A), len(B)) ^ crc(B)  (the zlib crc32_combine identity the kernel also uses per thread)
uint32_t crc32_shift(uint32_t crc, uint64_t nbytes_after) {
    uint32_t x = 0x40000000u, w = 0x80000000u;               // x^1, x^0 (reflected)
    for (uint64_t e = 8 * nbytes_after; e; e >>= 1) { if (e & 1) w = gf2_mul_host(x, w); x = gf2_mul_host(x, x); }
    return gf2_mul_host(w, crc);
}
// CRC-32/IEEE of ONE resident shard's row bytes (padding of a ragged pitch excluded), on the device
void shard_checksum(mvf_gpu_space* s, uint32_t* out_crc, double* out_ms) {
    mvf_gpu* g = s->g;
    const size_t rowbytes = (size_t)s->d * element_size((DataType)s->dtype);
    CUDA_CHECK(cudaSetDevice(g->device));
    if (!g->crc_tables.p) {
        // slicing-by-8 tables + x^(2^k) mod P, built once on the host
        std::vector<uint32_t> tab(8 * 256 + 64);
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int b = 0; b < 8; b++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tab[i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int t = 1; t < 8; t++) tab[t * 256 + i] = (tab[(t - 1) * 256 + i] >> 8) ^ tab[tab[(t - 1) * 256 + i] & 0xff];
        uint32_t x = 0x40000000u;                        // x^1
        for (int kb = 0; kb < 64; kb++) { tab[8 * 256 + kb] = x; x = gf2_mul_host(x, x); }
        g->crc_tables.reserve(tab.size() * 4 + 16);
        CUDA_CHECK(cudaMemcpyAsync(g->crc_tables.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, g->stream));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
    }
    k::CrcParams p;
    p.data = s->rows; p.nbytes = (uint64_t)s->local_n * rowbytes;
    if (rowbytes == s->pitch) {
        const uint64_t threads_target = (uint64_t)g->sm_count * 512;
        uint64_t seg = (p.nbytes + threads_target - 1) / threads_target;
        seg = std::max<uint64_t>(256, (seg + 127) & ~127ull);
        p.seg_len = seg; p.seg_stride = seg;
    } else {                                             // ragged pitch: one segment per row, the padding is skipped
        p.seg_len = rowbytes; p.seg_stride = s->pitch;
    }
    p.tables = (const uint32_t*)g->crc_tables.p; p.x2n = p.tables + 8 * 256;
    p.out = (uint32_t*)((unsigned char*)g->crc_tables.p + (8 * 256 + 64) * 4);
    CUDA_CHECK(cudaMemsetAsync(p.out, 0, 4, g->stream));
    uint32_t crc = 0;
    float ms = 0.f;
    if (p.nbytes) {
        const uint64_t nthreads = (p.nbytes + p.seg_len - 1) / p.seg_len;
        const uint32_t grid = (uint32_t)std::max<uint64_t>(1, (nthreads + 255) / 256);
        cudaEvent_t a, b; CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
        cudaEventRecord(a, g->stream);
        k::crc32_resident<<<grid, 256, 0, g->stream>>>(p);
        cudaEventRecord(b, g->stream);
        CUDA_CHECK(cudaGetLastError());
        g->all_launches++;
        CUDA_CHECK(cudaMemcpyAsync(&crc, p.out, 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b);
    }
    *out_crc = crc;
    if (out_ms) *out_ms = ms;
}

void shard_get_vectors(mvf_gpu_space* s, const uint64_t* ids, uint32_t count, int as_f32, void* out) {
    mvf_gpu* g = s->g;
    CUDA_CHECK(cudaSetDevice(g->device));
    const size_t obytes = (size_t)count * s->d * (as_f32 ? 4 : element_size((DataType)s->dtype));
    g->gather_ids.reserve((size_t)count * 8);
    g->gather_out.reserve(obytes);
    CUDA_CHECK(cudaMemcpyAsync(g->gather_ids.p, ids, (size_t)count * 8, cudaMemcpyHostToDevice, g->stream));
    const uint32_t grid = std::min<uint32_t>((count + 7) / 8, (uint32_t)g->sm_count * 8);
    const uint64_t* di = (const uint64_t*)g->gather_ids.p;
    unsigned char* dout = (unsigned char*)g->gather_out.p;
#define MVF_GATHER(DTV, F32V) k::gather_rows<DTV, F32V><<<grid, 256, 0, g->stream>>>(s->rows, s->pitch, s->d, s->local_n, s->row_base, di, count, dout)
    switch (s->dtype) {
        case k::DT_F32: if (as_f32) MVF_GATHER(k::DT_F32, true); else MVF_GATHER(k::DT_F32, false); break;
        case k::DT_F16: if (as_f32) MVF_GATHER(k::DT_F16, true); else MVF_GATHER(k::DT_F16, false); break;
        case k::DT_I8: MVF_GATHER(k::DT_I8, false); break;
        default: MVF_GATHER(k::DT_U8, false); break;
    }
#undef MVF_GATHER
    CUDA_CHECK(cudaGetLastError());
    g->all_launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, dout, obytes, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
}

// ---- searches ---------------------------------------------------------------------------------------
// Host-buffer search of a single-device space: H2D queries, the three phases, D2H results, one stream sync.
void search_host(mvf_gpu_space* s, const void* queries, int query_dtype, uint32_t nq, uint32_t k, int metric,
                 uint64_t* out_ids, float* out_scores, int32_t* out_iscores) {
    mvf_gpu* g = s->g;
    CUDA_CHECK(cudaSetDevice(g->device));
    const size_t qbytes = (size_t)nq * s->d * element_size((DataType)query_dtype);
    const size_t cnt = (size_t)nq * k;
    g->queries.reserve(std::max<size_t>(qbytes, 16));
    // float queries of a batch headed for the tensor path are checked for Inf/NaN here, on the host copy
    int qfinite = QF_UNKNOWN;
    {
        const int m = metric == MVF_METRIC_SPACE_DEFAULT ? s->metric : metric;
        if ((query_dtype == k::DT_F32 || query_dtype == k::DT_F16) && m >= 0 && m <= 2 && k >= 1 && k <= 1024 &&
            tensor_candidate(s, query_dtype, nq, k, m))
            qfinite = host_queries_finite(queries, query_dtype, (size_t)nq * s->d) ? QF_FINITE : QF_NONFINITE;
    }
    // results live in one device block [ids | scores | iscores]; small results come back in a single
    // D2H copy through a pinned mirror (one driver call instead of three), large ones directly
    const size_t back = cnt * (out_iscores ? 16 : 12);
    CUDA_CHECK(cudaMemcpyAsync(g->queries.p, queries, qbytes, cudaMemcpyHostToDevice, g->stream));
    if (back <= 64 * 1024) {
        // small results: the merge kernel stores straight into mapped pinned host memory, so the only
        // host<->device traffic calls are the query copy and the stream sync
        if (!g->pinned_out) {
            CUDA_CHECK(cudaHostAlloc((void**)&g->pinned_out, 64 * 1024, cudaHostAllocMapped));
            CUDA_CHECK(cudaHostGetDevicePointer((void**)&g->pinned_out_dev, g->pinned_out, 0));
        }
        unsigned char* dev = g->pinned_out_dev;
        search_device(s, g->queries.p, query_dtype, nq, k, metric, qfinite, (uint64_t*)dev, (float*)(dev + cnt * 8),
                      out_iscores ? (int32_t*)(dev + cnt * 12) : nullptr);
        wait_stream(g);
        peer_exchange_check(g);
        std::memcpy(out_ids, g->pinned_out, cnt * 8);
        if (out_scores) std::memcpy(out_scores, g->pinned_out + cnt * 8, cnt * 4);
        if (out_iscores) std::memcpy(out_iscores, g->pinned_out + cnt * 12, cnt * 4);
    } else {
        g->out_ids.reserve(cnt * 16);
        unsigned char* dev = (unsigned char*)g->out_ids.p;
        uint64_t* d_ids = (uint64_t*)dev;
        float* d_scores = (float*)(dev + cnt * 8);
        int32_t* d_isc = (int32_t*)(dev + cnt * 12);
        search_device(s, g->queries.p, query_dtype, nq, k, metric, qfinite, d_ids, d_scores, out_iscores ? d_isc : nullptr);
        CUDA_CHECK(cudaMemcpyAsync(out_ids, d_ids, cnt * 8, cudaMemcpyDeviceToHost, g->stream));
        if (out_scores) CUDA_CHECK(cudaMemcpyAsync(out_scores, d_scores, cnt * 4, cudaMemcpyDeviceToHost, g->stream));
        if (out_iscores) CUDA_CHECK(cudaMemcpyAsync(out_iscores, d_isc, cnt * 4, cudaMemcpyDeviceToHost, g->stream));
        wait_stream(g);
        peer_exchange_check(g);
    }
}

// Search of a group space (one process, G devices).  `queries` is host memory (queries_on_device = false) or memory
// of member 0's device; outputs likewise.  Phase 1 is queued device by device (each starts as soon as it is
// launched), the exchange is one ncclGroup of G all-gathers, and only member 0 runs the final merge: every member
// would produce the same list, one copy is wanted.
void search_group(mvf_gpu_space* S, const void* queries, bool on_device, int query_dtype, uint32_t nq, uint32_t k, int metric,
                  uint64_t* out_ids, float* out_scores, int32_t* out_iscores) {
    mvf_gpu* grp = S->g;
    const size_t G = S->shards.size();
    if (metric == MVF_METRIC_SPACE_DEFAULT) metric = S->metric;
    validate_query(S->shards[0], query_dtype, nq, k, metric);
    const size_t qbytes = (size_t)nq * S->d * element_size((DataType)query_dtype);
    const size_t cnt = (size_t)nq * k;
    mvf_gpu* g0 = grp->members[0];
    int qfinite = QF_FINITE;
    if ((query_dtype == k::DT_F32 || query_dtype == k::DT_F16) && tensor_candidate(S->shards[0], query_dtype, nq, k, metric)) {
        if (on_device) { CUDA_CHECK(cudaSetDevice(g0->device)); qfinite = device_queries_finite(g0, queries, query_dtype, (size_t)nq * S->d); }
        else qfinite = host_queries_finite(queries, query_dtype, (size_t)nq * S->d) ? QF_FINITE : QF_NONFINITE;
    }
    if (on_device) {                                   // members wait for whatever produced the queries on member 0's stream
        CUDA_CHECK(cudaSetDevice(g0->device));
        if (!g0->xdev_event) CUDA_CHECK(cudaEventCreateWithFlags(&g0->xdev_event, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(g0->xdev_event, g0->stream));
    }
    // where member 0 leaves the final result: the caller's device buffers, or a device block copied back below
    uint64_t* f_ids = out_ids; float* f_scores = out_scores; int32_t* f_isc = out_iscores;
    if (!on_device) {
        CUDA_CHECK(cudaSetDevice(g0->device));
        g0->out_ids.reserve(cnt * 16);
        unsigned char* dev = (unsigned char*)g0->out_ids.p;
        f_ids = (uint64_t*)dev; f_scores = (float*)(dev + cnt * 8); f_isc = out_iscores ? (int32_t*)(dev + cnt * 12) : nullptr;
    }
    std::vector<SearchCtx> ctx(G);
    for (size_t i = 0; i < G; i++) {
        mvf_gpu_space* sh = S->shards[i];
        mvf_gpu* g = sh->g;
        CUDA_CHECK(cudaSetDevice(g->device));
        const void* dq = queries;
        if (!on_device || i > 0) {
            g->queries.reserve(std::max<size_t>(qbytes, 16));
            if (on_device) {
                CUDA_CHECK(cudaStreamWaitEvent(g->stream, g0->xdev_event, 0));
                CUDA_CHECK(cudaMemcpyPeerAsync(g->queries.p, g->device, queries, g0->device, qbytes, g->stream));
                if (!g->xdev_event) CUDA_CHECK(cudaEventCreateWithFlags(&g->xdev_event, cudaEventDisableTiming));
                CUDA_CHECK(cudaEventRecord(g->xdev_event, g->stream));      // member 0's stream will wait: the caller may reuse the buffer
            } else {
                CUDA_CHECK(cudaMemcpyAsync(g->queries.p, queries, qbytes, cudaMemcpyHostToDevice, g->stream));
            }
            dq = g->queries.p;
        }
        // (small payloads: the exchange and member 0's final merge may already ride on this launch, see search_local)
        if (i == 0) search_local(sh, dq, query_dtype, nq, k, metric, qfinite, f_ids, f_scores, f_isc, &ctx[i], true);
        else search_local(sh, dq, query_dtype, nq, k, metric, qfinite, nullptr, nullptr, nullptr, &ctx[i], false);
    }
    // small payloads: every member pushes its candidates into every member's window (peer stores) and member 0 merges
    // what arrived in its own -- no NCCL; the other members wait for the same cells so that nobody runs two searches
    // ahead of a reader (kernels.cuh exchange_merge).  Large payloads: one ncclGroup of G all-gathers, merge on member 0.
    if (ctx[0].peer) {
        for (size_t i = 1; i < G; i++) search_exchange_peer(S->shards[i], ctx[i], nullptr, nullptr, nullptr);
        search_exchange_peer(S->shards[0], ctx[0], f_ids, f_scores, f_isc);
    } else {
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        for (size_t i = 0; i < G; i++) search_exchange(S->shards[i], ctx[i]);
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
        search_final(S->shards[0], ctx[0], f_ids, f_scores, f_isc);
    }
    CUDA_CHECK(cudaSetDevice(g0->device));
    if (on_device) {
        for (size_t i = 1; i < G; i++) CUDA_CHECK(cudaStreamWaitEvent(g0->stream, S->shards[i]->g->xdev_event, 0));
        return;                                                             // asynchronous on member 0's stream
    }
    CUDA_CHECK(cudaMemcpyAsync(out_ids, f_ids, cnt * 8, cudaMemcpyDeviceToHost, g0->stream));
    if (out_scores) CUDA_CHECK(cudaMemcpyAsync(out_scores, f_scores, cnt * 4, cudaMemcpyDeviceToHost, g0->stream));
    if (out_iscores) CUDA_CHECK(cudaMemcpyAsync(out_iscores, f_isc, cnt * 4, cudaMemcpyDeviceToHost, g0->stream));
    wait_stream(g0);
    peer_exchange_check(g0);
}

void require_single(const mvf_gpu* g, const char* what) {
    if (g->is_group()) throw MvfError::build_error(std::string(what) + " applies to single-device engines (this one spans " +
                                                   std::to_string(g->members.size()) + " devices)");
}

}  // namespace

// =================================================================================================
extern "C" {

int mvf_gpu_init(int device, mvf_gpu** out) {
    return guard([&] {
        *out = nullptr;
        *out = create_engine(device);
    });
}

// One process, n devices: the drop-in for the single call site of find_top_k_similar
// (examples/similarity_search.rs:117) that wants all GPUs of the box.
int mvf_gpu_init_multi(const int* devices, int n, mvf_gpu** out) {
    return guard([&] {
        *out = nullptr;
        if (!devices || n < 1) throw MvfError::build_error("mvf_gpu_init_multi needs at least one device");
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++)
                if (devices[i] == devices[j]) throw MvfError::build_error("mvf_gpu_init_multi: device listed twice");
        if (n == 1) { *out = create_engine(devices[0]); return; }
        std::unique_ptr<mvf_gpu, void (*)(mvf_gpu*)> grp(new mvf_gpu(), destroy_engine);
        for (int i = 0; i < n; i++) {
            mvf_gpu* m = create_engine(devices[i]);
            m->rank = i; m->world = n; m->parent_group = grp.get();
            grp->members.push_back(m);
        }
        grp->device = devices[0]; grp->world = n;
        grp->sm_count = grp->members[0]->sm_count; grp->cc_major = grp->members[0]->cc_major; grp->cc_minor = grp->members[0]->cc_minor;
        std::vector<NcclComm> comms((size_t)n, nullptr);
        nccl_check(nccl().CommInitAll(comms.data(), n, devices), "ncclCommInitAll");
        for (int i = 0; i < n; i++) grp->members[(size_t)i]->comm = comms[(size_t)i];
        peer_exchange_setup_group(grp.get());
        *out = grp.release();
    });
}

void mvf_gpu_shutdown(mvf_gpu* g) { destroy_engine(g); }

int mvf_gpu_set_stream(mvf_gpu* g, void* cuda_stream) {
    return guard([&] { EngineLock engine_lock(g->mu);
        require_single(g, "mvf_gpu_set_stream");
        CUDA_CHECK(cudaSetDevice(g->device));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        if (g->own_stream) { cudaStreamDestroy(g->stream); g->own_stream = false; }
        g->stream = (cudaStream_t)cuda_stream;
    });
}

int mvf_gpu_device_count(const mvf_gpu* g, int* out) {
    return guard([&] { *out = g->is_group() ? (int)g->members.size() : 1; });
}

int mvf_gpu_device_info(const mvf_gpu* g, int* sm_count, int* cc_major, int* cc_minor, uint64_t* free_bytes,
                        uint64_t* total_bytes) {
    return guard([&] {
        // a group reports the shape of its first device and the memory of all of them
        uint64_t fsum = 0, tsum = 0;
        auto one = [&](const mvf_gpu* e) {
            CUDA_CHECK(cudaSetDevice(e->device));
            size_t f = 0, t = 0;
            CUDA_CHECK(cudaMemGetInfo(&f, &t));
            fsum += f; tsum += t;
        };
        if (g->is_group()) for (const mvf_gpu* m : g->members) one(m); else one(g);
        if (sm_count) *sm_count = g->sm_count;
        if (cc_major) *cc_major = g->cc_major;
        if (cc_minor) *cc_minor = g->cc_minor;
        if (free_bytes) *free_bytes = fsum;
        if (total_bytes) *total_bytes = tsum;
    });
}

int mvf_gpu_comm_unique_id(void* out_128_bytes) {
    return guard([&] {
        NcclUniqueId id;
        nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(out_128_bytes, &id, 128);
    });
}

int mvf_gpu_comm_init(mvf_gpu* g, const void* unique_id_128_bytes, int rank, int world) {
    return guard([&] { EngineLock engine_lock(g->mu);
        require_single(g, "mvf_gpu_comm_init");
        if (world < 1 || rank < 0 || rank >= world) throw MvfError::nccl("invalid rank/world");
        if (g->comm) throw MvfError::nccl("this engine already has a communicator");
        CUDA_CHECK(cudaSetDevice(g->device));
        g->rank = rank; g->world = world;
        if (world == 1) return;
        NcclUniqueId id; std::memcpy(&id, unique_id_128_bytes, 128);
        nccl_check(nccl().CommInitRank(&g->comm, world, id, rank), "ncclCommInitRank");
        peer_exchange_setup_ipc(g);
    });
}

int mvf_gpu_space_upload(mvf_gpu* g, const mvf_space* sp, mvf_gpu_space** out) {
    return guard([&] { EngineLock engine_lock(g->mu);
        *out = nullptr;
        const VectorSpace& vs = sp->vs;
        if (vs.vector_type() != VectorType::Dense)
            throw MvfError(ErrorKind::InvalidVectorType, "Invalid vector type: expected Dense, got Sparse", 0 /* Dense */, 1 /* Sparse */);
        if (!element_size(vs.data_type())) throw MvfError::build_error("Unsupported vector data type");
        // exactly the rows `for i in 0..total_vectors()` + get_vector can visit (similarity_search.rs:147-148)
        const uint64_t n = vs.scannable_vectors();
        if (g->is_group()) {
            std::vector<mvf_gpu_space*> shards;
            try {
                for (mvf_gpu* m : g->members) {
                    uint64_t begin = 0, count = n;
                    shard_range(n, m->world, m->rank, &begin, &count);
                    shards.push_back(shard_from_file(m, vs, n, begin, count));
                }
            } catch (...) { for (auto* sh : shards) free_space(sh); throw; }
            *out = wrap_group_space(g, shards);
            return;
        }
        uint64_t begin = 0, count = n;
        shard_range(n, g->world, g->rank, &begin, &count);
        *out = shard_from_file(g, vs, n, begin, count);
    });
}

int mvf_gpu_space_from_host(mvf_gpu* g, const void* rows, int data_type, int metric, uint64_t local_n, uint32_t dimension,
                            uint64_t row_base, uint64_t global_n, mvf_gpu_space** out) {
    return guard([&] { EngineLock engine_lock(g->mu);
        *out = nullptr;
        if (row_base + local_n > global_n) throw MvfError::index_out_of_bounds(row_base + local_n, global_n);
        if (data_type < 0 || data_type > 3) throw MvfError::build_error("Unsupported vector data type");
        const size_t rowbytes = (size_t)dimension * element_size((DataType)data_type);
        auto one = [&](mvf_gpu* e, uint64_t begin, uint64_t count, bool ranged) {
            mvf_gpu_space* s = alloc_space(e, data_type, metric, count, dimension, row_base + begin, global_n);
            try {
                s->ranged = ranged;
                upload_rows(s, (const unsigned char*)rows + begin * rowbytes, count, rowbytes, false);
                compute_norms(s);
                CUDA_CHECK(cudaStreamSynchronize(e->stream));
                publish_facts(s);                                // with a communicator: collective, like every space creation
            } catch (...) { free_space(s); throw; }
            return s;
        };
        if (g->is_group()) {
            // the caller's rows are cut once more, contiguously, over the members
            const bool whole = row_base == 0 && local_n == global_n;
            std::vector<mvf_gpu_space*> shards;
            try {
                for (mvf_gpu* m : g->members) {
                    uint64_t begin = 0, count = local_n;
                    shard_range(local_n, m->world, m->rank, &begin, &count);
                    shards.push_back(one(m, begin, count, whole));
                }
            } catch (...) { for (auto* sh : shards) free_space(sh); throw; }
            *out = wrap_group_space(g, shards);
            return;
        }
        *out = one(g, 0, local_n, false);
    });
}

int mvf_gpu_space_synthetic(mvf_gpu* g, int data_type, int metric, uint64_t global_n, uint32_t dimension, uint64_t seed,
                            mvf_gpu_space** out) {
    return guard([&] { EngineLock engine_lock(g->mu);
        *out = nullptr;
        if (g->is_group()) {
            std::vector<mvf_gpu_space*> shards;
            try {
                for (mvf_gpu* m : g->members) {
                    uint64_t begin = 0, count = global_n;
                    shard_range(global_n, m->world, m->rank, &begin, &count);
                    shards.push_back(shard_synthetic(m, data_type, metric, global_n, dimension, seed, begin, count));
                }
            } catch (...) { for (auto* sh : shards) free_space(sh); throw; }
            *out = wrap_group_space(g, shards);
            return;
        }
        uint64_t begin = 0, count = global_n;
        shard_range(global_n, g->world, g->rank, &begin, &count);
        *out = shard_synthetic(g, data_type, metric, global_n, dimension, seed, begin, count);
    });
}

void mvf_gpu_space_free(mvf_gpu_space* s) {
    if (!s) return;
    EngineLock engine_lock(s->g->mu);
    free_space(s);
}

int mvf_gpu_space_set_tombstones(mvf_gpu_space* s, const uint8_t* bitmap, uint64_t nbits) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (s->is_group()) { for (auto* sh : s->shards) set_mask_part(sh, 0, bitmap, nbits); return; }
        set_mask_part(s, 0, bitmap, nbits); });
}
int mvf_gpu_space_set_ids(mvf_gpu_space* s, const uint64_t* ids, uint64_t count) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (s->is_group()) { set_ids(s->shards[0], ids, count); return; }     // only member 0 emits final results
        set_ids(s, ids, count); });
}
int mvf_gpu_space_set_row_mask(mvf_gpu_space* s, const uint8_t* bitmap, uint64_t nbits) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (s->is_group()) { for (auto* sh : s->shards) set_row_mask(sh, bitmap, nbits); return; }
        set_row_mask(s, bitmap, nbits); });
}

int mvf_gpu_space_info(const mvf_gpu_space* s, uint64_t* local_n, uint64_t* row_base, uint64_t* global_n,
                       uint32_t* dimension, int* data_type, int* metric) {
    return guard([&] {
        if (local_n) *local_n = s->local_n;          // a group space: the rows resident over all its devices
        if (row_base) *row_base = s->row_base;
        if (global_n) *global_n = s->global_n;
        if (dimension) *dimension = s->d;
        if (data_type) *data_type = s->dtype;
        if (metric) *metric = s->metric;
    });
}

// facts the scan-vs-tensor choice depends on besides the shape: rows holding Inf/NaN (all shards) and whether the
// tf32 planes of a Float32 space are resident (on every shard)
int mvf_gpu_space_path_facts(const mvf_gpu_space* s, uint64_t* nonfinite_rows, int* tf32_planes_resident, uint64_t* decision_rows_out) {
    return guard([&] {
        if (nonfinite_rows) *nonfinite_rows = s->nonfinite_rows;
        if (tf32_planes_resident) *tf32_planes_resident = s->split_everywhere ? 1 : 0;
        if (decision_rows_out) *decision_rows_out = decision_rows(s->is_group() ? s->shards[0] : s);
    });
}

int mvf_gpu_space_upload_stats(const mvf_gpu_space* s, int* path, double* seconds, uint64_t* bytes) {
    return guard([&] {
        const mvf_gpu_space* first = s->is_group() ? s->shards[0] : s;
        double sec = s->upload_seconds;
        for (const auto* sh : s->shards) sec += sh->upload_seconds;           // members upload one after the other
        if (path) *path = first->upload_path;
        if (seconds) *seconds = sec;
        if (bytes) *bytes = s->local_n * (uint64_t)s->d * element_size((DataType)s->dtype);
    });
}

const char* mvf_gpu_space_upload_note(const mvf_gpu_space* s) {
    if (!s) return "";
    return s->is_group() ? s->shards[0]->upload_note.c_str() : s->upload_note.c_str();
}

int mvf_gpu_space_read_rows(const mvf_gpu_space* s, uint64_t local_start, uint64_t count, void* out) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (local_start + count > s->local_n) throw MvfError::index_out_of_bounds(local_start + count, s->local_n);
        if (!count) return;
        const size_t rowbytes = (size_t)s->d * element_size((DataType)s->dtype);
        auto one = [&](const mvf_gpu_space* sh, uint64_t start, uint64_t cnt, unsigned char* dst) {
            CUDA_CHECK(cudaSetDevice(sh->g->device));
            CUDA_CHECK(cudaStreamSynchronize(sh->g->stream));
            CUDA_CHECK(cudaMemcpy2D(dst, rowbytes, sh->rows + start * sh->pitch, sh->pitch, rowbytes, cnt, cudaMemcpyDeviceToHost));
        };
        if (!s->is_group()) { one(s, local_start, count, (unsigned char*)out); return; }
        uint64_t base = 0;                                                      // rows of the group, shard after shard
        for (const auto* sh : s->shards) {
            const uint64_t lo = std::max(local_start, base), hi = std::min(local_start + count, base + sh->local_n);
            if (lo < hi) one(sh, lo - base, hi - lo, (unsigned char*)out + (lo - local_start) * rowbytes);
            base += sh->local_n;
        }
    });
}

int mvf_gpu_space_read_norms(const mvf_gpu_space* s, float* out_norm, float* out_sq) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        auto one = [&](const mvf_gpu_space* sh, uint64_t off) {
            CUDA_CHECK(cudaSetDevice(sh->g->device));
            CUDA_CHECK(cudaStreamSynchronize(sh->g->stream));
            if (out_norm) CUDA_CHECK(cudaMemcpy(out_norm + off, sh->norm, sh->local_n * 4, cudaMemcpyDeviceToHost));
            if (out_sq) CUDA_CHECK(cudaMemcpy(out_sq + off, sh->sq, sh->local_n * 4, cudaMemcpyDeviceToHost));
        };
        if (!s->is_group()) { one(s, 0); return; }
        uint64_t off = 0;
        for (const auto* sh : s->shards) { if (sh->local_n) one(sh, off); off += sh->local_n; }
    });
}

// CRC-32 of the resident rows of the WHOLE space, computed on the device(s).  Each shard hashes its own rows; shard
// CRCs are moved to their place in the block by the same GF(2) shift the kernel uses per thread and XOR-ed
// (block CRCs are per block, src/builder.rs:251, so a sharded space is verified exactly like a whole one).
// With a communicator this is a collective: every rank calls it and gets the same value.
int mvf_gpu_space_checksum(mvf_gpu_space* s, uint32_t* out_crc, double* out_ms) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        mvf_gpu* g = s->g;
        const uint64_t rowbytes = (uint64_t)s->d * element_size((DataType)s->dtype);
        uint32_t total = 0; double ms = 0.0;
        if (s->is_group()) {
            for (auto* sh : s->shards) {
                uint32_t c = 0; double m = 0.0;
                shard_checksum(sh, &c, &m);
                total ^= crc32_shift(c, (s->global_n - sh->row_base - sh->local_n) * rowbytes);
                ms = std::max(ms, m);
            }
        } else {
            uint32_t mine = 0;
            shard_checksum(s, &mine, &ms);
            if (g->world > 1 && g->comm && s->local_n != s->global_n) {
                g->scalar_tmp.reserve(16 + 8 * (size_t)g->world);
                unsigned char* base = (unsigned char*)g->scalar_tmp.p;
                const uint64_t word = crc32_shift(mine, (s->global_n - s->row_base - s->local_n) * rowbytes);
                CUDA_CHECK(cudaMemcpyAsync(base, &word, 8, cudaMemcpyHostToDevice, g->stream));
                nccl_check(nccl().AllGather(base, base + 16, 1, kNcclUint64, g->comm, g->stream), "ncclAllGather");
                std::vector<uint64_t> all((size_t)g->world);
                CUDA_CHECK(cudaMemcpyAsync(all.data(), base + 16, 8 * (size_t)g->world, cudaMemcpyDeviceToHost, g->stream));
                CUDA_CHECK(cudaStreamSynchronize(g->stream));
                for (uint64_t w : all) total ^= (uint32_t)w;
            } else {
                if (s->local_n != s->global_n) throw MvfError::build_error("this shard holds part of the space and the engine has no communicator to combine the shard checksums");
                total = mine;
            }
        }
        if (out_crc) *out_crc = total;
        if (out_ms) *out_ms = ms;
    });
}

// integrity on upload: device CRC of the resident block against the manifest entry of the file
int mvf_gpu_space_verify_checksum(mvf_gpu_space* s, const mvf_space* sp) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        const VectorSpace& vs = sp->vs;
        const auto& blocks = vs.reader()->blocks();
        const SpaceInfo* info = nullptr;
        for (const auto& si : vs.reader()->spaces()) if (&si.name == &vs.name()) info = &si;
        if (!info || info->vectors_block_index >= blocks.size()) throw MvfError::corrupted_data("Invalid vector block index");
        const DataBlock& blk = blocks[info->vectors_block_index];
        const uint64_t resident = (uint64_t)s->global_n * s->d * element_size((DataType)s->dtype);
        if (blk.checksum == 0) return;                               // nothing recorded (src/reader.rs:183)
        if (resident != blk.size) throw MvfError::build_error("resident rows do not cover the whole block (total_vectors < block rows)");
        uint32_t crc = 0;
        int rc = mvf_gpu_space_checksum(s, &crc, nullptr);
        if (rc != MVF_OK) throw MvfError((ErrorKind)rc, mvf_last_error());
        if (crc != blk.checksum)
            throw MvfError::corrupted_data("Block checksum mismatch: expected " + std::to_string(blk.checksum) + ", got " + std::to_string(crc));
    });
}

int mvf_gpu_space_get_vectors(mvf_gpu_space* s, const uint64_t* ids, uint32_t count, int as_f32, void* out) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (!count) return;
        if (!ids || !out) throw MvfError::build_error("null argument");
        const bool fl = s->dtype == k::DT_F32 || s->dtype == k::DT_F16;
        if (as_f32 && !fl) throw MvfError::build_error("Cannot convert to f32");     // Vector::as_f32, src/vectors/vector.rs:90
        if (!s->is_group()) { shard_get_vectors(s, ids, count, as_f32, out); return; }
        // group: every member gathers the ids of its own row range, rows land at their positions in `out`
        const size_t rb = (size_t)s->d * (as_f32 ? 4 : element_size((DataType)s->dtype));
        std::memset(out, 0, (size_t)count * rb);
        for (auto* sh : s->shards) {
            std::vector<uint64_t> mine; std::vector<uint32_t> at;
            for (uint32_t i = 0; i < count; i++)
                if (ids[i] >= sh->row_base && ids[i] - sh->row_base < sh->local_n) { mine.push_back(ids[i]); at.push_back(i); }
            if (mine.empty()) continue;
            std::vector<unsigned char> tmp(mine.size() * rb);
            shard_get_vectors(sh, mine.data(), (uint32_t)mine.size(), as_f32, tmp.data());
            for (size_t j = 0; j < mine.size(); j++) std::memcpy((unsigned char*)out + (size_t)at[j] * rb, tmp.data() + j * rb, rb);
        }
    });
}

int mvf_gpu_search_device(mvf_gpu_space* s, const void* d_queries, int query_dtype, uint32_t nq, uint32_t k, int metric,
                          uint64_t* d_out_ids, float* d_out_scores, int32_t* d_out_iscores) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (query_dtype < 0 || query_dtype > 3) throw MvfError::build_error("unsupported query data type");
        if (s->is_group()) search_group(s, d_queries, true, query_dtype, nq, k, metric, d_out_ids, d_out_scores, d_out_iscores);
        else search_device(s, d_queries, query_dtype, nq, k, metric, QF_UNKNOWN, d_out_ids, d_out_scores, d_out_iscores);
    });
}

int mvf_gpu_search(mvf_gpu_space* s, const void* queries, int query_dtype, uint32_t nq, uint32_t k, int metric,
                   uint64_t* out_ids, float* out_scores, int32_t* out_iscores) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        if (!queries || !out_ids) throw MvfError::build_error("null argument");
        if (nq == 0) throw MvfError::build_error("search needs at least one query");
        if (query_dtype < 0 || query_dtype > 3) throw MvfError::build_error("unsupported query data type");
        if (s->is_group()) search_group(s, queries, false, query_dtype, nq, k, metric, out_ids, out_scores, out_iscores);
        else search_host(s, queries, query_dtype, nq, k, metric, out_ids, out_scores, out_iscores);
    });
}

int mvf_gpu_tc_dump(mvf_gpu_space* s, const void* queries, int query_dtype, uint32_t nq, uint32_t k, int metric,
                    float* out, uint64_t out_elems, uint32_t* out_rows, uint32_t* out_cols) {
    return guard([&] { EngineLock engine_lock(s->g->mu);
        mvf_gpu* g = s->g;
        require_single(g, "mvf_gpu_tc_dump");
        if (metric == MVF_METRIC_SPACE_DEFAULT) metric = s->metric;
        if (!mvf::tc::eligible(g->tc, 2, s->dtype, query_dtype, metric, nq, k, s->local_n, s->d, s->pitch) ||
            (s->dtype == k::DT_F32 && !s->rows_split))
            throw MvfError::build_error("this space/query/metric combination has no tensor-core path");
        CUDA_CHECK(cudaSetDevice(g->device));
        const size_t qbytes = (size_t)nq * s->d * element_size((DataType)query_dtype);
        g->queries.reserve(std::max<size_t>(qbytes, 16));
        CUDA_CHECK(cudaMemcpyAsync(g->queries.p, queries, qbytes, cudaMemcpyHostToDevice, g->stream));
        if ((s->dtype == k::DT_I8 || s->dtype == k::DT_U8) && metric == k::M_L2 && !s->acc_seed && s->local_n) {
            const uint64_t padded = ((s->local_n + mvf::tc::BN - 1) / mvf::tc::BN) * mvf::tc::BN;
            CUDA_CHECK(cudaMalloc((void**)&s->acc_seed, padded * 4));
            k::build_acc_seed<<<(uint32_t)g->sm_count * 8, 256, 0, g->stream>>>(s->sq, s->mask, s->local_n, padded, (int)0x80000000 + (1 << 25), s->acc_seed);
            CUDA_CHECK(cudaGetLastError());
        }
        mvf::tc::SearchArgs a;
        a.acc_seed = s->acc_seed;
        a.rows = s->rows; a.rows_split = s->rows_split; a.norm = s->norm; a.sq = s->sq; a.mask = s->mask; a.n = s->local_n; a.d = s->d;
        a.pitch = s->pitch; a.dtype = s->dtype; a.metric = metric; a.queries = g->queries.p; a.nq = nq; a.k = k; a.dump = true;
        mvf::tc::SearchResult r = mvf::tc::run(g->tc, a, g->stream, g->sm_count);
        const uint32_t rows = ((nq + mvf::tc::BM - 1) / mvf::tc::BM) * mvf::tc::BM;
        if (out_rows) *out_rows = rows;
        if (out_cols) *out_cols = r.dump_cols;
        if ((uint64_t)rows * r.dump_cols > out_elems) throw MvfError::build_error("dump buffer too small");
        CUDA_CHECK(cudaMemcpyAsync(out, r.dump, (size_t)rows * r.dump_cols * 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
    });
}

// Host-only planning query: the launch shape K2 would use for this problem on a GPU with `sm_count` SMs.
// out[0..15] = mtiles, ntiles, chunks, tiles_per_chunk, cluster, mgroups, items, grid, kblocks, cap, trigger,
// sync_window, epilogue groups, stages, pair form (0/1), dynamic shared memory bytes.  No CUDA call is made.
int mvf_tc_plan(int data_type, uint32_t nq, uint32_t k, uint64_t n, uint32_t dimension, int sm_count, uint32_t* out16) {
    return guard([&] {
        if (!out16 || sm_count < 2 || nq == 0 || k == 0 || k > (uint32_t)mvf::tc::MAX_K || n == 0 || dimension == 0 || data_type < 0 || data_type > 3)
            throw MvfError::build_error("no tensor-core plan for these arguments");
        mvf::tc::TcState st;
        mvf::tc::read_env(st);
        if (!mvf::tc::eligible(st, 2, data_type, data_type, k::M_L2, nq, k, n, dimension, 0))
            throw MvfError::build_error("no tensor-core plan for these arguments");
        const uint32_t es = mvf::tc::elem_bytes(mvf::tc::kind_of(data_type));
        const uint32_t pitch = (dimension * es + 15u) & ~15u;
        const mvf::tc::Plan pl = mvf::tc::make_plan(st, data_type, nq, k, n, dimension, pitch, sm_count);
        const uint32_t v[16] = {pl.mtiles, pl.ntiles, pl.chunks, pl.tiles_per_chunk, pl.CL, pl.mgroups, pl.items, pl.grid, pl.kblocks,
                                pl.cap, pl.trigger, pl.sync_window, (uint32_t)pl.EG, (uint32_t)pl.stages, pl.pair ? 1u : 0u,
                                (uint32_t)pl.smem};
        for (int i = 0; i < 16; i++) out16[i] = v[i];
    });
}

// Host arithmetic only (no GPU needed): the row count every rank of a space sharded by mvf_shard_range(global_n, world)
// takes the scan-vs-tensor decision on -- the smallest non-empty shard -- and whether a batch of nq queries / top-k of
// that dtype and metric would then go to the tensor path in auto mode (default thresholds; non-finite data aside).
int mvf_path_decision(uint64_t global_n, int world, int data_type, int metric, uint32_t nq, uint32_t k, uint32_t dimension,
                      uint64_t* decision_rows_out, int* tensor_path) {
    return guard([&] {
        if (world < 1) throw MvfError::build_error("world must be >= 1");
        const uint64_t rows = world > 1 ? smallest_nonempty_shard(global_n, world) : global_n;
        if (decision_rows_out) *decision_rows_out = rows;
        if (tensor_path) {
            mvf::tc::TcState st;
            mvf::tc::read_env(st);
            *tensor_path = mvf::tc::eligible(st, 0, data_type, data_type, metric, nq, k, rows, dimension, 0) ? 1 : 0;
        }
    });
}

int mvf_gpu_set_path(mvf_gpu* g, int path) {
    return guard([&] { EngineLock engine_lock(g->mu);
        if (path < 0 || path > 2) throw MvfError::build_error("path must be 0 (auto), 1 (scan) or 2 (tensor)");
        g->path = path;
        for (mvf_gpu* m : g->members) m->path = path;
    });
}

int mvf_gpu_profile(mvf_gpu* g, int enable) {
    return guard([&] { EngineLock engine_lock(g->mu);
        auto one = [&](mvf_gpu* e) {
            CUDA_CHECK(cudaSetDevice(e->device));
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            flush_profile(e);
            e->profile = enable != 0;
            e->scan_ms_total = 0.0; e->scan_launches = 0; e->all_launches = 0;
        };
        if (g->is_group()) for (mvf_gpu* m : g->members) one(m); else one(g);
    });
}

// a group reports member 0's dominant-kernel time (the members run concurrently) and the launches of all members
int mvf_gpu_profile_read(mvf_gpu* g, double* scan_ms_total, uint64_t* scan_launches, uint64_t* all_launches, int* last_path) {
    return guard([&] { EngineLock engine_lock(g->mu);
        auto one = [&](mvf_gpu* e) {
            CUDA_CHECK(cudaSetDevice(e->device));
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            flush_profile(e);
        };
        uint64_t all = 0;
        if (g->is_group()) for (mvf_gpu* m : g->members) { one(m); all += m->all_launches; } else { one(g); all = g->all_launches; }
        const mvf_gpu* e0 = g->is_group() ? g->members[0] : g;
        if (scan_ms_total) *scan_ms_total = e0->scan_ms_total;
        if (scan_launches) *scan_launches = e0->scan_launches;
        if (all_launches) *all_launches = all;
        if (last_path) *last_path = e0->last_path;
    });
}

int mvf_gpu_synchronize(mvf_gpu* g) {
    return guard([&] { EngineLock engine_lock(g->mu);
        auto one = [&](mvf_gpu* e) { CUDA_CHECK(cudaSetDevice(e->device)); CUDA_CHECK(cudaStreamSynchronize(e->stream)); peer_exchange_check(e); };
        if (g->is_group()) for (mvf_gpu* m : g->members) one(m); else one(g);
    });
}

// How the exchange step of sharded searches runs on this engine: *peer_ready = 1 when small payloads go through the
// peer-memory kernel (X1), the searches that took each route so far, and (when the peer path is off) why.
int mvf_gpu_exchange_info(const mvf_gpu* g, int* peer_ready, uint64_t* peer_searches, uint64_t* nccl_searches) {
    return guard([&] {
        const mvf_gpu* e = g->is_group() ? g->members[0] : g;
        if (peer_ready) *peer_ready = e->px.ready ? 1 : 0;
        if (peer_searches) *peer_searches = e->peer_searches;
        if (nccl_searches) *nccl_searches = e->nccl_searches;
    });
}
const char* mvf_gpu_exchange_note(const mvf_gpu* g) {
    if (!g) return "";
    const mvf_gpu* e = g->is_group() ? g->members[0] : g;
    return e->px.note.c_str();
}

}  // extern "C"
