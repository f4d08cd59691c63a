// tc_kernels.cuh -- K2, the batched path: queries x database contraction on tcgen05 tensor cores
// with the per-query top-k selection fused into the accumulator epilogue (sm_100a only).
//
// Shape of the work: S = Q (nq x d) . X^T (n x d), both operands K-major exactly as stored
// (row-major queries, row-major resident rows), so no transposes.  One CTA tile = 128 queries x
// 256 database rows; K is walked in 128-byte blocks (64 f16 / 128 int8).  Operands are staged by
// TMA (cp.async.bulk.tensor, 128-byte swizzle) into a shared-memory ring (3-4 stages of 48 KiB, or
// 5-6 of 32 KiB in the CTA-pair form), multiplied by single-thread tcgen05.mma instructions
// (kind::f16 with f32 accumulation, kind::i8 with s32 accumulation; cta_group::2 spanning a
// cluster of two CTAs for long rows) into TMEM, and the two 128x256 accumulators in TMEM are
// double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.  The distance matrix
// never leaves TMEM / registers: each epilogue thread owns one query (one TMEM lane), reads its
// 256 scores with tcgen05.ld, turns them into the metric's ranking value and keeps only values
// that beat the query's current k-th best (threshold in a register, a 3-input min tree per 32
// columns as the common-case test, rare appends to a per-query buffer, warp-cooperative bitonic
// compaction between tiles once a buffer passes its trigger level).
//
// Warp roles (64 + 128*EG threads): warp 0 = TMA producer (also paces itself against the other
// clusters working on the same database tiles, so those stay in L2), warp 1 = TMEM allocator +
// MMA issuer, then EG epilogue warpgroups (TMEM lane quarter = warp_id % 4, column slice = group),
// each warp decoupled from the others (no CTA-wide barrier in the steady state).
// Replaces the same reference loop as K1 (examples/similarity_search.rs:147-169) for large batches.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "tc_path.hpp"

namespace mvf {
namespace tc {

using k::kInf;
static_assert(BM == 128 && BN == 256 && BKB == 128, "tile shape the descriptors below are written for");

struct TcParams {
    const float* norm;             // sqrt(sum x^2) per resident row
    const float* sq;               // sum x^2 per row (f32, or int32 bits for integer spaces)
    const uint32_t* mask;          // optional row mask (bit set = excluded), same layout as ScanParams::mask
    const float* qconst;           // per query constant (prep_queries); the kernel itself reads it for integer cosine only (|q|)
    const int* acc_seed;           // integer L2: per row, padded to whole tiles, what the accumulator column starts from:
                                   // -(|x|^2 >> 1) for live rows, kSeedDead for rows past the end / deleted / filtered out
    u64* cand;                     // [gridDim.x][EG][BM][cap] scratch candidate buffers
    u64* out_lists;                // [nq][chunks*EG][k] sorted lists (key = ordered ranking value, id = local row)
    float* dump;                   // optional [mtiles*BM][ntiles*BN] raw ranking values (tests)
    uint32_t* gthr;                // [(mtiles rounded up to the cluster size)*BM] tightest bound on the k-th-best key any CTA has established per query (atomicMin)
    uint32_t* gslot;               // [same rows][TC_SLOTS] best ceil(k / TC_SLOTS)-th key of the lists in each slot class (see share_bound)
    uint64_t n;                    // resident rows
    uint32_t nq, k, cap;
    uint32_t trigger;              // a buffer holding more than this many entries after a tile is compacted (<= cap - columns per thread)
    uint32_t kblocks;              // ceil(d * elem / 128)
    uint32_t mtiles, ntiles, chunks, tiles_per_chunk;
    int metric;                    // k::M_L2 / M_IP / M_COS
    int kind;                      // KIND_F16: f16 operands / f32 accum, KIND_I8: int8 / s32, KIND_TF32: two tf32 planes / f32
    uint32_t idesc;                // tcgen05 instruction descriptor
    uint32_t l2_prefetch_tiles;    // database tiles the producer prefetches into L2 ahead of its TMA loads (0 = off)
    uint32_t stages;               // depth of the operand ring (<= MAX_STAGES)
    uint32_t* progress;            // [items] database tiles each work item's producer has issued (pacing, see sync_window)
    uint32_t sync_window;          // 0 = off; else a producer stays within this many tiles of the slowest peer on its rows
    unsigned long long* stats;     // -DMVF_TC_STATS builds only: cycle counters per role (see tc_path.cuh), else unused
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"   // suspend-time hint: sleep in hardware, not in the loop
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
#define MVF_UMMA_ASM(GROUP, KINDSTR)                                                              \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                \
                 "tcgen05.mma.cta_group::" GROUP ".kind::" KINDSTR " [%0], %1, %2, %3, p;\n\t}"     \
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory")
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_I8) MVF_UMMA_ASM("1", "i8");
    else if (KIND == KIND_TF32) MVF_UMMA_ASM("1", "tf32");
    else MVF_UMMA_ASM("1", "f16");
}
// ---- CTA-pair (cta_group::2) forms: one MMA spans the two SMs of a cluster, M = 256 (128 query rows per CTA),
// each CTA supplies half of the B operand from its own shared memory ----
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_I8) MVF_UMMA_ASM("2", "i8");
    else if (KIND == KIND_TF32) MVF_UMMA_ASM("2", "tf32");
    else MVF_UMMA_ASM("2", "f16");
}
#undef MVF_UMMA_ASM
// arrives on the barrier at the same shared-memory offset in every CTA of `mask` once the pair's MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {     // issued by the same warp of both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the barrier at the same offset in the pair's
// leader (even) CTA: shared-window addresses of a pair differ in bit 24 only
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// relaxed: the barrier only hands TMEM back (the tcgen05.ld results are already in registers after wait::ld), no
// generic-memory data rides on it; the default release at cluster scope costs a MEMBAR.ALL.GPU per warp per tile,
// which also waits for the candidate appends in flight (ncu: stall_membar 8 % of samples)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // arrivals may come from the peer CTA
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
    uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
          "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
          "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 128 bytes, 8-row groups
// 1024 bytes apart): start>>4 | LBO=1 | SBO=64 | version=1 | layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// warp-cooperative bitonic sort of P (power of two) words in shared memory
__device__ __forceinline__ void warp_bitonic_sort(u64* a, int P, int lane) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int i = lane; i < (P >> 1); i += 32) {
                int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                int hi = lo | stride;
                bool up = (lo & size) == 0;
                u64 x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncwarp();
}

// Accumulator seed of a column that must never produce a candidate (integer L2): with |q.x| < 2^24 the column's
// value stays below every threshold the epilogue can hold (thr_i >= -2^30 - 1).
constexpr int kSeedDead = (int)0x80000000 + (1 << 25);

// Cross-list threshold.  A query's rows are cut into many independent lists (work items x column groups); a list's own
// k-th best only bounds the global k-th best as tightly as ITS slice of the rows allows (config C: 74 lists of 135k
// rows each, so every list kept collecting ~k candidates of its own).  Lists are disjoint sets of rows, so if each of
// TC_SLOTS different lists holds j = ceil(k / TC_SLOTS) rows with key <= T, at least k rows do: T bounds the global k-th
// best.  Every list publishes its j-th best into slot (list % TC_SLOTS) by atomicMin (a slot keeps the best j-th key
// of its class; values only fall, so a stale read is merely conservative) and whoever publishes folds the maximum of
// the slots into gthr, where every CTA already looks.
constexpr int TC_SLOTS = 8;
__device__ __forceinline__ void share_bound(const TcParams& p, uint32_t q, uint32_t list, uint32_t kth_key, bool have_kth,
                                            uint32_t jth_key, bool have_jth) {
    if (have_kth) atomicMin(p.gthr + q, kth_key);
    if (!have_jth) return;
    uint32_t* slots = p.gslot + (size_t)q * TC_SLOTS;
    atomicMin(slots + (list % TC_SLOTS), jth_key);
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < TC_SLOTS; i++) t = max(t, __ldcg(slots + i));       // 0xffffffff while a class has not published
    if (t != 0xffffffffu) atomicMin(p.gthr + q, t);
}

// Compacts the candidate buffer of one query (owned by lane `owner` of this warp) to its k best,
// sorted ascending; returns the new entry count and the k-th best word (kInf while fewer than k).
__device__ __forceinline__ u64 compact_buffer(u64* gbuf, int cnt, int k, int cap, u64* scratch, int lane, int* new_cnt, u64* jth = nullptr) {
    int P = 32;
    while (P < cnt) P <<= 1;
    if (P > cap) P = cap;
    for (int i = lane; i < P; i += 32) scratch[i] = i < cnt ? gbuf[i] : kInf;
    warp_bitonic_sort(scratch, P, lane);
    const int keep = cnt < k ? cnt : k;
    for (int i = lane; i < keep; i += 32) gbuf[i] = scratch[i];
    const u64 kth = (cnt >= k) ? scratch[k - 1] : kInf;
    if (jth) { const int j = (k + TC_SLOTS - 1) / TC_SLOTS; *jth = (cnt >= j) ? scratch[j - 1] : kInf; }   // see share_bound
    __syncwarp();
    *new_cnt = keep;
    return kth;
}

#ifdef MVF_TC_STATS
#define TC_STAT_BEGIN(var) const long long var = clock64()
#define TC_STAT_ADD(acc, var) acc += clock64() - var
#else
#define TC_STAT_BEGIN(var)
#define TC_STAT_ADD(acc, var)
#endif

// MODE selects the ranking value computed from the accumulator `a` and the per-column parameter:
//   MODE 0  (L2)            s = prm - 2a        prm = |x|^2           (float: fma; int: exact)
//   MODE 1  (cosine / dot)  s = a * prm         prm = -1/|x| or -1    (float)   |  s = -a (int dot)
//   MODE 3  (integer cosine) the exact key is 1 - a / (|q| |x|) in the scan kernel's float operations (a division per
//           element); the common path tests the approximate s = float(a) * (-1/|x|) against a bound widened by more than
//           the rounding of both forms, and only the rare candidates get the exact key -- which is what their buffers,
//           the shared bounds and the merges then compare, so the result is bit-identical to the scan kernel's.
// CL = 2: CTA pairs (thread-block clusters) work on the same database tiles for two different query
// tiles; each CTA fetches half of the 256-row database tile and TMA-multicasts it into both CTAs'
// shared memory, so L2 -> SM operand traffic per tile drops from 576 KB to 384 KB and the pair
// stays in lockstep on the same rows.
// PAIR (with CL = 2): the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per K step: 256 queries x 256 rows,
// each CTA stages its own 128 query rows and HALF of the database tile (32 KiB per stage instead of 48, so the
// ring is 5-6 deep), the leader CTA issues, both epilogues drain their own TMEM.  Against the multicast form
// this cuts the bytes written into each SM's shared memory per tile by a third and halves the tensor core's
// shared-memory reads of the B operand.
// KIND_TF32 (K2', Float32 spaces): operands come as two tf32-exact planes (map_q/map_x = hi, map_q2/map_x2 = lo);
// a stage holds [Q_hi | Q_lo | X_hi | X_lo] and every K step issues three MMAs (hi.hi, lo.hi, hi.lo), so the
// accumulator carries q.x to ~2^-21 relative (the dropped lo.lo term) -- inside the float tolerance of the scan.
template <int KIND, int MODE, int CL, int EG, bool PAIR>
// (64 + 128 * EG threads, one CTA per SM.  With four epilogue warpgroups that is 18 warps, five on two of the four SM
// sub-partitions, whose 16K-register files then cap a thread at 102 registers: ptxas stops at 96 and a __maxnreg__ of
// 104 is refused at launch -- measured.)
__global__ void __launch_bounds__(threads_for(EG), 1)
gemm_topk_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
             const __grid_constant__ CUtensorMap map_q2, const __grid_constant__ CUtensorMap map_x2, const TcParams p) {
    static_assert(!PAIR || CL == 2, "the pair form needs clusters of two");
    static_assert(KIND != KIND_TF32 || PAIR || CL == 1, "split-tf32 stages come in the pair form or for single CTAs");
    constexpr bool INT = KIND == KIND_I8;
    constexpr bool SPLIT = KIND == KIND_TF32;
    // Integer L2 ("accumulator-init" form): the epilogue warps pre-load every accumulator column with -(|x|^2 >> 1)
    // (tcgen05.st, two tiles ahead) and the MMAs accumulate on top of it, so a tile comes out of TMEM as
    // r = q.x - (|x|^2 >> 1) and the ranking value is s = |x|^2 - 2 q.x = (|x|^2 & 1) - 2r: the common path is a
    // max tree over r (one VIMNMX3 per two elements, like the dot-product metric) against the conservative bound
    // r >= -(thr >> 1); the exact s (with the parity bit) is only formed for the rare candidates.  The per-element
    // IADD3 (|x|^2 - 2a) and the parameter LDS of the old form were 2/3 of the epilogue's instructions on short
    // rows (config E batched: 512 MMA cycles per tile against ~2300 measured).
    constexpr bool INIT_ACC = INT && MODE == 0;
    constexpr bool ICOS = INT && MODE == 3;
    static_assert(MODE != 3 || INT, "MODE 3 is the integer cosine form");
    constexpr int EPI_GROUPS = EG, EPI_THREADS = 128 * EG, EPI_COLS = BN / EG, LDG = 1;
    constexpr int PSLOTS = 3;                                           // column-parameter slots per epilogue warp (tc_path.hpp smem_bytes)
    constexpr int SB = stage_bytes_of(KIND, PAIR);
    constexpr int XB = PAIR ? B_BYTES / 2 : B_BYTES;                    // bytes of one database-tile plane in this CTA
    constexpr int X_OFF = planes_of(KIND) * A_BYTES;                    // database planes follow the query planes
    constexpr int KELEMS = KIND == KIND_I8 ? 128 : (KIND == KIND_F16 ? 64 : 32);   // elements per 128-byte K block
    const uint32_t STAGES = p.stages;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve shared memory: operand ring (1024-byte aligned) | sort scratch | column params | barriers
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = smem;
    u64* scratch = reinterpret_cast<u64*>(ring + STAGES * SB);           // [epilogue warps][cap]
    float* prm = reinterpret_cast<float*>(scratch + (EPI_THREADS / 32) * p.cap);  // [epilogue warps][PSLOTS][EPI_COLS], warp-private
    uint64_t* bars = reinterpret_cast<uint64_t*>(prm + (EPI_THREADS / 32) * PSLOTS * EPI_COLS);   // full[S] empty[S] tfull[2] tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
    uint32_t* valid_s = tmem_slot + 4;                                                    // [epilogue warps][PSLOTS][EPI_COLS / 32] live-column words

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MAX_STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * MAX_STAGES), tempty0 = smem_u32(bars + 2 * MAX_STAGES + 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_x);
        if (SPLIT) { tma_prefetch_desc(&map_q2); tma_prefetch_desc(&map_x2); }
        for (uint32_t s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, PAIR ? 1 : CL); }
        // pair form: the leader's MMA thread waits for the epilogue warps of BOTH CTAs before it overwrites an accumulator
        for (int b = 0; b < 2; b++) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, (PAIR ? 2 : 1) * (EPI_THREADS / 32)); }
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(smem_u32(tmem_slot), 512); else tmem_alloc(smem_u32(tmem_slot), 512); }
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                     // peers' barriers exist before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work items: (query-tile group of CL tiles, run of database tiles); the CTAs of a cluster take the
    // same item and differ only in the query tile
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    const uint32_t mgroups = (p.mtiles + CL - 1) / CL;
    const uint32_t items = mgroups * p.chunks;
    const uint32_t w_first = blockIdx.x / CL, w_step = gridDim.x / CL;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            [[maybe_unused]] long long st_total = 0, st_empty = 0, st_pace = 0;
            TC_STAT_BEGIN(c_all);
            for (uint32_t w = w_first; w < items; w += w_step) {
                const uint32_t mt = (w % mgroups) * CL + crank, ch = w / mgroups;
                const uint32_t t0 = ch * p.tiles_per_chunk;
                const uint32_t t1 = min(t0 + p.tiles_per_chunk, p.ntiles);
                // Pacing.  The mgroups work items of one run of database tiles execute at the same time on different
                // clusters and should find those tiles in L2; nothing else keeps them together, and once they drift
                // apart by more than L2 holds, every one of them streams the rows from HBM again (measured on config
                // C, 16 items per run: 142 GB read for 15 GB of rows).  So a producer does not run more than
                // sync_window tiles ahead of the slowest item on the same rows in the same scheduling round.  All CTAs
                // are resident (grid <= SM count) and an item only ever waits for items of its own round, so the wait
                // cannot cycle; it is bounded anyway and falls back to "just go".
                const bool pace = p.sync_window != 0 && crank == 0;
                const uint32_t round_lo = (w / w_step) * w_step;
                const uint32_t g_lo = max(ch * mgroups, round_lo), g_hi = min(min((ch + 1) * mgroups, round_lo + w_step), items);
                for (uint32_t t = t0; t < t1; t++) {
                    if (pace && ((t - t0) & 3u) == 0) {
                        const uint32_t r = t - t0;
                        st_relaxed_gpu(p.progress + w, r);
                        TC_STAT_BEGIN(c_p);
                        if (r > p.sync_window) {
                            for (int spin = 0; spin < 2048; spin++) {
                                uint32_t slowest = 0xffffffffu;
                                for (uint32_t x = g_lo; x < g_hi; x++) slowest = min(slowest, ld_relaxed_gpu(p.progress + x));
                                if (r <= slowest + p.sync_window) break;
                                __nanosleep(512);
                            }
                        }
                        TC_STAT_ADD(st_pace, c_p);
                    }
                    for (uint32_t kb = 0; kb < p.kblocks; kb++) {
                        TC_STAT_BEGIN(c_e);
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);          // every CTA of the cluster released the slot
                        TC_STAT_ADD(st_empty, c_e);
                        const uint32_t a_dst = smem_u32(ring + stage * SB);
                        const int kc = (int)(kb * (uint32_t)KELEMS);     // element coordinate along K
                        if (PAIR) {
                            // both CTAs' bytes are counted on the leader's barrier, which alone expects them
                            if (crank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * SB);
                            tma_load_2d_pair(a_dst, &map_q, full0 + 8 * stage, kc, (int)(mt * BM));
                            tma_load_2d_pair(a_dst + X_OFF, &map_x, full0 + 8 * stage, kc, (int)(t * BN + crank * (BN / 2)));
                            if (SPLIT) {
                                tma_load_2d_pair(a_dst + A_BYTES, &map_q2, full0 + 8 * stage, kc, (int)(mt * BM));
                                tma_load_2d_pair(a_dst + X_OFF + XB, &map_x2, full0 + 8 * stage, kc, (int)(t * BN + crank * (BN / 2)));
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(full0 + 8 * stage, SB);
                        tma_load_2d(a_dst, &map_q, full0 + 8 * stage, kc, (int)(mt * BM));
                        if (SPLIT) {                                       // single-CTA split form: four plain tile loads
                            tma_load_2d(a_dst + A_BYTES, &map_q2, full0 + 8 * stage, kc, (int)(mt * BM));
                            tma_load_2d(a_dst + X_OFF, &map_x, full0 + 8 * stage, kc, (int)(t * BN));
                            tma_load_2d(a_dst + X_OFF + XB, &map_x2, full0 + 8 * stage, kc, (int)(t * BN));
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        if (p.l2_prefetch_tiles && t + p.l2_prefetch_tiles < t1)   // pull a later tile's rows from HBM into L2 now
                            tma_prefetch_l2_2d(&map_x, kc, (int)((t + p.l2_prefetch_tiles) * BN + crank * (BN / CL)));
                        if (CL > 1)      // this CTA's share of the database tile, delivered to every CTA of the cluster
                            tma_load_2d_mc(a_dst + X_OFF + crank * (B_BYTES / CL), &map_x, full0 + 8 * stage, kc,
                                           (int)(t * BN + crank * (BN / CL)), (uint16_t)((1u << CL) - 1));
                        else
                            tma_load_2d(a_dst + X_OFF, &map_x, full0 + 8 * stage, kc, (int)(t * BN));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if (pace) st_relaxed_gpu(p.progress + w, 0xffffffffu);     // finished: never holds anyone back
            }
#ifdef MVF_TC_STATS
            TC_STAT_ADD(st_total, c_all);
            if (p.stats) { atomicAdd(p.stats + 0, (unsigned long long)st_total); atomicAdd(p.stats + 1, (unsigned long long)st_empty); atomicAdd(p.stats + 2, (unsigned long long)st_pace); }
#endif
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && (!PAIR || crank == 0)) {                       // pair form: the leader CTA issues for both
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            [[maybe_unused]] long long st_total = 0, st_full = 0, st_tempty = 0;
            TC_STAT_BEGIN(c_all);
            for (uint32_t w = w_first; w < items; w += w_step) {
                const uint32_t ch = w / mgroups;
                const uint32_t t0 = ch * p.tiles_per_chunk;
                const uint32_t t1 = min(t0 + p.tiles_per_chunk, p.ntiles);
                for (uint32_t t = t0; t < t1; t++) {
                    TC_STAT_BEGIN(c_t);
                    // INIT_ACC: the epilogue also arrives once per accumulator up front (after seeding it), so there is
                    // no free first pass
                    if (PAIR) mbar_wait_cluster(tempty0 + 8 * acc, INIT_ACC ? acc_phase : acc_phase ^ 1);   // both CTAs' epilogues have drained it
                    else mbar_wait(tempty0 + 8 * acc, INIT_ACC ? acc_phase : acc_phase ^ 1);      // epilogue has drained (and re-seeded) this accumulator
                    TC_STAT_ADD(st_tempty, c_t);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    for (uint32_t kb = 0; kb < p.kblocks; kb++) {
                        TC_STAT_BEGIN(c_f);
                        mbar_wait(full0 + 8 * stage, phase);           // TMA bytes have landed
                        TC_STAT_ADD(st_full, c_f);
                        tcgen05_fence_after();
                        const uint32_t a_addr = smem_u32(ring + stage * SB);
                        const uint64_t adesc = make_sdesc(a_addr), bdesc = make_sdesc(a_addr + X_OFF);
                        [[maybe_unused]] const uint64_t adesc_lo = make_sdesc(a_addr + A_BYTES), bdesc_lo = make_sdesc(a_addr + X_OFF + XB);
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ks++) {            // 4 x 32 bytes of K per 128-byte block
                            if (PAIR) umma_pair<KIND>(d_tmem, adesc + 2 * ks, bdesc + 2 * ks, p.idesc, INIT_ACC || (kb | ks) != 0);
                            else umma<KIND>(d_tmem, adesc + 2 * ks, bdesc + 2 * ks, p.idesc, INIT_ACC || (kb | ks) != 0);
                            if (SPLIT) {                                 // the two cross terms of the split product
                                if (PAIR) {
                                    umma_pair<KIND>(d_tmem, adesc_lo + 2 * ks, bdesc + 2 * ks, p.idesc, 1u);
                                    umma_pair<KIND>(d_tmem, adesc + 2 * ks, bdesc_lo + 2 * ks, p.idesc, 1u);
                                } else {
                                    umma<KIND>(d_tmem, adesc_lo + 2 * ks, bdesc + 2 * ks, p.idesc, 1u);
                                    umma<KIND>(d_tmem, adesc + 2 * ks, bdesc_lo + 2 * ks, p.idesc, 1u);
                                }
                            }
                        }
                        if (PAIR) umma_commit_pair(empty0 + 8 * stage, 3);
                        else if (CL > 1) umma_commit_mc(empty0 + 8 * stage, (uint16_t)((1u << CL) - 1));   // slot free in every CTA
                        else umma_commit(empty0 + 8 * stage);            // frees the smem slot when the MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (PAIR) umma_commit_pair(tfull0 + 8 * acc, 3);      // each CTA's half of the accumulator is ready
                    else umma_commit(tfull0 + 8 * acc);                   // accumulator ready for the epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
#ifdef MVF_TC_STATS
            TC_STAT_ADD(st_total, c_all);
            if (p.stats) { atomicAdd(p.stats + 3, (unsigned long long)st_total); atomicAdd(p.stats + 4, (unsigned long long)st_full); atomicAdd(p.stats + 5, (unsigned long long)st_tempty); }
#endif
        }
    } else {
        // ===== epilogue: EG warpgroups x 4 warps; a thread owns one query (TMEM lane) and the
        // warpgroup's slice of the tile's columns =====
        const int quarter = warp & 3;                                   // TMEM lanes 32*quarter .. +31
        const int grp = (warp - 2) >> 2;                                // column slice
        const int qrow = quarter * 32 + lane;                           // query row inside the tile
        u64* wscratch = scratch + (size_t)(warp - 2) * p.cap;
        u64* const cta_cand = p.cand + ((size_t)blockIdx.x * EPI_GROUPS + grp) * BM * p.cap;
        u64* mybuf = cta_cand + (size_t)qrow * p.cap;
        // Per-column parameter of the ranking value and whether the column is a live row (inside the database,
        // not deleted / filtered out).  Every epilogue warp stages the EPI_COLS columns of ITS slice in a private
        // shared-memory slice ahead of time (EPI_COLS / 32 columns per lane), so the epilogue needs no CTA-wide
        // barrier.  PSLOTS slots: float kinds work one tile ahead (slot = accumulator), integer kinds three.
        constexpr int CPLANE = EPI_COLS / 32;
        float* my_prm = prm + (size_t)(warp - 2) * PSLOTS * EPI_COLS;
        uint32_t* my_valid = valid_s + (warp - 2) * PSLOTS * CPLANE;
        // Split in two so the global loads of a later tile are ISSUED before the current tile is processed and only
        // CONSUMED after it (their ~1 us latency used to sit in front of every tile's accumulator wait): load_raw
        // returns the untouched words, finish_param turns them into (parameter, live).
        const uint32_t n32 = (uint32_t)p.n;                             // a resident shard holds < 2^32 rows
        const uint32_t col0 = (uint32_t)(grp * EPI_COLS + lane);
        auto load_raw = [&](uint32_t t, int j, float* raw, uint32_t* mword) {
            const uint32_t row = t * BN + col0 + j * 32;                  // (the padded tile range stays below 2^32 + BN: wraps only past n)
            if (INIT_ACC) {                                               // the seed array is padded to whole tiles and already knows dead rows
                *raw = __int_as_float(p.acc_seed[(size_t)row]); *mword = 0u;
                return;
            }
            const bool in = row < n32 && row >= t * BN;
            *mword = (in && p.mask != nullptr) ? p.mask[row >> 5] : 0u;
            float v = 0.f;
            if (in) {
                if (MODE == 0) v = p.sq[row];                             // |x|^2 (int32 bits for integer spaces)
                else if ((!INT && p.metric == k::M_COS) || ICOS) v = p.norm[row];
            }
            *raw = v;
        };
        auto finish_param = [&](uint32_t t, int j, float raw, uint32_t mword, bool* live) -> float {
            const uint32_t row = t * BN + col0 + j * 32;
            const bool ok = row < n32 && row >= t * BN && ((mword >> (row & 31)) & 1u) == 0;
            *live = ok;
            if (!ok) return __int_as_float(0);
            if (MODE == 0) return raw;
            if (INT && !ICOS) return __int_as_float(0);
            if (ICOS || p.metric == k::M_COS) return raw > 0.f ? -1.0f / raw : 0.0f;
            return -1.0f;
        };
        // publishes the parameters of database tile t in slot `slot` of the warp's private slice (from raw words loaded earlier)
        auto publish_params = [&](uint32_t slot, uint32_t t, const float* raw, const uint32_t* mw) {
            float* prn = my_prm + slot * EPI_COLS;
#pragma unroll
            for (int j = 0; j < CPLANE; j++) {
                if (INIT_ACC) { prn[j * 32 + lane] = raw[j]; continue; }    // the seed is all the epilogue needs per column
                bool lv;
                prn[j * 32 + lane] = finish_param(t, j, raw[j], mw[j], &lv);
                const unsigned vw = __ballot_sync(0xffffffffu, lv);
                if (lane == 0) my_valid[slot * CPLANE + j] = vw;
            }
        };
        // INIT_ACC: seeds accumulator `a` (this warp's lane quarter and column slice) with the values published in `slot`
        [[maybe_unused]] auto seed_accumulator = [&](uint32_t a, uint32_t slot) {
            const uint4* pi = reinterpret_cast<const uint4*>(my_prm + slot * EPI_COLS);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + grp * EPI_COLS;
#pragma unroll 1                                                    // (unrolled even by two, ptxas hoists the loads: tile values + 32-64 seeds live -> spills)
            for (int c = 0; c < EPI_COLS / 16; c++) {
                uint32_t sv[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const uint4 x = pi[c * 4 + j4];
                    sv[4 * j4] = x.x; sv[4 * j4 + 1] = x.y; sv[4 * j4 + 2] = x.z; sv[4 * j4 + 3] = x.w;
                }
                tmem_st16(taddr + c * 16, sv);
            }
            tmem_wait_st();
        };
        auto release_accumulator = [&](uint32_t a) {                    // all epilogue warps -> accumulator free
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(mapa_cluster(tempty0 + 8 * a, 0));
                else mbar_arrive(tempty0 + 8 * a);
            }
        };
        auto item_tiles = [&](uint32_t w, uint32_t* a, uint32_t* b) {
            const uint32_t c0 = (w / mgroups) * p.tiles_per_chunk;
            *a = c0; *b = min(c0 + p.tiles_per_chunk, p.ntiles);
        };
        // Integer kinds: the CTA's tiles form ONE sequence across its work items (tile number i, parameter slot
        // i % 3, accumulator i % 2) and the epilogue works ahead of the drain: parameters are published three
        // tiles ahead, accumulators (INIT_ACC) are seeded two tiles ahead.
        [[maybe_unused]] uint32_t la_t1 = 0;                              // end of the look-ahead cursor's item
        auto seq_advance = [&](uint32_t* w, uint32_t* t) -> bool {       // false: the CTA has no further tile
            if (*t + 1 < la_t1) { (*t)++; return true; }
            *w += w_step;
            if (*w >= items) return false;
            uint32_t a;
            item_tiles(*w, &a, &la_t1);                                   // (one integer division per work item, not per tile)
            *t = a;
            return true;
        };
        const bool dump = p.dump != nullptr;                            // tests only
        uint32_t acc = 0, acc_phase = 0;
        [[maybe_unused]] uint32_t seq_i = 0, seq_total = 0;             // running tile number / tiles of this CTA
        [[maybe_unused]] uint32_t seq_slot = 0;                         // seq_i % 3, kept incrementally
        [[maybe_unused]] uint32_t la_w = w_first, la_t = 0;             // look-ahead cursor: the next tile to publish
        [[maybe_unused]] bool la_ok = w_first < items;
        if (INT) {
            for (uint32_t w = w_first; w < items; w += w_step) { uint32_t a, b; item_tiles(w, &a, &b); seq_total += b - a; }
            if (la_ok) item_tiles(la_w, &la_t, &la_t1);
            // prologue: tiles #0..#2 get their parameters; INIT_ACC seeds both accumulators (tiles #0, #1) and hands
            // them to the issuer before any MMA is issued
            for (uint32_t slot = 0; slot < 3; slot++) {
                float raw[CPLANE]; uint32_t mw[CPLANE];
#pragma unroll
                for (int j = 0; j < CPLANE; j++) { raw[j] = 0.f; mw[j] = 0; }
                if (la_ok) {
#pragma unroll
                    for (int j = 0; j < CPLANE; j++) load_raw(la_t, j, &raw[j], &mw[j]);
                }
                publish_params(slot, la_ok ? la_t : 0u, raw, mw);
                if (la_ok) la_ok = seq_advance(&la_w, &la_t);
            }
            __syncwarp();
            if (INIT_ACC) {
                for (uint32_t a = 0; a < 2; a++) { seed_accumulator(a, a); release_accumulator(a); }
            }
        }
        [[maybe_unused]] long long st_total = 0, st_tfull = 0, st_compact = 0, st_tail = 0, st_drain = 0, st_seed = 0, st_proc = 0, st_pub = 0;
        TC_STAT_BEGIN(c_all);
        for (uint32_t w = w_first; w < items; w += w_step) {
            const uint32_t mt = (w % mgroups) * CL + crank, ch = w / mgroups;
            const uint32_t t0 = ch * p.tiles_per_chunk;
            const uint32_t t1 = min(t0 + p.tiles_per_chunk, p.ntiles);
            const uint32_t q = mt * BM + qrow;
            int cnt = 0;
            // thr_key: tightest known bound on this query's k-th best key (local list or any CTA's, via gthr);
            // loc_key: k-th best key of the local list.  Rows arrive in increasing index order, so a row that
            // only TIES the local k-th can never displace it (strict compare); a tie with a bound that came from
            // another CTA may still win on the index, so it is kept (<=).  One compare per element either way:
            // the strict case compares against the next smaller key.
            uint32_t thr_key = 0xffffffffu, loc_key = 0xffffffffu;
            float thr_f = __int_as_float(0x7f800000);                   // +inf: everything passes until k are known
            // Integer spaces never negate in the common path: they look for the LARGEST accumulator value.
            //   dot (MODE 1): s = -a, a hit is a > ~t (== -a <= t for every t, no overflow at either end).  (Written as
            //     min(-a0, -a1, -a2), ptxas 12.9 folded the negations into VIMNMX3 incorrectly and real hits were skipped.)
            //   L2 (MODE 0, INIT_ACC): the tile holds r = a - (|x|^2 >> 1), s = (|x|^2 & 1) - 2r; s <= t implies
            //     r >= -(t >> 1), so the common path looks for r > thr_i = -(t >> 1) - 1 and the exact test s <= thr_s
            //     runs on the rare candidates only (no overflow: |t >> 1| <= 2^30).
            int thr_i = 0x7fffffff;
            [[maybe_unused]] int thr_s = 0x7fffffff;
            // integer cosine: |q| (as the scan kernel computes it), the widened bound on the approximate ranking value and
            // the exact bound on the key
            [[maybe_unused]] const float qn = (ICOS && q < p.nq) ? p.qconst[q] : 0.0f;
            [[maybe_unused]] float thr_fc = __int_as_float(0x7f800000);
            [[maybe_unused]] uint32_t thr_ck = 0xffffffffu;
            auto set_thr_i = [&](int t) { if (INIT_ACC) { thr_s = t; thr_i = -(t >> 1) - 1; } else thr_i = ~t; };
            auto refresh_threshold = [&]() {
                if (q >= p.nq) { thr_f = thr_fc = __int_as_float(0xff800000); set_thr_i((int)0x80000000); return; }   // padding row of the query tile: nothing passes
                if (thr_key == 0xffffffffu) { thr_f = thr_fc = __int_as_float(0x7f800000); thr_ck = 0xffffffffu; set_thr_i(0x7fffffff); return; }
                const bool ties = thr_key < loc_key;
                if (!ties && thr_key == 0) { thr_f = thr_fc = __int_as_float(0xff800000); set_thr_i((int)0x80000000); return; }
                const uint32_t ck = ties ? thr_key : thr_key - 1;
                if (ICOS) {
                    // key <= ck  <=>  1 - a / (|q||x|) <= tv; in the approximate form s = -a / |x|:  s <= (tv - 1) |q|.  The bound
                    // is widened by the rounding of both forms (1 - x rounds to 6e-8 absolute, the products to ~1e-7 relative);
                    // a zero query (every key is exactly 1) sends everything to the exact test.
                    thr_ck = ck;
                    const float tv = k::unordered_f32(ck);
                    thr_fc = (qn > 0.0f && tv == tv) ? (tv - 1.0f + 4e-7f + 4e-6f * fabsf(tv - 1.0f)) * qn : __int_as_float(0x7f800000);
                    return;
                }
                if (INT) set_thr_i((int)(ck ^ 0x80000000u)); else thr_f = k::unordered_f32(ck);
            };
            refresh_threshold();
            {
                if (!INT) {                                              // (integer kinds: published three tiles ago)
                    float raw0[CPLANE]; uint32_t mw0[CPLANE];
#pragma unroll
                    for (int j = 0; j < CPLANE; j++) load_raw(t0, j, &raw0[j], &mw0[j]);
                    publish_params(acc, t0, raw0, mw0);
                }
                const uint32_t g0 = __ldcg(p.gthr + q);
                if (g0 < thr_key) { thr_key = g0; refresh_threshold(); }
                __syncwarp();
            }
            // appends the candidates of one 32-column group whose bit is set in `mask` (rare per thread)
            auto append_hits = [&](uint32_t mask, const uint32_t* v, uint32_t id0) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (mask & (1u << j)) {
                        uint32_t key;
                        if (ICOS) {
                            // the scan kernel's key, operation for operation (scan_kernels.cuh finalize): 1 - a / (|q| |x|), zero norm -> 1
                            const float den = __fmul_rn(qn, p.norm[id0 + j]);
                            const float vv = den > 0.0f ? __fsub_rn(1.0f, __fdiv_rn((float)(int)v[j], den)) : 1.0f;
                            key = k::ordered_f32(vv);
                            if (key > thr_ck) continue;                // passed the widened bound only
                        } else if (INT) {
                            // exact integer ranking value: s = (|x|^2 & 1) - 2r (L2; the parity comes from the row's |x|^2) or -a (dot)
                            const int s = INIT_ACC ? (__float_as_int(p.sq[id0 + j]) & 1) - 2 * (int)v[j] : -(int)v[j];
                            if (INIT_ACC && s > thr_s) continue;       // passed the conservative bound on r only
                            key = k::ordered_i32(s);
                        } else {
                            key = k::ordered_f32(__uint_as_float(v[j]));
                        }
                        mybuf[cnt++] = ((u64)key << 32) | (id0 + j);
                    }
                }
            };
            for (uint32_t t = t0; t < t1; t++) {
                const uint32_t slot = INT ? seq_slot : acc;               // parameter slot of this tile
                float nraw[CPLANE];
                uint32_t nmw[CPLANE];
#pragma unroll
                for (int j = 0; j < CPLANE; j++) { nraw[j] = 0.f; nmw[j] = 0; }
                if (INT ? la_ok : (t + 1 < t1)) {
#pragma unroll
                    for (int j = 0; j < CPLANE; j++) load_raw(INT ? la_t : t + 1, j, &nraw[j], &nmw[j]);
                }
                const uint32_t* vs = my_valid + slot * CPLANE;
                // any CTA's k-th best for this query bounds the global k-th best: adopt the tightest seen.
                // The word is read from L2 now and applied after this tile, so its latency is hidden.
                const uint32_t gk = __ldcg(p.gthr + q);
                TC_STAT_BEGIN(c_w);
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                TC_STAT_ADD(st_tfull, c_w);
                tcgen05_fence_after();
                TC_STAT_BEGIN(c_d);
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + grp * EPI_COLS;
                if constexpr (INT) {
                    // Integer kinds: once the LAST 32-column load of the tile has landed in registers
                    // the accumulator is re-seeded for the tile two steps ahead (INIT_ACC) and handed back to the
                    // issuer BEFORE its values are looked at: on short rows (config E: 512 MMA cycles per tile) the
                    // chain tfull -> ld -> seed -> tempty is what the tensor pipe waits for, the max tree is not.
                    constexpr int NCH = EPI_COLS / 32;
#pragma unroll
                    for (int cc = 0; cc < NCH; cc++) {
                        uint32_t v[32];
                        tmem_ld32(taddr + cc * 32, v);
                        tmem_wait_ld();
                        if (cc == NCH - 1) {
                            TC_STAT_ADD(st_drain, c_d);
                            TC_STAT_BEGIN(c_s);
                            if (INIT_ACC && seq_i + 2 < seq_total) seed_accumulator(acc, seq_slot == 0 ? 2u : seq_slot - 1u);   // slot of tile i + 2
                            release_accumulator(acc);
                            TC_STAT_ADD(st_seed, c_s);
                        }
                        TC_STAT_BEGIN(c_p);
                        const uint32_t id0 = t * BN + grp * EPI_COLS + cc * 32;
                        uint32_t mask = 0;
                        if constexpr (ICOS) {
                            // approximate ranking value s = float(a) * (-1 / |x|) and its minimum over the group
                            const float4* pr4 = reinterpret_cast<const float4*>(my_prm + slot * EPI_COLS + cc * 32);
                            float f0 = __int_as_float(0x7f800000), f1 = f0;
#pragma unroll
                            for (int j4 = 0; j4 < 8; j4++) {
                                const float4 pq = pr4[j4];
                                f0 = fminf(f0, fminf((float)(int)v[4 * j4] * pq.x, (float)(int)v[4 * j4 + 2] * pq.z));
                                f1 = fminf(f1, fminf((float)(int)v[4 * j4 + 1] * pq.y, (float)(int)v[4 * j4 + 3] * pq.w));
                            }
                            if (fminf(f0, f1) <= thr_fc) {
                                const float* prs = my_prm + slot * EPI_COLS + cc * 32;
#pragma unroll
                                for (int j = 0; j < 32; j++) if ((float)(int)v[j] * prs[j] <= thr_fc) mask |= 1u << j;
                            }
                        } else {
                            int m0 = (int)0x80000000, m1 = (int)0x80000000;   // two chains: max of a (dot) / of r (L2)
#pragma unroll
                            for (int j = 0; j < 32; j += 2) { m0 = max(m0, (int)v[j]); m1 = max(m1, (int)v[j + 1]); }
                            if (max(m0, m1) > thr_i) {
#pragma unroll
                                for (int j = 0; j < 32; j++) if ((int)v[j] > thr_i) mask |= 1u << j;
                            }
                        }
                        if (!INIT_ACC) mask &= vs[cc];                   // rows past the end, deleted or filtered out (INIT_ACC: dead seeds)
                        if (dump) {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const bool in = id0 + j < n32;
                                p.dump[((size_t)mt * BM + qrow) * ((size_t)p.ntiles * BN) + id0 + j] =
                                    ICOS ? (float)(int)v[j] * my_prm[slot * EPI_COLS + cc * 32 + j]
                                         : (float)(INIT_ACC ? (in ? (__float_as_int(p.sq[id0 + j]) & 1) - 2 * (int)v[j] : 0) : -(int)v[j]);
                            }
                        }
                        if (mask) append_hits(mask, v, id0);
                        TC_STAT_ADD(st_proc, c_p);
                    }
                } else {
                    const float* pr = my_prm + slot * EPI_COLS;
                    for (int c2 = 0; c2 < EPI_COLS / (32 * LDG); c2++) {
                        // LDG 32-column TMEM loads per wait
                        uint32_t vv[LDG][32];
#pragma unroll
                        for (int hh = 0; hh < LDG; hh++) tmem_ld32(taddr + c2 * (32 * LDG) + hh * 32, vv[hh]);
                        tmem_wait_ld();
                        // the tile's last columns are in registers: the accumulator goes back to the issuer before they are looked at
                        if (c2 == EPI_COLS / (32 * LDG) - 1) release_accumulator(acc);
#pragma unroll
                        for (int hh = 0; hh < LDG; hh++) {
                            uint32_t (&v)[32] = vv[hh];
                            const int cc = c2 * LDG + hh;
                            const uint32_t id0 = t * BN + grp * EPI_COLS + cc * 32;
                            // ranking values in place, and their minimum: almost every 32-column group holds nothing that
                            // beats the threshold, so the per-element compare + mask is only built when the minimum does
                            float fmin_ = __int_as_float(0x7f800000);
                            const float4* pr4 = reinterpret_cast<const float4*>(pr + cc * 32);
#pragma unroll
                            for (int j4 = 0; j4 < 8; j4++) {
                                const float4 pq = pr4[j4];
                                const float pj[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    const int j = j4 * 4 + e;
                                    const float a = __uint_as_float(v[j]);
                                    const float s = MODE == 0 ? fmaf(a, -2.0f, pj[e]) : a * pj[e];
                                    v[j] = __float_as_uint(s);
                                    fmin_ = fminf(fmin_, s);
                                }
                            }
                            uint32_t mask = 0;
                            if (fmin_ <= thr_f) {
#pragma unroll
                                for (int j = 0; j < 32; j++) if (__uint_as_float(v[j]) <= thr_f) mask |= 1u << j;
                            }
                            mask &= vs[cc];                              // rows past the end, deleted or filtered out
                            if (dump) {
#pragma unroll
                                for (int j = 0; j < 32; j++)
                                    p.dump[((size_t)mt * BM + qrow) * ((size_t)p.ntiles * BN) + id0 + j] = __uint_as_float(v[j]);
                            }
                            if (mask) append_hits(mask, v, id0);
                        }
                    }
                    TC_STAT_ADD(st_drain, c_d);
                }
                // A buffer past its trigger level is compacted now, by the whole warp -- after the accumulator went back to
                // the issuer, so the sort overlaps the next tiles' MMAs instead of holding up every other warp's tile.
                // trigger <= cap - EPI_COLS: whatever the next tile appends still fits.
                {
                    unsigned need = __ballot_sync(0xffffffffu, cnt > (int)p.trigger);
                    TC_STAT_BEGIN(c_c);
                    while (need) {
                        const int owner = __ffs(need) - 1;
                        need &= need - 1;
                        const int ocnt = __shfl_sync(0xffffffffu, cnt, owner);
                        u64* obuf = cta_cand + (size_t)(quarter * 32 + owner) * p.cap;
                        __syncwarp();
                        int ncnt;
                        u64 jth;
                        const u64 kth = compact_buffer(obuf, ocnt, (int)p.k, (int)p.cap, wscratch, lane, &ncnt, &jth);
                        if (lane == owner) {
                            cnt = ncnt;
                            if (kth != kInf) {
                                const uint32_t kk = (uint32_t)(kth >> 32);
                                loc_key = kk;
                                if (kk < thr_key) thr_key = kk;
                                refresh_threshold();
                            }
                            share_bound(p, q, ch * EPI_GROUPS + grp, (uint32_t)(kth >> 32), kth != kInf, (uint32_t)(jth >> 32), jth != kInf);
                        }
                    }
                    TC_STAT_ADD(st_compact, c_c);
                }
                if (gk < thr_key) { thr_key = gk; refresh_threshold(); }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                TC_STAT_BEGIN(c_u);
                if (INT) {
                    // the slot this tile just vacated takes the parameters of the tile three steps ahead
                    __syncwarp();
                    if (la_ok) { publish_params(slot, la_t, nraw, nmw); la_ok = seq_advance(&la_w, &la_t); }
                    seq_i++;
                    seq_slot = seq_slot == 2 ? 0u : seq_slot + 1u;
                } else if (t + 1 < t1) {
                    publish_params(acc, t + 1, nraw, nmw);                 // the next tile's parameters (warp-private)
                }
                __syncwarp();
                TC_STAT_ADD(st_pub, c_u);
            }
            // end of the work item: every query's buffer becomes a sorted list of its k best
            TC_STAT_BEGIN(c_tail);
            for (int owner = 0; owner < 32; owner++) {
                const int ocnt = __shfl_sync(0xffffffffu, cnt, owner);
                u64* obuf = cta_cand + (size_t)(quarter * 32 + owner) * p.cap;
                __syncwarp();
                int ncnt;
                u64 jth;
                const u64 kth = compact_buffer(obuf, ocnt, (int)p.k, (int)p.cap, wscratch, lane, &ncnt, &jth);
                const uint32_t qo = mt * BM + quarter * 32 + owner;
                if (lane == 0) share_bound(p, qo, ch * EPI_GROUPS + grp, (uint32_t)(kth >> 32), kth != kInf, (uint32_t)(jth >> 32), jth != kInf);
                if (qo < p.nq) {
                    u64* dst = p.out_lists + ((size_t)qo * (p.chunks * EPI_GROUPS) + ch * EPI_GROUPS + grp) * p.k;
                    for (int i = lane; i < (int)p.k; i += 32) dst[i] = i < ncnt ? wscratch[i] : kInf;
                }
                __syncwarp();
            }
            TC_STAT_ADD(st_tail, c_tail);
        }
#ifdef MVF_TC_STATS
        TC_STAT_ADD(st_total, c_all);
        if (p.stats && lane == 0 && warp == 2) {
            atomicAdd(p.stats + 6, (unsigned long long)st_total); atomicAdd(p.stats + 7, (unsigned long long)st_tfull);
            atomicAdd(p.stats + 8, (unsigned long long)st_compact); atomicAdd(p.stats + 9, (unsigned long long)st_tail);
            atomicAdd(p.stats + 10, (unsigned long long)st_drain); atomicAdd(p.stats + 11, (unsigned long long)st_seed);
            atomicAdd(p.stats + 12, (unsigned long long)st_proc); atomicAdd(p.stats + 13, (unsigned long long)st_pub);
        }
#endif
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                     // no CTA leaves while a peer may still write into it
    if (warp == 1) { tcgen05_fence_after(); if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// query preparation: copy the batch into a 16-byte-pitched operand buffer and compute the
// per-query constant of the final score (cosine: 1/|q|; L2: |q|^2; integer L2: sum q^2 as int bits)
template <bool INT, bool SIGNED>
__global__ void __launch_bounds__(128) prep_queries(const void* q, unsigned char* out, uint32_t nq, uint32_t d, uint32_t pitch,
                                                    float* qconst, int metric) {
    const uint32_t qi = blockIdx.x;
    if (qi >= nq) return;
    const uint32_t es = INT ? 1 : 2;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(q) + (size_t)qi * d * es;
    unsigned char* dst = out + (size_t)qi * pitch;
    float fs = 0.f; int is = 0;
    for (uint32_t j = threadIdx.x; j < pitch / es; j += blockDim.x) {
        if (INT) {
            unsigned char b = j < d ? src[j] : 0;
            dst[j] = b;
            int v = SIGNED ? (int)(signed char)b : (int)b;
            is += v * v;
        } else {
            __half h = j < d ? reinterpret_cast<const __half*>(src)[j] : __float2half(0.f);
            reinterpret_cast<__half*>(dst)[j] = h;
            float v = __half2float(h);
            fs = fmaf(v, v, fs);
        }
    }
    __shared__ float rf[4]; __shared__ int ri[4];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { fs += __shfl_xor_sync(0xffffffffu, fs, off); is += __shfl_xor_sync(0xffffffffu, is, off); }
    if ((threadIdx.x & 31) == 0) { rf[threadIdx.x >> 5] = fs; ri[threadIdx.x >> 5] = is; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float f = rf[0] + rf[1] + rf[2] + rf[3]; int i = ri[0] + ri[1] + ri[2] + ri[3];
        float c;
        if (INT) c = metric == k::M_L2 ? __int_as_float(i) : (metric == k::M_COS ? __fsqrt_rn((float)i) : 0.f);   // cosine: |q| as the scan kernel computes it
        else if (metric == k::M_COS) { float nm = sqrtf(f); c = nm > 0.f ? 1.0f / nm : 0.0f; }
        else c = f;
        qconst[qi] = c;
    }
}

// ---- K2' (Float32 spaces): the two tf32-exact planes of an operand ------------------------------------------
// hi = tf32(x) (round to nearest, low 13 mantissa bits zero), lo = tf32(x - hi); x - hi is exact in f32, so
// x = hi + lo + O(2^-22 |x|) and both planes pass through the tensor core's tf32 inputs unchanged whatever
// rounding it applies to wider values.
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float x, float* hi, float* lo) {
    const float h = tf32_rn(x);
    *hi = h; *lo = tf32_rn(x - h);
}
#ifdef MVF_TC_PATH_TU   // non-template kernels: defined in tc_path.cu's translation unit only
// rows of an f32 shard -> [hi plane | lo plane], same pitch; one thread per 16-byte chunk
__global__ void __launch_bounds__(256) split_rows_tf32_kernel(const unsigned char* rows, uint64_t nbytes, unsigned char* out) {
    const uint64_t chunks = nbytes >> 4;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = k::ldg_stream16(rows + c * 16);
        const float x[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; e++) tf32_split(x[e], &h[e], &l[e]);
        *reinterpret_cast<float4*>(out + c * 16) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(out + nbytes + c * 16) = make_float4(l[0], l[1], l[2], l[3]);
    }
}
// query preparation for the split form: planes hi (out) and lo (out + plane_bytes), per-query constant from the f32 values
__global__ void __launch_bounds__(128) prep_queries_tf32(const float* q, unsigned char* out, size_t plane_bytes, uint32_t nq, uint32_t d,
                                                         uint32_t pitch, float* qconst, int metric) {
    const uint32_t qi = blockIdx.x;
    if (qi >= nq) return;
    const float* src = q + (size_t)qi * d;
    float* dh = reinterpret_cast<float*>(out + (size_t)qi * pitch);
    float* dl = reinterpret_cast<float*>(out + plane_bytes + (size_t)qi * pitch);
    float fs = 0.f;
    for (uint32_t j = threadIdx.x; j < pitch / 4; j += blockDim.x) {
        const float v = j < d ? src[j] : 0.f;
        float h, l;
        tf32_split(v, &h, &l);
        dh[j] = h; dl[j] = l;
        fs = fmaf(v, v, fs);
    }
    __shared__ float rf[4];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, off);
    if ((threadIdx.x & 31) == 0) rf[threadIdx.x >> 5] = fs;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float f = rf[0] + rf[1] + rf[2] + rf[3];
        float c;
        if (metric == k::M_COS) { const float nm = sqrtf(f); c = nm > 0.f ? 1.0f / nm : 0.0f; }
        else c = f;
        qconst[qi] = c;
    }
}
#endif  // MVF_TC_PATH_TU

}  // namespace tc
}  // namespace mvf
