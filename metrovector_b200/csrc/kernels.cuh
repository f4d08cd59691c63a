// kernels.cuh -- hand-written sm_100a kernels of the bandwidth path.
//   K0  row_norms        sqrt(sum x^2) and sum x^2 per row, once per resident space
//   K1  scan_topk        distance + running top-k, one pass over the HBM-resident rows
//   K3  merge_topk       per-query merge of candidate lists (per-CTA lists, or per-rank lists)
//   X1  exchange_merge   cross-GPU exchange of the local top-k over peer memory + the merge, one kernel
//   synth_fill           device-side synthetic rows (bench/tests)
// What they replace: the loop body of find_top_k_similar, examples/similarity_search.rs:147-169
// of the reference (get_vector -> as_f32 -> sum of squared differences -> heap push/pop).
// Candidates are 64-bit words (order-preserving 32-bit key << 32 | 32-bit row id inside the shard), so
// "smallest score, lowest index on ties" is plain unsigned order everywhere.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "synth.h"

namespace mvf {
namespace k {

typedef unsigned long long u64;
constexpr u64 kInf = ~0ull;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kPadRows = 256;   // zero rows kept after the last resident row (scan kernels over-read, never clamp)
inline int scan_rows_in_flight(int NQ) { return NQ <= 2 ? 8 : NQ == 4 ? 4 : 2; }   // R of scan_topk_*<.., NQ, R>
const void* scan_kernel_f32(int op, int NQ, bool wide);
const void* scan_kernel_f16(int op, int NQ, bool wide);
const void* scan_kernel_i8(int op, int NQ, bool wide);
const void* scan_kernel_u8(int op, int NQ, bool wide);

enum { DT_F32 = 0, DT_F16 = 1, DT_I8 = 2, DT_U8 = 3 };
enum { M_L2 = 0, M_IP = 1, M_COS = 2 };
enum { OP_L2 = 0, OP_DOT = 1 };   // inner-loop flavour: squared differences vs products

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {   // streaming 128-bit load, no L1 allocation
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ordered_f32(float v) {   // -0 -> +0, NaN last
    if (v != v) return 0xffffffffu;
    v = v + 0.0f;
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unordered_f32(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ uint32_t ordered_i32(int v) { return (uint32_t)v ^ 0x80000000u; }
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
    return ((u64)hi << 32) | lo;
}

template <int DT> struct Traits;
template <> struct Traits<DT_F32> { static constexpr int EPC = 4; static constexpr bool is_float = true; };
template <> struct Traits<DT_F16> { static constexpr int EPC = 8; static constexpr bool is_float = true; };
template <> struct Traits<DT_I8> { static constexpr int EPC = 16; static constexpr bool is_float = false; };
template <> struct Traits<DT_U8> { static constexpr int EPC = 16; static constexpr bool is_float = false; };

// widen one 16-byte chunk of a float space to f32 (Vector::as_f32, src/vectors/vector.rs:71-92)
template <int DT> __device__ __forceinline__ void widen(const uint4& x, float* out);
template <> __device__ __forceinline__ void widen<DT_F32>(const uint4& x, float* out) {
    out[0] = __uint_as_float(x.x); out[1] = __uint_as_float(x.y); out[2] = __uint_as_float(x.z); out[3] = __uint_as_float(x.w);
}
template <> __device__ __forceinline__ void widen<DT_F16>(const uint4& x, float* out) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        float2 f = __half22float2(h);
        out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
}
template <int DT> __device__ __forceinline__ int dp4(uint32_t a, uint32_t b, int c);
template <> __device__ __forceinline__ int dp4<DT_I8>(uint32_t a, uint32_t b, int c) { return __dp4a((int)a, (int)b, c); }
template <> __device__ __forceinline__ int dp4<DT_U8>(uint32_t a, uint32_t b, int c) { return (int)__dp4a(a, b, (unsigned)c); }

// ---- per-warp running top-k: unsorted list in shared memory + (max, argmax) in registers --------
__device__ __forceinline__ void list_insert(u64* list, int k, u64 cand, u64& thr, int& pos, int lane) {
    __syncwarp();
    if (lane == 0) list[pos] = cand;
    __syncwarp();
    u64 m = 0; int mp = 0x7fffffff;
    for (int j = lane; j < k; j += 32) { u64 v = list[j]; if (v > m || mp == 0x7fffffff) { m = v; mp = j; } }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        u64 om = shfl_xor_u64(m, off); int op = __shfl_xor_sync(0xffffffffu, mp, off);
        if (om > m || (om == m && op < mp)) { m = om; mp = op; }
    }
    thr = m; pos = mp;
}

// block-wide bitonic sort (ascending) of P (power of two) 64-bit words in shared memory
__device__ __forceinline__ void block_bitonic_sort(u64* a, int P, int tid, int nthreads) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < (P >> 1); i += nthreads) {
                int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                int hi = lo | stride;
                bool up = (lo & size) == 0;
                u64 x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// ---- K1 parameters (kernels in scan_kernels.cuh) -------------------------------------------------
struct ScanParams {
    const unsigned char* rows;   // resident shard, row pitch `pitch` bytes (multiple of 16, zero padded)
    const float* norm;           // sqrt(sum x^2) per row (cosine)
    const uint32_t* mask;        // optional row mask, bit (row & 31) of word row >> 5 set = row excluded (tombstones / filters)
    const void* queries;         // nq_total x d, dtype qdtype, device
    u64* cand;                   // [nq_launch][gridDim.x][k] sorted per-CTA lists
    uint64_t n;                  // rows in the shard
    uint32_t d;                  // components per row
    uint32_t pitch;              // bytes
    uint32_t k;
    uint32_t P;                  // pow2 >= kScanWarps * k (per-query list region, 64-bit words)
    int metric;                  // M_L2 / M_IP / M_COS
    int qdtype;                  // DT_F32 or the space's dtype
    uint32_t q0;                 // first query of this launch
};

// per-query constants computed in the prologue
struct QueryConst { float qnorm; int qq; };

// ------------------------------------------------------------------------------------------------
// K3 -- merge_topk: one CTA per query merges `nlists` lists of k candidates; optionally decodes.
// ------------------------------------------------------------------------------------------------
struct MergeParams {
    const u64* in;            // candidate (q, list, j) at in[q*q_stride + list*list_stride + j]
    size_t q_stride, list_stride;
    uint32_t nlists, k;
    uint32_t P;               // pow2 sort buffer (64-bit words), > k (bitonic flavour)
    uint32_t seg_lists;       // lists per CTA (tournament flavour)
    uint64_t id_base;         // first GLOBAL row of the shard the candidates' 32-bit LOCAL row ids count from (used when ids are emitted)
    const uint64_t* list_base = nullptr;   // cross-rank merge (list = rank): first global row of every rank's shard.  Candidates keep
                              // LOCAL ids all the way, so only a shard (not the space) is limited to 2^32 rows; the merge orders
                              // equal keys by (rank, position in the rank's sorted list) = by global row (shards are contiguous)
    u64* out_packed;          // [nq][k] or null
    uint64_t* out_ids;        // [nq][k] or null
    float* out_scores;        // [nq][k] or null
    int* out_iscores;         // [nq][k] or null
    int metric, dtype;
    int tc_mode;              // keys hold K2's ranking value (tc_kernels.cuh); decode with qconst
    const float* qconst;      // per query: 1/|q| (cosine), |q|^2 (L2; int32 bits for integer spaces)
    const uint64_t* ext_ids = nullptr;   // optional external id per GLOBAL row (vector_ids_block_index, schema/core.fbs:54): reported instead of the position
};

// Cross-rank merges compare (key, rank, position j in the rank's list) instead of (key, local id): the low word of a
// candidate is replaced on load and resolved back to a global row when the result is emitted.
__device__ __forceinline__ u64 rank_word(u64 w, uint32_t list, uint32_t j) {
    return w == kInf ? kInf : ((w & 0xffffffff00000000ull) | ((u64)list << 16) | j);
}
// decode one merged candidate into the caller's output arrays; gid = its global row
__device__ __forceinline__ void emit_result_gid(const MergeParams& p, uint32_t q, uint32_t j, u64 v, uint64_t gid) {
    const bool valid = v != kInf;
    const size_t o = (size_t)q * p.k + j;
    if (p.out_packed) p.out_packed[o] = v;
    const uint32_t key = (uint32_t)(v >> 32);
    if (p.out_ids) p.out_ids[o] = valid ? (p.ext_ids ? p.ext_ids[gid] : gid) : ~0ull;
    const bool int_score = (p.dtype == DT_I8 || p.dtype == DT_U8) && p.metric != M_COS;
    if (int_score) {
        int s = (int)(key ^ 0x80000000u);
        if (p.tc_mode && p.metric == M_L2) s += __float_as_int(p.qconst[q]);      // |q|^2 + (|x|^2 - 2 q.x)
        if (p.metric == M_IP) s = -s;
        if (p.out_iscores) p.out_iscores[o] = valid ? s : 0x7fffffff;
        if (p.out_scores) p.out_scores[o] = valid ? (float)s : __uint_as_float(0x7fc00000u);
    } else {
        float s = unordered_f32(key);
        if (p.tc_mode) {
            const float c = p.qconst[q];
            if (p.metric == M_COS) s = fmaf(s, c, 1.0f);                           // 1 - q.x / (|q||x|)
            else if (p.metric == M_L2) s = __fsqrt_rn(fmaxf(c + s, 0.0f));          // sqrt(|q|^2 + |x|^2 - 2 q.x)
            else s = 0.0f - s;
        } else if (p.metric == M_IP) s = 0.0f - s;
        if (p.out_iscores) p.out_iscores[o] = 0;
        if (p.out_scores) p.out_scores[o] = valid ? s : __uint_as_float(0x7fc00000u);
    }
}
__device__ __forceinline__ void emit_result(const MergeParams& p, uint32_t q, uint32_t j, u64 v) {
    uint64_t gid = 0;
    if (v != kInf) {
        if (p.list_base != nullptr) {                          // ranked word: (list, position) -> the local id it stands for
            const uint32_t r = (uint32_t)(v >> 16) & 0xffffu, jj = (uint32_t)v & 0xffffu;
            gid = p.list_base[r] + (uint32_t)p.in[(size_t)q * p.q_stride + (size_t)r * p.list_stride + jj];
        } else {
            gid = p.id_base + (uint32_t)v;
        }
    }
    emit_result_gid(p, q, j, v, gid);
}

// K3, fast flavour: the lists are individually sorted, so the k best of their union come out of a
// two-level tournament.  All nlists*k words are staged in shared memory; each of the 32 warps
// merges its share of the lists (one or more list heads per lane, warp-min per output), then warp 0
// merges the 32 partial results.  ~100 cycles per output instead of a 16K-element sort.
constexpr int kMergeMaxHeads = 4;   // list heads per lane -> up to 32*32*4 lists

// minimum of one 64-bit word per lane: two hardware warp reductions (REDUX) on the halves instead of a five-round
// 64-bit shuffle butterfly -- every output of every tournament below costs one of these
__device__ __forceinline__ u64 warp_min_u64(u64 v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((u64)mh << 32) | ml;
}

// one warp merges `nl` sorted lists (list i at base + i*stride, k words each) into out[0..k)
__device__ __forceinline__ void warp_tournament(const u64* base, size_t stride, uint32_t nl, uint32_t k, u64* out, int lane) {
    uint32_t idx[kMergeMaxHeads];
    u64 head[kMergeMaxHeads];
#pragma unroll
    for (int h = 0; h < kMergeMaxHeads; h++) {
        idx[h] = 0;
        const uint32_t li = (uint32_t)h * 32 + lane;
        head[h] = li < nl ? base[(size_t)li * stride] : kInf;
    }
    for (uint32_t j = 0; j < k; j++) {
        u64 mine = head[0];
#pragma unroll
        for (int h = 1; h < kMergeMaxHeads; h++) mine = head[h] < mine ? head[h] : mine;
        const u64 m = warp_min_u64(mine);
        if (m == kInf) {                                   // every list is exhausted
            for (uint32_t t = j + lane; t < k; t += 32) out[t] = kInf;
            break;
        }
        const unsigned owners = __ballot_sync(0xffffffffu, mine == m);
        if (lane == __ffs(owners) - 1) {
            out[j] = m;
            bool adv = false;
#pragma unroll
            for (int h = 0; h < kMergeMaxHeads; h++) {
                if (!adv && head[h] == m) {
                    adv = true;
                    const uint32_t li = (uint32_t)h * 32 + lane;
                    idx[h]++;
                    head[h] = idx[h] < k ? base[(size_t)li * stride + idx[h]] : kInf;
                }
            }
        }
    }
    __syncwarp();
}

// X1 parameters (exchange_merge below; also the tail of a fused scan launch, scan_kernels.cuh)
struct ExchangeParams {
    const u64* mine;          // [nq][k] this rank's merged candidates (row ids local to its shard)
    uint4* const* peers;      // [world] window of rank r as mapped into this device (peers[rank] = window)
    uint4* window;            // this rank's own window
    uint32_t world, rank;
    uint32_t slot_cells;      // cells per (buffer, rank) slot, >= nq * k
    uint32_t buf, flag;       // buffer of this search (0/1) and its flag value (never 0, differs between consecutive uses of a buffer)
    uint32_t* error;          // mapped pinned word: set when a peer's cells did not arrive within the time limit
    int do_merge;             // 0: push and wait only (members of a one-process group other than the first)
};
__device__ __forceinline__ void st_cell(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_cell(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

// every rank's k words of query q (cells written by exchange_push of all ranks) from this rank's window into `cand`
// ([world][k]); polls until the cells carry this search's flag
__device__ __forceinline__ void exchange_gather(const ExchangeParams& x, uint32_t q, uint32_t k, u64* cand, uint32_t tid, uint32_t nthreads) {
    const size_t slot0 = (size_t)x.buf * x.world * x.slot_cells;
    const unsigned long long t0 = globaltimer_ns();
    for (uint32_t i = tid; i < x.world * k; i += nthreads) {
        const uint32_t r = i / k, j = i - r * k;
        const uint4* cell = x.window + slot0 + (size_t)r * x.slot_cells + (size_t)q * k + j;
        uint4 c = ld_cell(cell);
        while (c.y != x.flag || c.w != x.flag) {
            if (globaltimer_ns() - t0 > 20000000000ull) { *x.error = 1u; c = make_uint4(~0u, 0, ~0u, 0); break; }   // a peer is gone: fail, do not hang
            __nanosleep(64);
            c = ld_cell(cell);
        }
        cand[(size_t)r * k + j] = rank_word(((u64)c.z << 32) | c.x, r, j);   // cross-rank order: (key, rank, position)
    }
}
// the global row a ranked word of query q stands for: local id from the cell it came from + that rank's first row
__device__ __forceinline__ uint64_t exchange_gid(const ExchangeParams& x, const uint64_t* list_base, uint32_t q, uint32_t k, u64 w) {
    if (w == kInf) return 0;
    const uint32_t r = (uint32_t)(w >> 16) & 0xffffu, j = (uint32_t)w & 0xffffu;
    const size_t slot0 = (size_t)x.buf * x.world * x.slot_cells;
    return list_base[r] + ld_cell(x.window + slot0 + (size_t)r * x.slot_cells + (size_t)q * k + j).x;
}
// word j of query q of this rank into slot `rank` of every rank's window
__device__ __forceinline__ void exchange_push(const ExchangeParams& x, uint32_t q, uint32_t k, uint32_t j, uint32_t r, u64 w) {
    const size_t slot0 = (size_t)x.buf * x.world * x.slot_cells;
    st_cell(x.peers[r] + slot0 + (size_t)x.rank * x.slot_cells + (size_t)q * k + j, make_uint4((uint32_t)w, x.flag, (uint32_t)(w >> 32), x.flag));
}

// K1 + K3 in one launch: what the scan kernels need to merge their own per-CTA lists (the last CTA to finish does it)
struct ScanTail {
    uint32_t* counter;        // CTAs finished so far (zero between launches); nullptr = no fused merge, K3 is launched separately
    MergeParams mp;           // in / q_stride / nlists describe ScanParams::cand of the same launch
    int exchange;             // K1 + K3 + X1 in one launch: the last CTA also pushes the merged list to every rank, waits for theirs
    ExchangeParams x;         // and merges them (mp then describes the FINAL outputs, mp.list_base the ranks' first rows)
};

#ifdef MVF_ENGINE_TU   // non-template kernels: defined in engine.cu's translation unit only
__global__ void __launch_bounds__(1024) merge_topk(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    u64* buf = reinterpret_cast<u64*>(smem);
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const size_t total = (size_t)p.nlists * p.k;
    const u64* src = p.in + (size_t)q * p.q_stride;
    for (uint32_t i = tid; i < p.k; i += blockDim.x) buf[i] = kInf;
    size_t done = 0;
    do {
        const uint32_t room = p.P - p.k;
        for (uint32_t i = tid; i < room; i += blockDim.x) {
            size_t c = done + i;
            u64 v = kInf;
            if (c < total) {
                v = src[(c / p.k) * p.list_stride + (c % p.k)];
                if (p.list_base != nullptr) v = rank_word(v, (uint32_t)(c / p.k), (uint32_t)(c % p.k));
            }
            buf[p.k + i] = v;
        }
        done += room;
        block_bitonic_sort(buf, (int)p.P, (int)tid, (int)blockDim.x);
    } while (done < total);
    for (uint32_t j = tid; j < p.k; j += blockDim.x) emit_result(p, q, j, buf[j]);
}

__global__ void __launch_bounds__(1024) merge_topk_tournament(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    // gridDim.y > 1: segmented pass -- CTA (q, seg) merges lists [seg*seg_lists, ...) into one
    // intermediate sorted list at out_packed[(q*gridDim.y + seg)*k]; no decode
    const uint32_t q = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t first = seg * p.seg_lists;
    const uint32_t nlists = first < p.nlists ? min(p.seg_lists, p.nlists - first) : 0;
    u64* cand = reinterpret_cast<u64*>(smem);               // [nlists][k]
    u64* lvl1 = cand + (size_t)p.seg_lists * p.k;           // [nwarps][k]
    const u64* src = p.in + (size_t)q * p.q_stride + (size_t)first * p.list_stride;
    const size_t total = (size_t)nlists * p.k;
    for (size_t c = tid; c < total; c += blockDim.x) {
        u64 v = src[(c / p.k) * p.list_stride + (c % p.k)];
        if (p.list_base != nullptr) v = rank_word(v, first + (uint32_t)(c / p.k), (uint32_t)(c % p.k));   // (never segmented: engine.cu)
        cand[c] = v;
    }
    __syncthreads();
    // level 1: warp w owns lists [w*per, min((w+1)*per, nlists))
    const uint32_t per = (nlists + nwarps - 1) / nwarps;
    const uint32_t l0 = warp * per;
    const uint32_t nl = l0 < nlists ? min(per, nlists - l0) : 0;
    warp_tournament(cand + (size_t)l0 * p.k, p.k, nl, p.k, lvl1 + (size_t)warp * p.k, (int)lane);
    __syncthreads();
    if (warp == 0) {
        u64* out = cand;                                    // reuse: level-1 input is dead
        warp_tournament(lvl1, p.k, nwarps, p.k, out, (int)lane);
        if (gridDim.y > 1) {
            u64* dst = p.out_packed + ((size_t)q * gridDim.y + seg) * p.k;
            for (uint32_t j = lane; j < p.k; j += 32) dst[j] = out[j];
        } else {
            for (uint32_t j = lane; j < p.k; j += 32) emit_result(p, q, j, out[j]);
        }
    }
}


// K3, batch flavour: one WARP per query, list heads read straight from global memory (L2) with the next
// element of every list prefetched one step ahead.  With thousands of queries every SM holds dozens of
// independent merges, so the per-output latency is hidden by other warps instead of by shared-memory
// staging (the CTA-per-query tournament above is the low-latency choice for a handful of queries).
__global__ void __launch_bounds__(256) merge_topk_warp(const MergeParams p, uint32_t nq) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const u64* src = p.in + (size_t)q * p.q_stride;
    uint32_t idx[kMergeMaxHeads];
    u64 head[kMergeMaxHeads], nxt[kMergeMaxHeads];
#pragma unroll
    for (int h = 0; h < kMergeMaxHeads; h++) {
        idx[h] = 0;
        const uint32_t li = (uint32_t)h * 32 + lane;
        head[h] = li < p.nlists ? __ldcg(src + (size_t)li * p.list_stride) : kInf;
        nxt[h] = (li < p.nlists && p.k > 1) ? __ldcg(src + (size_t)li * p.list_stride + 1) : kInf;
        if (p.list_base != nullptr) { head[h] = rank_word(head[h], li, 0); nxt[h] = rank_word(nxt[h], li, 1); }
    }
    for (uint32_t j = 0; j < p.k; j++) {
        u64 mine = head[0];
#pragma unroll
        for (int h = 1; h < kMergeMaxHeads; h++) mine = head[h] < mine ? head[h] : mine;
        const u64 m = warp_min_u64(mine);
        if (lane == 0) emit_result(p, q, j, m);
        if (m == kInf) {                                   // every list is exhausted: pad the rest
            for (uint32_t t = j + 1 + lane; t < p.k; t += 32) emit_result(p, q, t, kInf);
            break;
        }
        const unsigned owners = __ballot_sync(0xffffffffu, mine == m);
        if (lane == (uint32_t)(__ffs(owners) - 1)) {
            bool adv = false;
#pragma unroll
            for (int h = 0; h < kMergeMaxHeads; h++) {
                if (!adv && head[h] == m) {
                    adv = true;
                    const uint32_t li = (uint32_t)h * 32 + lane;
                    idx[h]++;
                    head[h] = idx[h] < p.k ? nxt[h] : kInf;
                    nxt[h] = idx[h] + 1 < p.k ? __ldcg(src + (size_t)li * p.list_stride + idx[h] + 1) : kInf;
                    if (p.list_base != nullptr) nxt[h] = rank_word(nxt[h], li, idx[h] + 1);
                }
            }
        }
    }
}
// ------------------------------------------------------------------------------------------------
// X1 -- exchange_merge: the cross-GPU step of a sharded search as ONE kernel over peer memory (NVLink / NVSwitch
// loads and stores), instead of ncclAllGather + K3.  Every rank owns a window other ranks can write
// (cudaIpc* between processes, peer access inside one process): [2 buffers][world ranks][slot_cells] cells of 16
// bytes {lo32, flag, hi32, flag}.  CTA q pushes this rank's k candidates of query q into slot `rank` of EVERY
// rank's window, data and flag in the same 16-byte store (each 8-byte half carries its own copy of the flag, the
// way NCCL's LL protocol does, so no fence and no separate flag round trip is needed), then polls its own window
// until the k cells of every rank carry this search's flag, and merges the `world` sorted lists.  Two buffers
// alternate: a rank can start pushing search e+1 only after its own merge of search e, and it can finish the merge of
// e+1 only after everybody pushed e+1, so nobody is ever two searches ahead of a reader.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) exchange_merge(const ExchangeParams x, const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    u64* cand = reinterpret_cast<u64*>(smem);               // [world][k], then [k] merged
    const uint32_t q = blockIdx.x, tid = threadIdx.x, k = p.k, W = x.world;
    for (uint32_t i = tid; i < W * k; i += blockDim.x) {
        const uint32_t r = i / k, j = i - r * k;
        exchange_push(x, q, k, j, r, x.mine[(size_t)q * k + j]);
    }
    exchange_gather(x, q, k, cand, tid, blockDim.x);
    __syncthreads();
    if (!x.do_merge || tid >= 32) return;
    u64* out = cand + (size_t)W * k;
    warp_tournament(cand, k, W, k, out, (int)tid);
    for (uint32_t j = tid; j < k; j += 32) emit_result_gid(p, q, j, out[j], exchange_gid(x, p.list_base, q, k, out[j]));
}
#endif  // MVF_ENGINE_TU

// ------------------------------------------------------------------------------------------------
// K0 -- row norms (once per resident space): norm = sqrt(sum x^2) (f32), sq = sum x^2 (f32, or the
// exact int32 bit pattern for integer spaces).  One warp per row, 16-byte loads.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) row_norms(const unsigned char* rows, uint64_t n, uint32_t pitch, float* norm,
                                                  float* sq, unsigned long long* nonfinite) {
    constexpr int EPC = Traits<DT>::EPC;
    const int lane = threadIdx.x & 31;
    const uint32_t C = pitch >> 4;
    for (uint64_t row = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (uint64_t)gridDim.x * 8) {
        const unsigned char* rp = rows + row * pitch;
        float fs = 0.0f; int is = 0;
        for (uint32_t ci = lane; ci < C; ci += 32) {
            uint4 x = ldg_stream16(rp + (size_t)ci * 16);
            if (Traits<DT>::is_float) {
                float xf[EPC]; widen<DT>(x, xf);
#pragma unroll
                for (int e = 0; e < EPC; e++) fs = fmaf(xf[e], xf[e], fs);
            } else {
                is = dp4<DT>(x.x, x.x, is); is = dp4<DT>(x.y, x.y, is); is = dp4<DT>(x.z, x.z, is); is = dp4<DT>(x.w, x.w, is);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            fs += __shfl_xor_sync(0xffffffffu, fs, off);
            is += __shfl_xor_sync(0xffffffffu, is, off);
        }
        if (lane == 0) {
            if (Traits<DT>::is_float) {
                norm[row] = __fsqrt_rn(fs); sq[row] = fs;
                // rows holding Inf/NaN (or whose |x|^2 overflows f32): the space then stays on the scan kernel, whose
                // element-wise arithmetic defines the result for such rows (DESIGN.md "non-finite values")
                if (nonfinite != nullptr && !(fabsf(fs) <= 3.402823466e38f)) atomicAdd(nonfinite, 1ull);
            }
            else { norm[row] = __fsqrt_rn((float)is); reinterpret_cast<int*>(sq)[row] = is; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// crc32_resident -- CRC-32 (IEEE 802.3, what crc32fast::hash writes into the block manifest,
// src/builder.rs:251) of the resident rows, the check the reference's validate_with_checksum
// (src/reader.rs:172-221) never finished.  Every thread hashes one contiguous segment with
// slicing-by-8 tables in shared memory; a segment CRC is moved to its place in the whole message by
// multiplying with x^(8 * bytes_after) mod P (the zlib crc32_combine identity), and the
// contributions are XOR-ed together.
// ------------------------------------------------------------------------------------------------
struct CrcParams {
    const unsigned char* data; uint64_t nbytes; uint64_t seg_len;   // message bytes; bytes per segment (thread)
    uint64_t seg_stride;         // bytes between segment starts in memory (== seg_len for contiguous rows, the row pitch otherwise); multiple of 16
    const uint32_t* tables;      // [8][256] slicing tables
    const uint32_t* x2n;         // [64] x^(2^k) mod P, reflected representation
    uint32_t* out;               // XOR accumulator, zeroed by the caller
};
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        if (a & (1u << i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t crc_step8(const uint32_t (*T)[256], uint32_t crc, uint32_t a, uint32_t b) {
    a ^= crc;
    return T[7][a & 0xff] ^ T[6][(a >> 8) & 0xff] ^ T[5][(a >> 16) & 0xff] ^ T[4][a >> 24] ^
           T[3][b & 0xff] ^ T[2][(b >> 8) & 0xff] ^ T[1][(b >> 16) & 0xff] ^ T[0][b >> 24];
}
#ifdef MVF_ENGINE_TU
// K2 integer L2: what every accumulator column starts from (tc_kernels.cuh INIT_ACC): -(|x|^2 >> 1) for live rows, `dead`
// for rows past the end of the shard (the array is padded to whole tiles), deleted or filtered out.  Rebuilt when the
// row mask changes.
__global__ void __launch_bounds__(256) build_acc_seed(const float* sq, const uint32_t* mask, uint64_t n, uint64_t padded, int dead, int* out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool ok = i < n && (mask == nullptr || ((mask[i >> 5] >> (i & 31)) & 1u) == 0);
        out[i] = ok ? -(__float_as_int(sq[i]) >> 1) : dead;
    }
}
// sets *flag when any of `count` f16 / f32 values is Inf or NaN (exponent all ones): such query batches stay on the
// scan kernel (DESIGN.md "non-finite values")
__global__ void __launch_bounds__(256) flag_nonfinite(const void* data, uint64_t count, int is_f16, uint32_t* flag) {
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        if (is_f16) bad |= (reinterpret_cast<const unsigned short*>(data)[i] & 0x7c00u) == 0x7c00u;
        else bad |= (reinterpret_cast<const uint32_t*>(data)[i] & 0x7f800000u) == 0x7f800000u;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
__global__ void __launch_bounds__(256) crc32_resident(const CrcParams p) {
    __shared__ uint32_t T[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&T[0][0])[i] = p.tables[i];
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t off = t * p.seg_len;
    uint32_t contrib = 0;
    if (off < p.nbytes) {
        const uint64_t len = min(p.seg_len, p.nbytes - off);
        const unsigned char* s = p.data + t * p.seg_stride;
        uint32_t crc = 0xffffffffu;
        uint64_t i = 0;
        for (; i + 128 <= len; i += 128) {                  // one 128-byte line per iteration, 8 loads in flight
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = *reinterpret_cast<const uint4*>(s + i + 16 * j);
#pragma unroll
            for (int j = 0; j < 8; j++) { crc = crc_step8(T, crc, v[j].x, v[j].y); crc = crc_step8(T, crc, v[j].z, v[j].w); }
        }
        for (; i + 16 <= len; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(s + i);
            crc = crc_step8(T, crc, v.x, v.y); crc = crc_step8(T, crc, v.z, v.w);
        }
        for (; i < len; i++) crc = T[0][(crc ^ s[i]) & 0xff] ^ (crc >> 8);
        crc = ~crc;
        // weight x^(8 * bytes after this segment)
        uint64_t e = 8 * (p.nbytes - off - len);
        uint32_t w = 0x80000000u;                            // x^0
        for (int kbit = 0; e; e >>= 1, kbit++) if (e & 1) w = gf2_mulmod(p.x2n[kbit], w);
        contrib = gf2_mulmod(w, crc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib ^= __shfl_xor_sync(0xffffffffu, contrib, o);
    if ((threadIdx.x & 31) == 0 && contrib) atomicXor(p.out, contrib);
}
#endif

// ------------------------------------------------------------------------------------------------
// gather_rows -- the resident rows of a list of ids (the `vector` field of ScoredVector,
// examples/similarity_search.rs:18; VectorSpace::get_vectors_batch, src/vectors/vector_space.rs:230-243),
// raw or widened to f32 like Vector::as_f32.  One warp per id, 16-byte loads.
// ------------------------------------------------------------------------------------------------
template <int DT, bool AS_F32>
__global__ void __launch_bounds__(256) gather_rows(const unsigned char* rows, uint32_t pitch, uint32_t d, uint64_t n,
                                                    uint64_t row_base, const uint64_t* ids, uint32_t count, unsigned char* out) {
    constexpr int EPC = Traits<DT>::EPC;
    constexpr int ES = 16 / EPC;
    const int lane = threadIdx.x & 31;
    const uint32_t C = pitch >> 4;
    for (uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5); i < count; i += gridDim.x * 8) {
        const uint64_t gid = ids[i];
        const bool ok = gid >= row_base && gid - row_base < n;          // ids outside this shard (or invalid) give zero rows
        const unsigned char* rp = rows + (ok ? gid - row_base : 0) * pitch;
        for (uint32_t ci = lane; ci < C; ci += 32) {
            uint4 x = ok ? *reinterpret_cast<const uint4*>(rp + (size_t)ci * 16) : make_uint4(0, 0, 0, 0);
            if (AS_F32) {
                float xf[EPC];
                widen<DT>(x, xf);
                float* o = reinterpret_cast<float*>(out) + (size_t)i * d + (size_t)ci * EPC;
#pragma unroll
                for (int e = 0; e < EPC; e++) if (ci * EPC + e < d) o[e] = xf[e];
            } else {
                unsigned char* o = out + (size_t)i * d * ES + (size_t)ci * 16;
                const unsigned char* xb = reinterpret_cast<const unsigned char*>(&x);
#pragma unroll
                for (int e = 0; e < 16; e++) if ((size_t)ci * 16 + e < (size_t)d * ES) o[e] = xb[e];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// synth_fill -- one thread per 16-byte chunk of the (zero-padded) resident layout
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) synth_fill(unsigned char* rows, uint64_t n, uint32_t d, uint32_t pitch,
                                                   uint64_t seed_key, uint64_t row_base) {
    constexpr int EPC = Traits<DT>::EPC;
    const uint32_t C = pitch >> 4;
    const uint64_t total = n * C;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = t / C; const uint32_t ci = (uint32_t)(t % C);
        const uint64_t rk = synth::row_key(seed_key, row_base + row);
        const uint32_t c0 = ci * EPC;
        uint32_t w[4] = {0, 0, 0, 0};
        if (Traits<DT>::is_float) {
            float v[EPC];
#pragma unroll
            for (int e = 0; e < EPC; e += 2) synth::float_pair(rk, (c0 + e) >> 1, &v[e], &v[e + 1]);
#pragma unroll
            for (int e = 0; e < EPC; e++) if (c0 + e >= d) v[e] = 0.0f;
            if (DT == DT_F32) {
#pragma unroll
                for (int e = 0; e < 4; e++) w[e] = __float_as_uint(v[e % EPC]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    __half2 h = __floats2half2_rn(v[(2 * e) % EPC], v[(2 * e + 1) % EPC]);
                    w[e] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
        } else {
            uint64_t a = synth::byte_octet(rk, c0 >> 3), b = synth::byte_octet(rk, (c0 >> 3) + 1);
            w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32);
            unsigned char* bytes = reinterpret_cast<unsigned char*>(w);
#pragma unroll
            for (int e = 0; e < 16; e++) if (c0 + e >= d) bytes[e] = 0;
        }
        *reinterpret_cast<uint4*>(rows + row * pitch + (size_t)ci * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace k
}  // namespace mvf
