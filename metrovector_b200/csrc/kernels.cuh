// kernels.cuh -- hand-written sm_100a kernels of the bandwidth path.
//   K0  row_norms        sqrt(sum x^2) and sum x^2 per row, once per resident space
//   K1  scan_topk        distance + running top-k, one pass over the HBM-resident rows
//   K3  merge_topk       per-query merge of candidate lists (per-CTA lists, or per-rank lists)
//   synth_fill           device-side synthetic rows (bench/tests)
// What they replace: the loop body of find_top_k_similar, examples/similarity_search.rs:147-169
// of the reference (get_vector -> as_f32 -> sum of squared differences -> heap push/pop).
// Candidates are 64-bit words (order-preserving 32-bit key << 32 | 32-bit row id), so "smallest
// score, lowest index on ties" is plain unsigned order everywhere.
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

// ------------------------------------------------------------------------------------------------
// K1 -- scan_topk
// ------------------------------------------------------------------------------------------------
struct ScanParams {
    const unsigned char* rows;   // resident shard, row pitch `pitch` bytes (multiple of 16, zero padded)
    const float* norm;           // sqrt(sum x^2) per row (cosine)
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

template <int DT>
__device__ __forceinline__ u64 finalize(int metric, float facc, int iacc, int ixx, const QueryConst& qc, float xnorm,
                                        uint32_t row) {
    uint32_t key;
    if (Traits<DT>::is_float) {
        float v;
        if (metric == M_L2) v = __fsqrt_rn(facc);
        else if (metric == M_IP) v = 0.0f - facc;
        else { float den = __fmul_rn(qc.qnorm, xnorm); v = den > 0.0f ? __fsub_rn(1.0f, __fdiv_rn(facc, den)) : 1.0f; }
        key = ordered_f32(v);
    } else {
        if (metric == M_L2) key = ordered_i32(qc.qq + ixx - 2 * iacc);
        else if (metric == M_IP) key = ordered_i32(-iacc);
        else {
            float den = __fmul_rn(qc.qnorm, xnorm);
            float v = den > 0.0f ? __fsub_rn(1.0f, __fdiv_rn((float)iacc, den)) : 1.0f;
            key = ordered_f32(v);
        }
    }
    return ((u64)key << 32) | row;
}

// Shared prologue: stage NQ queries in shared memory (float spaces: widened to f32, chunk-split so
// that lane-strided 16-byte reads are conflict-free; integer spaces: raw bytes), zero padded to
// dpad components, and compute per-query constants.
template <int DT, int NQ>
__device__ __forceinline__ void stage_queries(const ScanParams& p, unsigned char* q_s, uint32_t dpad, QueryConst* qc_s,
                                              float* red_s) {
    constexpr int EPC = Traits<DT>::EPC;
    const int tid = threadIdx.x;
    for (int q = 0; q < NQ; q++) {
        float fs = 0.0f; int is = 0;
        if (Traits<DT>::is_float) {
            float* dst = reinterpret_cast<float*>(q_s) + (size_t)q * dpad;
            for (uint32_t j = tid; j < dpad; j += blockDim.x) {
                float v = 0.0f;
                if (j < p.d) {
                    size_t qi = (size_t)(p.q0 + q) * p.d + j;
                    v = p.qdtype == DT_F16 ? __half2float(reinterpret_cast<const __half*>(p.queries)[qi])
                                           : reinterpret_cast<const float*>(p.queries)[qi];
                }
                uint32_t ci = j / EPC, e = j % EPC;
                // f16 spaces: elements 0-3 of every chunk in the first half-plane, 4-7 in the second
                uint32_t at = (EPC == 8) ? ((e >> 2) * (dpad / 2) + ci * 4 + (e & 3)) : j;
                dst[at] = v;
                fs += v * v;
            }
        } else {
            unsigned char* dst = q_s + (size_t)q * dpad;
            for (uint32_t j = tid; j < dpad; j += blockDim.x) {
                unsigned char b = 0;
                if (j < p.d) b = reinterpret_cast<const unsigned char*>(p.queries)[(size_t)(p.q0 + q) * p.d + j];
                dst[j] = b;
                int v = DT == DT_I8 ? (int)(signed char)b : (int)b;
                is += v * v;
            }
        }
        // block reduction of the norm
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            fs += __shfl_xor_sync(0xffffffffu, fs, off);
            is += __shfl_xor_sync(0xffffffffu, is, off);
        }
        __syncthreads();
        if ((tid & 31) == 0) { red_s[tid >> 5] = fs; reinterpret_cast<int*>(red_s)[kScanWarps + (tid >> 5)] = is; }
        __syncthreads();
        if (tid == 0) {
            float f = 0.0f; int i = 0;
            for (int w = 0; w < kScanWarps; w++) { f += red_s[w]; i += reinterpret_cast<int*>(red_s)[kScanWarps + w]; }
            qc_s[q].qq = i;
            qc_s[q].qnorm = Traits<DT>::is_float ? __fsqrt_rn(f) : __fsqrt_rn((float)i);
        }
    }
    __syncthreads();
}

// accumulate one 16-byte chunk of one row against one query chunk
template <int DT, int OP>
__device__ __forceinline__ void accum_float(const float* xf, const float* qf, float& acc) {
    constexpr int EPC = Traits<DT>::EPC;
#pragma unroll
    for (int e = 0; e < EPC; e++) {
        if (OP == OP_L2) { float t = qf[e] - xf[e]; acc = fmaf(t, t, acc); }
        else acc = fmaf(qf[e], xf[e], acc);
    }
}

// CTA epilogue shared by both scan shapes: sort each query's kScanWarps*k candidates, emit best k
template <int NQ>
__device__ __forceinline__ void emit_cta_lists(const ScanParams& p, u64* lists) {
    __syncthreads();
    for (int q = 0; q < NQ; q++) {
        u64* l = lists + (size_t)q * p.P;
        block_bitonic_sort(l, (int)p.P, threadIdx.x, blockDim.x);
        u64* out = p.cand + ((size_t)q * gridDim.x + blockIdx.x) * p.k;
        for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = l[j];
    }
}

// WIDE rows (>= 32 chunks of 16 bytes): one warp per row, R rows in flight per warp.
template <int DT, int OP, int NQ, int R>
__global__ void __launch_bounds__(kScanThreads) scan_topk_wide(const ScanParams p) {
    constexpr int EPC = Traits<DT>::EPC;
    constexpr bool F = Traits<DT>::is_float;
    extern __shared__ __align__(16) unsigned char smem[];
    u64* lists = reinterpret_cast<u64*>(smem);
    const uint32_t C = p.pitch >> 4, CPL = (C + 31) >> 5, dpad = CPL * 32 * EPC;
    unsigned char* q_s = smem + (size_t)NQ * p.P * 8;
    const size_t q_bytes = (size_t)NQ * dpad * (F ? 4 : 1);
    QueryConst* qc_s = reinterpret_cast<QueryConst*>(q_s + ((q_bytes + 15) & ~(size_t)15));
    float* red_s = reinterpret_cast<float*>(qc_s + NQ);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (uint32_t i = threadIdx.x; i < NQ * p.P; i += blockDim.x) lists[i] = kInf;
    stage_queries<DT, NQ>(p, q_s, dpad, qc_s, red_s);

    u64 thr[NQ]; int pos[NQ]; QueryConst qc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) { thr[q] = kInf; pos[q] = 0; qc[q] = qc_s[q]; }

    const uint64_t groups = (p.n + R - 1) / R;
    for (uint64_t g = (uint64_t)blockIdx.x * kScanWarps + warp; g < groups; g += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t row0 = g * R;
        const unsigned char* rp[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint64_t row = row0 + r < p.n ? row0 + r : p.n - 1;
            rp[r] = p.rows + row * p.pitch + (size_t)lane * 16;
        }
        float xn = 0.0f;
        if (p.metric == M_COS && lane < R) xn = p.norm[row0 + lane < p.n ? row0 + lane : p.n - 1];
        float facc[NQ][R]; int iacc[NQ][R]; int ixx[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            ixx[r] = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++) { facc[q][r] = 0.0f; iacc[q][r] = 0; }
        }
        for (uint32_t c = 0; c < CPL; c++) {
            const uint32_t ci = c * 32 + lane;
            uint4 x[R];
            if (ci < C) {
#pragma unroll
                for (int r = 0; r < R; r++) x[r] = ldg_stream16(rp[r] + (size_t)c * 512);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) x[r] = make_uint4(0, 0, 0, 0);
            }
            if (F) {
                float xf[R][EPC];
#pragma unroll
                for (int r = 0; r < R; r++) widen<DT>(x[r], xf[r]);
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    float qf[EPC];
                    const float* qb = reinterpret_cast<const float*>(q_s) + (size_t)q * dpad;
                    float4 a = reinterpret_cast<const float4*>(qb)[ci];
                    qf[0] = a.x; qf[1] = a.y; qf[2] = a.z; qf[3] = a.w;
                    if (EPC == 8) {
                        float4 b = reinterpret_cast<const float4*>(qb + dpad / 2)[ci];
                        qf[4] = b.x; qf[5] = b.y; qf[6] = b.z; qf[7] = b.w;
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) accum_float<DT, OP>(xf[r], qf, facc[q][r]);
                }
            } else {
                if (OP == OP_L2) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        ixx[r] = dp4<DT>(x[r].x, x[r].x, ixx[r]); ixx[r] = dp4<DT>(x[r].y, x[r].y, ixx[r]);
                        ixx[r] = dp4<DT>(x[r].z, x[r].z, ixx[r]); ixx[r] = dp4<DT>(x[r].w, x[r].w, ixx[r]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    uint4 qv = reinterpret_cast<const uint4*>(q_s + (size_t)q * dpad)[ci];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        int a = iacc[q][r];
                        a = dp4<DT>(qv.x, x[r].x, a); a = dp4<DT>(qv.y, x[r].y, a);
                        a = dp4<DT>(qv.z, x[r].z, a); a = dp4<DT>(qv.w, x[r].w, a);
                        iacc[q][r] = a;
                    }
                }
            }
        }
        // warp reductions: every lane ends with the row totals
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                if (!F && OP == OP_L2) ixx[r] += __shfl_xor_sync(0xffffffffu, ixx[r], off);
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (F) facc[q][r] += __shfl_xor_sync(0xffffffffu, facc[q][r], off);
                    else iacc[q][r] += __shfl_xor_sync(0xffffffffu, iacc[q][r], off);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (row0 + r >= p.n) break;
            float xnr = __shfl_sync(0xffffffffu, xn, r);
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                u64 pk = finalize<DT>(p.metric, facc[q][r], iacc[q][r], ixx[r], qc[q], xnr, (uint32_t)(row0 + r));
                if (pk < thr[q]) list_insert(lists + (size_t)q * p.P + warp * p.k, (int)p.k, pk, thr[q], pos[q], lane);
            }
        }
    }
    emit_cta_lists<NQ>(p, lists);
}

// NARROW rows (<= 16 chunks): LPR = pow2 >= chunks lanes per row, 32/LPR rows per warp load,
// R loads in flight.  The lane's query chunk never changes, so it lives in registers.
template <int DT, int OP, int NQ, int R>
__global__ void __launch_bounds__(kScanThreads) scan_topk_narrow(const ScanParams p, const int LPR) {
    constexpr int EPC = Traits<DT>::EPC;
    constexpr bool F = Traits<DT>::is_float;
    extern __shared__ __align__(16) unsigned char smem[];
    u64* lists = reinterpret_cast<u64*>(smem);
    const uint32_t C = p.pitch >> 4, dpad = (uint32_t)LPR * EPC;
    unsigned char* q_s = smem + (size_t)NQ * p.P * 8;
    const size_t q_bytes = (size_t)NQ * dpad * (F ? 4 : 1);
    QueryConst* qc_s = reinterpret_cast<QueryConst*>(q_s + ((q_bytes + 15) & ~(size_t)15));
    float* red_s = reinterpret_cast<float*>(qc_s + NQ);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int li = lane & (LPR - 1), ri = lane / LPR, RPW = 32 / LPR;

    for (uint32_t i = threadIdx.x; i < NQ * p.P; i += blockDim.x) lists[i] = kInf;
    stage_queries<DT, NQ>(p, q_s, dpad, qc_s, red_s);

    u64 thr[NQ]; int pos[NQ]; QueryConst qc[NQ];
    float qf[NQ][F ? EPC : 1]; uint4 qv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        thr[q] = kInf; pos[q] = 0; qc[q] = qc_s[q];
        if (F) {
            const float* qb = reinterpret_cast<const float*>(q_s) + (size_t)q * dpad;
            float4 a = reinterpret_cast<const float4*>(qb)[li];
            qf[q][0] = a.x; qf[q][1] = a.y; qf[q][2] = a.z; qf[q][3] = a.w;
            if (EPC == 8) {
                float4 b = reinterpret_cast<const float4*>(qb + dpad / 2)[li];
                qf[q][4 % (F ? EPC : 1)] = b.x; qf[q][5 % (F ? EPC : 1)] = b.y;
                qf[q][6 % (F ? EPC : 1)] = b.z; qf[q][7 % (F ? EPC : 1)] = b.w;
            }
            qv[q] = make_uint4(0, 0, 0, 0);
        } else {
            qv[q] = reinterpret_cast<const uint4*>(q_s + (size_t)q * dpad)[li];
        }
    }
    const bool active = (uint32_t)li < C;
    const uint64_t rows_per_step = (uint64_t)R * RPW;
    const uint64_t groups = (p.n + rows_per_step - 1) / rows_per_step;
    for (uint64_t g = (uint64_t)blockIdx.x * kScanWarps + warp; g < groups; g += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t row0 = g * rows_per_step;
        uint4 x[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint64_t row = row0 + (uint64_t)r * RPW + ri;
            if (row >= p.n) row = p.n - 1;
            x[r] = active ? ldg_stream16(p.rows + row * p.pitch + (size_t)li * 16) : make_uint4(0, 0, 0, 0);
        }
        float xn[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            xn[r] = 0.0f;
            if (p.metric == M_COS) {
                uint64_t row = row0 + (uint64_t)r * RPW + ri;
                xn[r] = p.norm[row < p.n ? row : p.n - 1];
            }
        }
        float facc[NQ][R]; int iacc[NQ][R]; int ixx[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            ixx[r] = 0;
            if (F) {
                float xf[EPC];
                widen<DT>(x[r], xf);
#pragma unroll
                for (int q = 0; q < NQ; q++) { facc[q][r] = 0.0f; iacc[q][r] = 0; accum_float<DT, OP>(xf, qf[q], facc[q][r]); }
            } else {
                if (OP == OP_L2) {
                    ixx[r] = dp4<DT>(x[r].x, x[r].x, ixx[r]); ixx[r] = dp4<DT>(x[r].y, x[r].y, ixx[r]);
                    ixx[r] = dp4<DT>(x[r].z, x[r].z, ixx[r]); ixx[r] = dp4<DT>(x[r].w, x[r].w, ixx[r]);
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    int a = 0;
                    a = dp4<DT>(qv[q].x, x[r].x, a); a = dp4<DT>(qv[q].y, x[r].y, a);
                    a = dp4<DT>(qv[q].z, x[r].z, a); a = dp4<DT>(qv[q].w, x[r].w, a);
                    iacc[q][r] = a; facc[q][r] = 0.0f;
                }
            }
        }
        for (int off = LPR >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (!F && OP == OP_L2) ixx[r] += __shfl_xor_sync(0xffffffffu, ixx[r], off);
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (F) facc[q][r] += __shfl_xor_sync(0xffffffffu, facc[q][r], off);
                    else iacc[q][r] += __shfl_xor_sync(0xffffffffu, iacc[q][r], off);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint64_t row = row0 + (uint64_t)r * RPW + ri;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                u64 pk = kInf;
                if (row < p.n && li == 0)
                    pk = finalize<DT>(p.metric, facc[q][r], iacc[q][r], ixx[r], qc[q], xn[r], (uint32_t)row);
                unsigned hits = __ballot_sync(0xffffffffu, pk < thr[q]);
                while (hits) {
                    int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    u64 c = shfl_u64(pk, src);
                    if (c < thr[q]) list_insert(lists + (size_t)q * p.P + warp * p.k, (int)p.k, c, thr[q], pos[q], lane);
                }
            }
        }
    }
    emit_cta_lists<NQ>(p, lists);
}

// ------------------------------------------------------------------------------------------------
// K3 -- merge_topk: one CTA per query merges `nlists` lists of k candidates; optionally decodes.
// ------------------------------------------------------------------------------------------------
struct MergeParams {
    const u64* in;            // candidate (q, list, j) at in[q*q_stride + list*list_stride + j]
    size_t q_stride, list_stride;
    uint32_t nlists, k;
    uint32_t P;               // pow2 sort buffer (64-bit words), > k (bitonic flavour)
    uint32_t seg_lists;       // lists per CTA (tournament flavour)
    uint32_t id_add;          // added to the low 32 bits of every valid candidate (shard row base)
    u64* out_packed;          // [nq][k] or null
    uint64_t* out_ids;        // [nq][k] or null
    float* out_scores;        // [nq][k] or null
    int* out_iscores;         // [nq][k] or null
    int metric, dtype;
};

// decode one merged candidate into the caller's output arrays
__device__ __forceinline__ void emit_result(const MergeParams& p, uint32_t q, uint32_t j, u64 v) {
    const bool valid = v != kInf;
    if (valid) v += p.id_add;
    const size_t o = (size_t)q * p.k + j;
    if (p.out_packed) p.out_packed[o] = v;
    const uint32_t key = (uint32_t)(v >> 32);
    if (p.out_ids) p.out_ids[o] = valid ? (uint64_t)(uint32_t)v : ~0ull;
    const bool int_score = (p.dtype == DT_I8 || p.dtype == DT_U8) && p.metric != M_COS;
    if (int_score) {
        int s = (int)(key ^ 0x80000000u);
        if (p.metric == M_IP) s = -s;
        if (p.out_iscores) p.out_iscores[o] = valid ? s : 0x7fffffff;
        if (p.out_scores) p.out_scores[o] = valid ? (float)s : __uint_as_float(0x7fc00000u);
    } else {
        float s = unordered_f32(key);
        if (p.metric == M_IP) s = 0.0f - s;
        if (p.out_iscores) p.out_iscores[o] = 0;
        if (p.out_scores) p.out_scores[o] = valid ? s : __uint_as_float(0x7fc00000u);
    }
}

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
            if (c < total) v = src[(c / p.k) * p.list_stride + (c % p.k)];
            buf[p.k + i] = v;
        }
        done += room;
        block_bitonic_sort(buf, (int)p.P, (int)tid, (int)blockDim.x);
    } while (done < total);
    for (uint32_t j = tid; j < p.k; j += blockDim.x) emit_result(p, q, j, buf[j]);
}

// K3, fast flavour: the lists are individually sorted, so the k best of their union come out of a
// two-level tournament.  All nlists*k words are staged in shared memory; each of the 32 warps
// merges its share of the lists (one or more list heads per lane, warp-min per output), then warp 0
// merges the 32 partial results.  ~100 cycles per output instead of a 16K-element sort.
constexpr int kMergeMaxHeads = 4;   // list heads per lane -> up to 32*32*4 lists

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { u64 o = shfl_xor_u64(v, off); v = o < v ? o : v; }
    return v;
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
    for (size_t c = tid; c < total; c += blockDim.x) cand[c] = src[(c / p.k) * p.list_stride + (c % p.k)];
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

// ------------------------------------------------------------------------------------------------
// K0 -- row norms (once per resident space): norm = sqrt(sum x^2) (f32), sq = sum x^2 (f32, or the
// exact int32 bit pattern for integer spaces).  One warp per row, 16-byte loads.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) row_norms(const unsigned char* rows, uint64_t n, uint32_t pitch, float* norm,
                                                  float* sq) {
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
            if (Traits<DT>::is_float) { norm[row] = __fsqrt_rn(fs); sq[row] = fs; }
            else { norm[row] = __fsqrt_rn((float)is); reinterpret_cast<int*>(sq)[row] = is; }
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
