// scan_kernels.cuh -- K1, the bandwidth-bound scan with fused top-k (templates; instantiated per
// data type in scan_f32.cu / scan_f16.cu / scan_i8.cu / scan_u8.cu so the build parallelises).
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace mvf {
namespace k {

// ------------------------------------------------------------------------------------------------
// K1 -- scan_topk
// ------------------------------------------------------------------------------------------------
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

// CTA epilogue shared by both scan shapes: the best k of each query's kScanWarps*k candidates, ascending
// k <= 32 (register lists): the kScanWarps * k candidates of a query are ranked by counting -- thread t holds candidate t
// and counts the candidates that precede it (smaller word, or the same word at a lower position: the kInf fillers of a
// list that never filled up), then stores it at its rank if that is below k.  No sort, no tournament rounds, one block
// barrier between the warps' stores and the counting: the CTA epilogue is a fixed cost every search pays (it was a
// shuffle sort per warp plus k rounds of a warp tournament; a 256-row search took 12.3 us).
template <int NQ>
__device__ __forceinline__ void emit_cta_lists_reg(const ScanParams& p, u64* lists, const u64 (&regl)[NQ], int lane, int warp) {
    const uint32_t k = p.k, M = kScanWarps * k;
    __syncthreads();                                            // everybody is done with the query area / lists
#pragma unroll
    for (int q = 0; q < NQ; q++)
        if (lane < (int)k) lists[(size_t)q * p.P + warp * k + lane] = regl[q];
    __syncthreads();
    for (uint32_t idx = threadIdx.x; idx < NQ * M; idx += blockDim.x) {
        const uint32_t q = idx / M, t = idx - q * M;
        const u64* l = lists + (size_t)q * p.P;
        const u64 c = l[t];
        uint32_t rank = 0;
        for (uint32_t u = 0; u < M; u++) {
            const u64 o = l[u];
            rank += (o < c || (o == c && u < t)) ? 1u : 0u;
        }
        if (rank < k) p.cand[((size_t)q * gridDim.x + blockIdx.x) * k + rank] = c;
    }
}
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

// K1 + K3 fused (latency-bound configs: BASELINE configs[0] is 51 MB, 8 us of HBM time against two launches and a
// dependent grid start): every CTA counts itself done after its lists are visible; the last one stages all
// gridDim.x lists of a query in shared memory and runs the same two-level warp tournament as merge_topk_tournament,
// decoding straight into the caller's outputs.  The host only asks for this when the lists fit (engine.cu).
template <int NQ>
__device__ __forceinline__ void fused_merge_tail(const ScanParams& p, const ScanTail& tail, unsigned char* smem) {
    if (tail.counter == nullptr) return;
    __shared__ uint32_t s_last;
    __threadfence();                                        // this CTA's lists are visible before it is counted
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(tail.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *tail.counter = 0;                // zero again for the next launch on this stream
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nlists = gridDim.x, k = p.k;
    u64* cand_s = reinterpret_cast<u64*>(smem);             // [nlists][k]
    u64* lvl1 = cand_s + (size_t)nlists * k;                // [kScanWarps][k]
    const uint32_t per = (nlists + kScanWarps - 1) / kScanWarps;
    const uint32_t l0 = warp * per;
    const uint32_t nl = l0 < nlists ? min(per, nlists - l0) : 0;
    for (int q = 0; q < NQ; q++) {
        const u64* src = p.cand + (size_t)q * nlists * k;
        const uint32_t total = nlists * k;
        for (uint32_t c = threadIdx.x; c < total; c += blockDim.x) cand_s[c] = __ldcg(src + c);   // other CTAs wrote them: read at L2
        __syncthreads();
        warp_tournament(cand_s + (size_t)l0 * k, k, nl, k, lvl1 + (size_t)warp * k, (int)lane);
        __syncthreads();
        if (k <= 32) {
            // second level by counting (like the CTA epilogue): kScanWarps * k candidates, one per thread
            const uint32_t M = kScanWarps * k;
            for (uint32_t t = threadIdx.x; t < M; t += blockDim.x) {
                const u64 c = lvl1[t];
                uint32_t rank = 0;
                for (uint32_t u = 0; u < M; u++) {
                    const u64 o = lvl1[u];
                    rank += (o < c || (o == c && u < t)) ? 1u : 0u;
                }
                if (rank < k) cand_s[rank] = c;
            }
            __syncthreads();
        }
        if (warp == 0) {
            if (k > 32) warp_tournament(lvl1, k, kScanWarps, k, cand_s, (int)lane);
            if (!tail.exchange) {
                for (uint32_t j = lane; j < k; j += 32) emit_result(tail.mp, (uint32_t)q, j, cand_s[j]);
            } else {
                // sharded search, small payload: this rank's list goes straight into every rank's window (X1)
                for (uint32_t j = lane; j < k; j += 32) {
                    const u64 v = cand_s[j];                    // local row ids travel; the merge adds every rank's base
                    for (uint32_t r = 0; r < tail.x.world; r++) exchange_push(tail.x, p.q0 + (uint32_t)q, k, j, r, v);
                }
            }
        }
        __syncthreads();
        if (tail.exchange) {
            // ... and the lists of all ranks are merged here as soon as they have arrived: K1 + K3 + X1 in one launch
            u64* xc = cand_s;                               // [world][k] + [k] (the host sized the allocation for it)
            exchange_gather(tail.x, p.q0 + (uint32_t)q, k, xc, threadIdx.x, blockDim.x);
            __syncthreads();
            if (warp == 0 && tail.x.do_merge) {
                u64* out = xc + (size_t)tail.x.world * k;
                warp_tournament(xc, k, tail.x.world, k, out, (int)lane);
                for (uint32_t j = lane; j < k; j += 32)
                    emit_result_gid(tail.mp, (uint32_t)q, j, out[j], exchange_gid(tail.x, tail.mp.list_base, p.q0 + (uint32_t)q, k, out[j]));
            }
            __syncthreads();
        }
    }
}

// Transposed warp reduction.  Every lane enters with R partial sums (one per row of the group);
// the lanes of each LPR-lane team leave with the team totals spread over them: while more than one
// value is left, a step at offset OFF halves the value count (each lane keeps the half selected
// by its lane bit and receives the partner's share of it), afterwards plain butterfly steps.
// Cost: R-1 + log2(LPR/R) shuffles instead of R*log2(LPR).  `rbase` returns the row (0..R) of
// the lane's value 0; values p = 0..CNT-1 belong to rows rbase + p, CNT = max(1, R/LPR).
template <int CNT, int OFF, class T, int R>
__device__ __forceinline__ void reduce_step(T (&v)[R], int lane, int& rbase) {
    if constexpr (OFF >= 1) {
        if constexpr (CNT > 1) {
            constexpr int H = CNT / 2;
            const bool up = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const T send = up ? v[i] : v[i + H];
                const T keep = up ? v[i + H] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            if (up) rbase += H;
            reduce_step<H, OFF / 2>(v, lane, rbase);
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
            reduce_step<1, OFF / 2>(v, lane, rbase);
        }
    }
}
template <int R, int LPR, class T>
__device__ __forceinline__ int reduce_rows(T (&v)[R], int lane) {
    int rbase = 0;
    reduce_step<R, LPR / 2>(v, lane, rbase);
    return rbase;
}
template <int R, int LPR> struct RowsLeft { static constexpr int CNT = (R / LPR) > 1 ? (R / LPR) : 1; static constexpr int DUP = LPR > R ? LPR / R : 1; };

// offer the candidates held by the lanes of a warp (kInf = nothing) to the warp's running list
// k <= 32: the warp's running list lives in registers, one entry per lane (lanes >= k hold 0 and never
// are the maximum); an insert overwrites the lane holding the current maximum and re-reduces (max, lane).
// (max, lane) of the 64-bit entries comes from two hardware warp reductions (REDUX) on the halves and one ballot
// instead of a five-round shuffle butterfly: inserts are what a short scan is made of (config A: ~24 per warp for 42 rows).
__device__ __forceinline__ void reg_list_insert(u64& mine, int k, u64 cand, u64& thr, int& pos, int lane) {
    if (lane == pos) mine = cand;
    // while the list is filling up (thr is still a kInf filler) the fillers sit in lanes pos+1 .. k-1 -- they are consumed
    // lowest lane first -- so the next maximum is known without looking: the first k inserts of a warp cost nothing
    if (thr == kInf && pos + 1 < k) { pos++; return; }
    const bool real = lane < k;
    const uint32_t hi = real ? (uint32_t)(mine >> 32) : 0u, lo = real ? (uint32_t)mine : 0u;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned who = __ballot_sync(0xffffffffu, real && hi == mh && lo == ml);   // equal entries (kInf fillers): lowest lane
    thr = ((u64)mh << 32) | ml;
    pos = __ffs(who) - 1;
}
__device__ __forceinline__ void offer(u64 pk, u64* list, int k, u64& thr, int& pos, int lane, u64& mine, bool regmode) {
    unsigned hits = __ballot_sync(0xffffffffu, pk < thr);
    while (hits) {
        const int src = __ffs(hits) - 1;
        hits &= hits - 1;
        const u64 c = shfl_u64(pk, src);
        if (c < thr) {
            if (regmode) reg_list_insert(mine, k, c, thr, pos, lane);
            else list_insert(list, k, c, thr, pos, lane);
        }
    }
}

// WIDE rows (> 16 chunks of 16 bytes): one warp per row, R consecutive rows per warp step.
// Loads are software-pipelined in registers: each 16-byte register tile is refilled with the
// next chunk (or the next row group's first chunk) right after it is consumed, so every lane keeps
// R loads in flight all the time.  The resident buffer carries kPadRows zero rows past the end
// (engine.cu), so the last partial group needs no clamping; rows >= n are dropped before insert.
template <int DT, int OP, int NQ, int R>
__global__ void __launch_bounds__(kScanThreads, 2) scan_topk_wide(const ScanParams p, const ScanTail tail) {
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

    const uint32_t n = (uint32_t)p.n;
    const uint32_t pitch = p.pitch;
    const bool ragged = (C & 31) != 0;
    const unsigned char* lane_base = p.rows + (size_t)lane * 16;
    int myrow;
    {   // which row of a group this lane reports after the transposed reduction
        float t[R];
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = 0.f;
        myrow = reduce_rows<R, 32>(t, lane);
    }
    const bool reporter = (lane & (RowsLeft<R, 32>::DUP - 1)) == 0;

    const uint32_t groups = (uint32_t)(((uint64_t)n + R - 1) / R), gstride = gridDim.x * kScanWarps;
    uint32_t g = blockIdx.x * kScanWarps + warp;
    const unsigned char* ptr = lane_base;
    uint4 x[R];
    float xn = 0.0f;
    if (g < groups) {
        ptr = lane_base + (size_t)g * R * pitch;
#pragma unroll
        for (int r = 0; r < R; r++) x[r] = ldg_stream16(ptr + (size_t)r * pitch);
        if (p.metric == M_COS) xn = p.norm[g * R + myrow];
    }
    // (the first row tile is on its way from HBM while the queries are staged: on short scans -- BASELINE configs[0] is
    // 42 rows per warp -- the prologue is a visible part of the search)
    stage_queries<DT, NQ>(p, q_s, dpad, qc_s, red_s);

    u64 thr[NQ]; int pos[NQ]; QueryConst qc[NQ];
    u64 regl[NQ];                            // register-resident list entry of this lane (k <= 32)
    const bool regmode = p.k <= 32;
#pragma unroll
    for (int q = 0; q < NQ; q++) { thr[q] = kInf; pos[q] = 0; qc[q] = qc_s[q]; regl[q] = lane < (int)p.k ? kInf : 0ull; }

    // the row loop is compiled twice: ragged rows (chunk count not a multiple of 32) zero the lanes past
    // the row end when a tile is consumed; the common shapes skip that work entirely
    auto run = [&](auto ragged_tag) {
    constexpr bool RAGGED = decltype(ragged_tag)::value;
    float facc[NQ][R]; int iacc[NQ][R]; int ixx[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        ixx[r] = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) { facc[q][r] = 0.0f; iacc[q][r] = 0; }
    }
    uint32_t c = 0;                          // chunk of the current group the step consumes
    float xn_n = 0.0f;
    // One step = one 16-byte chunk of all R rows.  Two register tiles alternate: the step consumes `cur`
    // while `nxt` is refilled with the following chunk (or the next group's first one), so a load never
    // lands in registers that still have to be copied somewhere else.
    auto step = [&](uint4 (&cur)[R], uint4 (&nxt)[R]) {
        const bool lastc = c + 1 == CPL;
        const uint32_t gn = g + gstride;
        const bool more = gn < groups && gn > g;
        // the refill is unconditional (a predicated load would force register copies): after the very last
        // chunk it just re-reads the current one
        const unsigned char* tp = lastc ? (more ? lane_base + (size_t)gn * R * pitch : ptr) : ptr + (size_t)(c + 1) * 512;
        if (lastc && p.metric == M_COS && more) xn_n = p.norm[gn * R + myrow];
        const uint32_t ci = c * 32 + lane;
        // ragged rows: lanes past the row end read whatever follows (always inside the padded buffer) and
        // the tile is zeroed when consumed
        const bool dead = RAGGED && ci >= C;
        float qf[NQ][F ? EPC : 1]; uint4 qv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            if (F) {
                const float* qb = reinterpret_cast<const float*>(q_s) + (size_t)q * dpad;
                float4 a = reinterpret_cast<const float4*>(qb)[ci];
                qf[q][0] = a.x; qf[q][1 % EPC] = a.y; qf[q][2 % EPC] = a.z; qf[q][3 % EPC] = a.w;
                if (EPC == 8) {
                    float4 b = reinterpret_cast<const float4*>(qb + dpad / 2)[ci];
                    qf[q][4 % EPC] = b.x; qf[q][5 % EPC] = b.y; qf[q][6 % EPC] = b.z; qf[q][7 % EPC] = b.w;
                }
            } else {
                qv[q] = reinterpret_cast<const uint4*>(q_s + (size_t)q * dpad)[ci];
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            nxt[r] = ldg_stream16(tp + (size_t)r * pitch);
            uint4 cv = cur[r];
            if (RAGGED && dead) cv = make_uint4(0, 0, 0, 0);
            if (F) {
                float xf[EPC];
                widen<DT>(cv, xf);
#pragma unroll
                for (int q = 0; q < NQ; q++) accum_float<DT, OP>(xf, qf[q], facc[q][r]);
            } else {
                if (OP == OP_L2) {
                    ixx[r] = dp4<DT>(cv.x, cv.x, ixx[r]); ixx[r] = dp4<DT>(cv.y, cv.y, ixx[r]);
                    ixx[r] = dp4<DT>(cv.z, cv.z, ixx[r]); ixx[r] = dp4<DT>(cv.w, cv.w, ixx[r]);
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    int a = iacc[q][r];
                    a = dp4<DT>(qv[q].x, cv.x, a); a = dp4<DT>(qv[q].y, cv.y, a);
                    a = dp4<DT>(qv[q].z, cv.z, a); a = dp4<DT>(qv[q].w, cv.w, a);
                    iacc[q][r] = a;
                }
            }
        }
        if (!lastc) { c++; return; }
        // end of the row group -- transposed reductions: one row total per lane, then one finalize per lane
        const uint32_t row = g * R + myrow;
        bool mine = reporter && row < n;
        if (p.mask != nullptr && mine) mine = ((p.mask[row >> 5] >> (row & 31)) & 1u) == 0;   // deleted / filtered rows never compete
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            float fv = 0.0f; int iv = 0;
            if (F) { reduce_rows<R, 32>(facc[q], lane); fv = facc[q][0]; }
            else {
                if (OP == OP_L2) {   // ||x||^2 - 2 q.x reduces as one value
#pragma unroll
                    for (int r = 0; r < R; r++) iacc[q][r] = ixx[r] - 2 * iacc[q][r];
                }
                reduce_rows<R, 32>(iacc[q], lane); iv = iacc[q][0];
            }
            u64 pk = kInf;
            if (mine) {
                if (!F && OP == OP_L2) pk = ((u64)ordered_i32(qc[q].qq + iv) << 32) | row;
                else pk = finalize<DT>(p.metric, fv, iv, 0, qc[q], xn, row);
            }
            offer(pk, lists + (size_t)q * p.P + warp * p.k, (int)p.k, thr[q], pos[q], lane, regl[q], regmode);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            ixx[r] = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++) { facc[q][r] = 0.0f; iacc[q][r] = 0; }
        }
        c = 0; g = more ? gn : groups; ptr = tp; xn = xn_n;
    };
    uint4 y[R];                              // the second register tile
    while (g < groups) {
        step(x, y);
        if (g >= groups) break;
        step(y, x);
    }
    };
    if (ragged) run(std::true_type{}); else run(std::false_type{});
    if (regmode) emit_cta_lists_reg<NQ>(p, lists, regl, lane, warp);
    else emit_cta_lists<NQ>(p, lists);
    fused_merge_tail<NQ>(p, tail, smem);
}

// NARROW rows (<= 16 chunks): LPR = pow2 >= chunks lanes per row, 32/LPR rows per warp load,
// R loads in flight per lane (software-pipelined like the wide kernel).  The lane's query chunk
// never changes, so it lives in registers.
template <int DT, int OP, int NQ, int R, int LPR>
__device__ __forceinline__ void narrow_body(const ScanParams& p, u64* lists, const unsigned char* q_s, uint32_t dpad,
                                            const QueryConst* qc_s) {
    constexpr int EPC = Traits<DT>::EPC;
    constexpr bool F = Traits<DT>::is_float;
    constexpr int RPW = 32 / LPR, CNT = RowsLeft<R, LPR>::CNT, DUP = RowsLeft<R, LPR>::DUP;
    constexpr uint32_t RPS = (uint32_t)R * RPW;                 // rows per warp step (<= kPadRows)
    const uint32_t C = p.pitch >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int li = lane & (LPR - 1), ri = lane / LPR;

    u64 thr[NQ]; int pos[NQ]; QueryConst qc[NQ];
    u64 regl[NQ];                            // register-resident list entry of this lane (k <= 32)
    const bool regmode = p.k <= 32;
    float qf[NQ][F ? EPC : 1]; uint4 qv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        thr[q] = kInf; pos[q] = 0; qc[q] = qc_s[q]; regl[q] = lane < (int)p.k ? kInf : 0ull;
        if (F) {
            const float* qb = reinterpret_cast<const float*>(q_s) + (size_t)q * dpad;
            float4 a = reinterpret_cast<const float4*>(qb)[li];
            qf[q][0] = a.x; qf[q][1 % EPC] = a.y; qf[q][2 % EPC] = a.z; qf[q][3 % EPC] = a.w;
            if (EPC == 8) {
                float4 b = reinterpret_cast<const float4*>(qb + dpad / 2)[li];
                qf[q][4 % EPC] = b.x; qf[q][5 % EPC] = b.y; qf[q][6 % EPC] = b.z; qf[q][7 % EPC] = b.w;
            }
            qv[q] = make_uint4(0, 0, 0, 0);
        } else {
            qv[q] = reinterpret_cast<const uint4*>(q_s + (size_t)q * dpad)[li];
        }
    }
    const uint32_t n = (uint32_t)p.n;
    const uint32_t pitch = p.pitch;
    const bool active = (uint32_t)li < C;
    // lane's first row inside a step and its byte offset; row r of the step adds r*RPW rows
    const unsigned char* lane_base = p.rows + (size_t)ri * pitch + (size_t)li * 16;
    const uint32_t step_pitch = (uint32_t)RPW * pitch;
    int myr;
    {
        float t[R];
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = 0.f;
        myr = reduce_rows<R, LPR>(t, li);
    }
    const bool reporter = (li & (DUP - 1)) == 0;

    const uint32_t groups = (uint32_t)(((uint64_t)n + RPS - 1) / RPS), gstride = gridDim.x * kScanWarps;
    uint32_t g = blockIdx.x * kScanWarps + warp;
    uint4 x[R];
    float xn[CNT];
#pragma unroll
    for (int c = 0; c < CNT; c++) xn[c] = 0.0f;
    if (g < groups) {
        const unsigned char* ptr = lane_base + (size_t)g * RPS * pitch;
#pragma unroll
        for (int r = 0; r < R; r++) x[r] = ldg_stream16(ptr + (size_t)r * step_pitch);
        if (p.metric == M_COS) {
#pragma unroll
            for (int c = 0; c < CNT; c++) {
                uint32_t row = g * RPS + (uint32_t)(myr + c) * RPW + ri;
                xn[c] = p.norm[row < n ? row : n - 1];
            }
        }
    }
    // one step = one row group; two register tiles alternate like in the wide kernel
    auto step = [&](uint4 (&cur)[R], uint4 (&nxt)[R]) {
        const uint32_t gn = g + gstride;
        const bool more = gn < groups && gn > g;
        const unsigned char* ptr_n = lane_base + (size_t)(more ? gn : g) * RPS * pitch;   // unconditional refill
        float xn_n[CNT];
#pragma unroll
        for (int c = 0; c < CNT; c++) {
            xn_n[c] = 0.0f;
            if (p.metric == M_COS && more) {
                uint32_t row = gn * RPS + (uint32_t)(myr + c) * RPW + ri;
                xn_n[c] = p.norm[row < n ? row : n - 1];
            }
        }
        float facc[NQ][R]; int iacc[NQ][R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            nxt[r] = ldg_stream16(ptr_n + (size_t)r * step_pitch);
            uint4 cv = cur[r];
            if (!active) cv = make_uint4(0, 0, 0, 0);
            if (F) {
                float xf[EPC];
                widen<DT>(cv, xf);
#pragma unroll
                for (int q = 0; q < NQ; q++) { facc[q][r] = 0.0f; iacc[q][r] = 0; accum_float<DT, OP>(xf, qf[q], facc[q][r]); }
            } else {
                int xx = 0;
                if (OP == OP_L2) {
                    xx = dp4<DT>(cv.x, cv.x, xx); xx = dp4<DT>(cv.y, cv.y, xx);
                    xx = dp4<DT>(cv.z, cv.z, xx); xx = dp4<DT>(cv.w, cv.w, xx);
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    int a = 0;
                    a = dp4<DT>(qv[q].x, cv.x, a); a = dp4<DT>(qv[q].y, cv.y, a);
                    a = dp4<DT>(qv[q].z, cv.z, a); a = dp4<DT>(qv[q].w, cv.w, a);
                    iacc[q][r] = OP == OP_L2 ? xx - 2 * a : a; facc[q][r] = 0.0f;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            if (F) reduce_rows<R, LPR>(facc[q], li); else reduce_rows<R, LPR>(iacc[q], li);
#pragma unroll
            for (int c = 0; c < CNT; c++) {
                const uint32_t row = g * RPS + (uint32_t)(myr + c) * RPW + ri;
                u64 pk = kInf;
                bool live = reporter && row < n;
                if (p.mask != nullptr && live) live = ((p.mask[row >> 5] >> (row & 31)) & 1u) == 0;
                if (live) {
                    if (!F && OP == OP_L2) pk = ((u64)ordered_i32(qc[q].qq + iacc[q][c]) << 32) | row;
                    else pk = finalize<DT>(p.metric, facc[q][c], iacc[q][c], 0, qc[q], xn[c], row);
                }
                offer(pk, lists + (size_t)q * p.P + warp * p.k, (int)p.k, thr[q], pos[q], lane, regl[q], regmode);
            }
        }
        g = more ? gn : groups;
#pragma unroll
        for (int c = 0; c < CNT; c++) xn[c] = xn_n[c];
    };
    uint4 y[R];
    while (g < groups) {
        step(x, y);
        if (g >= groups) break;
        step(y, x);
    }
    // CTA epilogue (all threads of the CTA run the same narrow_body instantiation)
    if (regmode) emit_cta_lists_reg<NQ>(p, lists, regl, lane, warp);
    else emit_cta_lists<NQ>(p, lists);
}

template <int DT, int OP, int NQ, int R>
__global__ void __launch_bounds__(kScanThreads, 2) scan_topk_narrow(const ScanParams p, const int LPR, const ScanTail tail) {
    constexpr int EPC = Traits<DT>::EPC;
    constexpr bool F = Traits<DT>::is_float;
    extern __shared__ __align__(16) unsigned char smem[];
    u64* lists = reinterpret_cast<u64*>(smem);
    const uint32_t dpad = (uint32_t)LPR * EPC;
    unsigned char* q_s = smem + (size_t)NQ * p.P * 8;
    const size_t q_bytes = (size_t)NQ * dpad * (F ? 4 : 1);
    QueryConst* qc_s = reinterpret_cast<QueryConst*>(q_s + ((q_bytes + 15) & ~(size_t)15));
    float* red_s = reinterpret_cast<float*>(qc_s + NQ);
    for (uint32_t i = threadIdx.x; i < NQ * p.P; i += blockDim.x) lists[i] = kInf;
    stage_queries<DT, NQ>(p, q_s, dpad, qc_s, red_s);
    switch (LPR) {
        case 1: narrow_body<DT, OP, NQ, R, 1>(p, lists, q_s, dpad, qc_s); break;
        case 2: narrow_body<DT, OP, NQ, R, 2>(p, lists, q_s, dpad, qc_s); break;
        case 4: narrow_body<DT, OP, NQ, R, 4>(p, lists, q_s, dpad, qc_s); break;
        case 8: narrow_body<DT, OP, NQ, R, 8>(p, lists, q_s, dpad, qc_s); break;
        default: narrow_body<DT, OP, NQ, R, 16>(p, lists, q_s, dpad, qc_s); break;
    }
    fused_merge_tail<NQ>(p, tail, smem);
}

// kernel lookup for one data type (called from engine.cu through scan_kernel_for())
template <int DT, int OP>
inline const void* scan_kernel_op(int NQ, bool wide) {
    switch (NQ) {
        case 1: return wide ? (const void*)scan_topk_wide<DT, OP, 1, 8> : (const void*)scan_topk_narrow<DT, OP, 1, 8>;
        case 2: return wide ? (const void*)scan_topk_wide<DT, OP, 2, 8> : (const void*)scan_topk_narrow<DT, OP, 2, 8>;
        case 4: return wide ? (const void*)scan_topk_wide<DT, OP, 4, 4> : (const void*)scan_topk_narrow<DT, OP, 4, 4>;
        default: return wide ? (const void*)scan_topk_wide<DT, OP, 8, 2> : (const void*)scan_topk_narrow<DT, OP, 8, 2>;
    }
}
template <int DT>
inline const void* scan_kernel_dt(int op, int NQ, bool wide) {
    return op == OP_L2 ? scan_kernel_op<DT, OP_L2>(NQ, wide) : scan_kernel_op<DT, OP_DOT>(NQ, wide);
}

}  // namespace k
}  // namespace mvf
