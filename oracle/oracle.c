/*
 * oracle.c -- CPU restatement of the reference's brute-force scan.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file's shared object, and there only as the checker or the timed CPU baseline.  The product
 * (metrovector_b200/) never links, imports or calls it and has no CPU compute fallback.
 *
 * What it restates (all paths relative to /root/reference):
 *   - find_top_k_similar            examples/similarity_search.rs:140-176
 *   - the distance expression       examples/similarity_search.rs:152-157, examples/simple.rs:81-86
 *   - ScoredVector ordering         examples/similarity_search.rs:14-37
 *   - the 1-NN strict-< loop        examples/simple.rs:73-94  (lowest index wins ties)
 *   - Vector::as_f32 (f16 widening) src/vectors/vector.rs:71-92   (half 2.6.0 f16::to_f32)
 *   - f16::from_f32 on the write    src/builder.rs:187            (half 2.6.0, round-nearest-even)
 *
 * PARITY STATUS: "parity unpinned" for search results.  The reference cannot be compiled here
 * (no cargo/rustc/flatc) and none of its 99 unit tests computes a
 * distance or a top-k, so there is no reference-produced golden vector for this path.  What
 * pins this file instead: (1) the known answers derived by hand from the example programs'
 * datasets with the reference's exact arithmetic (tests/golden/
 * known_answers.json), (2) the IEEE definition of f16<->f32 conversion that `half` implements
 * (tests check pi/e from src/vectors/vector.rs:227-241 bit-for-bit).  Cosine, inner product,
 * Int8/UInt8 arithmetic and tie-breaking are NOT computed anywhere in the reference; the
 * definitions below are ours and are flagged "ours" where they occur.
 *
 * Arithmetic fidelity: Rust/LLVM never contracts a*b+c into FMA and never reassociates float
 * sums without explicit fast-math, so the reference's `.map(|(a,b)| (a-b).powi(2)).sum::<f32>()`
 * is exactly: s = 0; for j { t = a_j - b_j; s = s + t*t; } with each op rounded to f32.
 * Build with -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* enum values = schema/types.fbs:3-25 */
enum { DT_F32 = 0, DT_F16 = 1, DT_I8 = 2, DT_U8 = 3 };
enum { M_L2 = 0, M_IP = 1, M_COS = 2 };
enum { MODE_NEAREST = 0, MODE_AS_WRITTEN = 1 };

/* ------------------------------------------------------------------------------------------
 * f16 <-> f32.  to_f32 is exact widening (vector.rs:85-86).  from_f32 is round-to-nearest-even
 * (builder.rs:187 -> half::f16::from_f32).
 * ---------------------------------------------------------------------------------------- */
float oracle_f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

uint16_t oracle_f32_to_f16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t exp = (x >> 23) & 0xffu;
    uint32_t man = x & 0x7fffffu;
    if (exp == 255) { /* inf / nan */
        if (man == 0) return (uint16_t)(sign | 0x7c00u);
        return (uint16_t)(sign | 0x7c00u | 0x200u | (man >> 13));
    }
    int e = (int)exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);             /* overflow -> inf */
    if (e <= 0) {                                               /* subnormal or zero */
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1u))) half_man++;
        return (uint16_t)(sign | half_man);
    }
    uint32_t half_man = man >> 13;
    uint32_t rem = man & 0x1fffu;
    uint16_t out = (uint16_t)(sign | ((uint32_t)e << 10) | half_man);
    if (rem > 0x1000u || (rem == 0x1000u && (half_man & 1u))) out++; /* carries into exp ok */
    return out;
}

/* ------------------------------------------------------------------------------------------
 * The reference distance expression, similarity_search.rs:152-157 / simple.rs:81-86.
 * ---------------------------------------------------------------------------------------- */
float oracle_l2_distance_f32(const float* q, const float* x, uint32_t d) {
    float s = 0.0f;
    for (uint32_t j = 0; j < d; j++) {
        float t = q[j] - x[j];
        s = s + t * t;
    }
    return sqrtf(s);
}

/* ours (unpinned): sequential-f32 dot and squared norm, same no-FMA discipline */
static float dot_f32(const float* q, const float* x, uint32_t d) {
    float s = 0.0f;
    for (uint32_t j = 0; j < d; j++) s = s + q[j] * x[j];
    return s;
}

/* order-preserving map f32 -> u32 (ours): -0 -> +0, NaN -> last */
static uint32_t ordered_key_f32(float v) {
    if (v != v) return 0xffffffffu;
    if (v == 0.0f) v = 0.0f;
    uint32_t b; memcpy(&b, &v, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
uint32_t oracle_ordered_key_f32(float v) { return ordered_key_f32(v); }
static uint32_t ordered_key_i32(int32_t v) { return (uint32_t)v ^ 0x80000000u; }

typedef struct { uint32_t key; uint64_t idx; float fscore; int32_t iscore; } cand_t;

static int cand_less(const cand_t* a, const cand_t* b) { /* (key, index) lexicographic */
    if (a->key != b->key) return a->key < b->key;
    return a->idx < b->idx;
}

/* sorted insertion into best-first array of capacity k */
static void topk_push(cand_t* best, uint32_t* count, uint32_t k, const cand_t* c) {
    if (k == 0) return;
    if (*count == k && !cand_less(c, &best[k - 1])) return;
    uint32_t pos = (*count < k) ? (*count)++ : k - 1;
    while (pos > 0 && cand_less(c, &best[pos - 1])) { best[pos] = best[pos - 1]; pos--; }
    best[pos] = *c;
}

/* Row score per (dtype, metric).  x = raw row bytes, q = query in f32 (float spaces) or in the
 * space's own integer dtype (integer spaces).  Fills key + reported scores.
 *   L2  float : sqrtf(sum (q-x)^2)                       reference-pinned arithmetic
 *   L2  int   : exact i32 squared distance, no sqrt      ours
 *   IP        : key = -dot, reported score = dot          ours
 *   COS float : 1 - dot / (sqrtf(nq) * sqrtf(nx)), zero norm -> 1      ours
 *   COS int   : same from exact integer dot / norms converted to f32   ours
 */
typedef struct {
    int dtype, metric;
    uint32_t d;
    const float* qf;       /* float spaces */
    const void* qi;        /* integer spaces */
    float qnorm;           /* sqrtf(sum q^2) for cosine */
    float* scratch;        /* d floats, the as_f32() materialisation */
} rowctx_t;

static void widen_row(const rowctx_t* c, const void* x, float* out) {
    if (c->dtype == DT_F32) {
        memcpy(out, x, (size_t)c->d * 4); /* f32::from_le_bytes, vector.rs:74-79 (LE host) */
    } else {
        const uint16_t* h = (const uint16_t*)x;
        for (uint32_t j = 0; j < c->d; j++) out[j] = oracle_f16_to_f32(h[j]);
    }
}

static void score_row(const rowctx_t* c, const void* x, cand_t* out) {
    out->iscore = 0;
    if (c->dtype == DT_F32 || c->dtype == DT_F16) {
        float* xf = c->scratch;
        widen_row(c, x, xf);
        float v;
        if (c->metric == M_L2) {
            v = oracle_l2_distance_f32(c->qf, xf, c->d);
            out->fscore = v;
        } else if (c->metric == M_IP) {
            float dot = dot_f32(c->qf, xf, c->d);
            v = 0.0f - dot;
            out->fscore = dot;
        } else {
            float dot = dot_f32(c->qf, xf, c->d);
            float xn = sqrtf(dot_f32(xf, xf, c->d));
            float den = c->qnorm * xn;
            v = (den > 0.0f) ? 1.0f - dot / den : 1.0f;
            out->fscore = v;
        }
        out->key = ordered_key_f32(v);
        if (out->fscore == 0.0f) out->fscore = 0.0f;
    } else {
        int32_t dot = 0, nx = 0, nq = 0, l2 = 0;
        if (c->dtype == DT_U8) {
            const uint8_t* a = (const uint8_t*)c->qi; const uint8_t* b = (const uint8_t*)x;
            for (uint32_t j = 0; j < c->d; j++) {
                int32_t qa = a[j], xb = b[j], t = qa - xb;
                dot += qa * xb; nx += xb * xb; nq += qa * qa; l2 += t * t;
            }
        } else {
            const int8_t* a = (const int8_t*)c->qi; const int8_t* b = (const int8_t*)x;
            for (uint32_t j = 0; j < c->d; j++) {
                int32_t qa = a[j], xb = b[j], t = qa - xb;
                dot += qa * xb; nx += xb * xb; nq += qa * qa; l2 += t * t;
            }
        }
        if (c->metric == M_L2) {
            out->key = ordered_key_i32(l2); out->iscore = l2; out->fscore = (float)l2;
        } else if (c->metric == M_IP) {
            out->key = ordered_key_i32(-dot); out->iscore = dot; out->fscore = (float)dot;
        } else {
            float den = sqrtf((float)nq) * sqrtf((float)nx);
            float v = (den > 0.0f) ? 1.0f - (float)dot / den : 1.0f;
            out->key = ordered_key_f32(v); out->fscore = (v == 0.0f) ? 0.0f : v;
        }
    }
}

static size_t elem_size(int dtype) { return dtype == DT_F32 ? 4 : dtype == DT_F16 ? 2 : 1; }

static void prep_ctx(rowctx_t* c, int dtype, int metric, uint32_t d, const void* query, int qdtype,
                     float* qbuf, float* scratch) {
    c->dtype = dtype; c->metric = metric; c->d = d; c->scratch = scratch; c->qf = NULL; c->qi = NULL;
    c->qnorm = 0.0f;
    if (dtype == DT_F32 || dtype == DT_F16) {
        if (qdtype == DT_F16) {
            const uint16_t* h = (const uint16_t*)query;
            for (uint32_t j = 0; j < d; j++) qbuf[j] = oracle_f16_to_f32(h[j]);
        } else {
            memcpy(qbuf, query, (size_t)d * 4);
        }
        c->qf = qbuf;
        c->qnorm = sqrtf(dot_f32(qbuf, qbuf, d));
    } else {
        c->qi = query;
    }
}

/*
 * oracle_topk -- find_top_k_similar (similarity_search.rs:140-176) for one query.
 *   mode 0: the documented intent, k NEAREST by (key, index), best first.
 *   mode 1: the code AS WRITTEN: the reversed Ord (:23-31) makes BinaryHeap::pop (:165-168)
 *           discard the smallest distance, so the k FARTHEST survive; then sorted ascending
 *           (:172-173).  Kept only to document the defect (defect 1).
 * row_base is added to every reported index (shard emulation).  Returns the number of results
 * (min(k, n)), or -1 on bad arguments.
 */
int oracle_topk(const void* db, int dtype, uint64_t n, uint32_t d, int metric, const void* query,
                int qdtype, uint32_t k, int mode, uint64_t row_base, uint64_t* out_ids,
                float* out_scores, int32_t* out_iscores) {
    if (dtype < 0 || dtype > 3 || metric < 0 || metric > 2) return -1;
    size_t rb = (size_t)d * elem_size(dtype);
    cand_t* best = (cand_t*)malloc(sizeof(cand_t) * (k ? k : 1));
    float* qbuf = (float*)malloc(sizeof(float) * (d ? d : 1));
    float* scratch = (float*)malloc(sizeof(float) * (d ? d : 1));
    rowctx_t c; prep_ctx(&c, dtype, metric, d, query, qdtype, qbuf, scratch);
    uint32_t count = 0;
    for (uint64_t i = 0; i < n; i++) {
        cand_t cd; score_row(&c, (const uint8_t*)db + i * rb, &cd);
        cd.idx = row_base + i;
        if (mode == MODE_AS_WRITTEN) cd.key = ~cd.key; /* keep the largest */
        topk_push(best, &count, k, &cd);
    }
    for (uint32_t r = 0; r < count; r++) {
        uint32_t src = (mode == MODE_AS_WRITTEN) ? count - 1 - r : r; /* final ascending sort */
        if (out_ids) out_ids[r] = best[src].idx;
        if (out_scores) out_scores[r] = best[src].fscore;
        if (out_iscores) out_iscores[r] = best[src].iscore;
    }
    free(best); free(qbuf); free(scratch);
    return (int)count;
}

/* score of a single row, for spot checks at sizes where a full scan is too slow */
int oracle_score_row(const void* row, int dtype, uint32_t d, int metric, const void* query,
                     int qdtype, float* out_score, int32_t* out_iscore, uint32_t* out_key) {
    float* qbuf = (float*)malloc(sizeof(float) * (d ? d : 1));
    float* scratch = (float*)malloc(sizeof(float) * (d ? d : 1));
    rowctx_t c; prep_ctx(&c, dtype, metric, d, query, qdtype, qbuf, scratch);
    cand_t cd; score_row(&c, row, &cd);
    if (out_score) *out_score = cd.fscore;
    if (out_iscore) *out_iscore = cd.iscore;
    if (out_key) *out_key = cd.key;
    free(qbuf); free(scratch);
    return 0;
}

/* sqrtf(sum x^2) per row (float dtypes: sequential f32; ints: exact i32 then f32) -- ours */
void oracle_row_norms(const void* db, int dtype, uint64_t n, uint32_t d, float* out_norm,
                      float* out_sq_f32, int32_t* out_sq_i32) {
    size_t rb = (size_t)d * elem_size(dtype);
    float* xf = (float*)malloc(sizeof(float) * (d ? d : 1));
    rowctx_t c; memset(&c, 0, sizeof c); c.dtype = dtype; c.d = d;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t* x = (const uint8_t*)db + i * rb;
        if (dtype == DT_F32 || dtype == DT_F16) {
            widen_row(&c, x, xf);
            float s = dot_f32(xf, xf, d);
            if (out_sq_f32) out_sq_f32[i] = s;
            if (out_norm) out_norm[i] = sqrtf(s);
        } else {
            int32_t s = 0;
            for (uint32_t j = 0; j < d; j++) {
                int32_t v = dtype == DT_U8 ? (int32_t)x[j] : (int32_t)((const int8_t*)x)[j];
                s += v * v;
            }
            if (out_sq_i32) out_sq_i32[i] = s;
            if (out_sq_f32) out_sq_f32[i] = (float)s;
            if (out_norm) out_norm[i] = sqrtf((float)s);
        }
    }
    free(xf);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic data (ours): counter-based, any row reproducible.
 *   rk      = sm64(sm64(seed) ^ row)
 *   f32/f16 : h = sm64(rk + (col>>1)); w = 32 bits of h selected by col&1;
 *             v = (float)((w & 0xffff) + (w >> 16) - 65535) * 2^-15   (triangular, |v| < 2)
 *             f16 = round-nearest-even(v)
 *   u8/i8   : h = sm64(rk + (col>>3)); byte (col&7) of h
 * The product carries its own host+device implementation of the same definition
 * (metrovector_b200/csrc/synth.h); tests require the two to agree bit-for-bit.
 * ---------------------------------------------------------------------------------------- */
static uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_synth_fill(void* out, int dtype, uint64_t seed, uint64_t row0, uint64_t nrows,
                       uint32_t d) {
    uint64_t sk = sm64(seed);
    size_t rb = (size_t)d * elem_size(dtype);
    for (uint64_t r = 0; r < nrows; r++) {
        uint64_t rk = sm64(sk ^ (row0 + r));
        uint8_t* row = (uint8_t*)out + r * rb;
        if (dtype == DT_F32 || dtype == DT_F16) {
            for (uint32_t c = 0; c < d; c++) {
                uint64_t h = sm64(rk + (c >> 1));
                uint32_t w = (uint32_t)(h >> (32 * (c & 1)));
                int32_t s = (int32_t)((w & 0xffffu) + (w >> 16)) - 65535;
                float v = (float)s * 0x1p-15f;
                if (dtype == DT_F32) ((float*)row)[c] = v;
                else ((uint16_t*)row)[c] = oracle_f32_to_f16(v);
            }
        } else {
            for (uint32_t c = 0; c < d; c++) {
                uint64_t h = sm64(rk + (c >> 3));
                row[c] = (uint8_t)(h >> (8 * (c & 7)));
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Timed CPU baseline (bench.py cpu_baseline / --impl reference).
 *   faithful=1 : what the reference does per row -- heap-allocate a Vec<f32> (vector.rs:73,83),
 *                convert, distance, push into a binary heap and pop when len > k
 *                (similarity_search.rs:147-169); one thread (the reference is single-threaded).
 *   faithful=0 : same arithmetic, no per-row allocation, rows sharded over nthreads, merged.
 * Returns elapsed seconds for nq queries; results of the last query are written out.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const void* db; int dtype; uint64_t row0, nrows; uint32_t d; int metric; const void* query;
    int qdtype; uint32_t k; int faithful; cand_t* best; uint32_t count;
} job_t;

static void heap_sift_up(cand_t* h, uint32_t i) { /* max-heap on (key,idx): worst on top */
    while (i > 0) {
        uint32_t p = (i - 1) / 2;
        if (cand_less(&h[p], &h[i])) { cand_t t = h[p]; h[p] = h[i]; h[i] = t; i = p; } else break;
    }
}
static void heap_sift_down(cand_t* h, uint32_t n) {
    uint32_t i = 0;
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(&h[m], &h[l])) m = l;
        if (r < n && cand_less(&h[m], &h[r])) m = r;
        if (m == i) break;
        cand_t t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
}

static void* scan_job(void* arg) {
    job_t* j = (job_t*)arg;
    size_t rb = (size_t)j->d * elem_size(j->dtype);
    float* qbuf = (float*)malloc(sizeof(float) * (j->d ? j->d : 1));
    float* scratch = (float*)malloc(sizeof(float) * (j->d ? j->d : 1));
    rowctx_t c; prep_ctx(&c, j->dtype, j->metric, j->d, j->query, j->qdtype, qbuf, scratch);
    j->count = 0;
    if (j->faithful) {
        cand_t* heap = j->best; /* capacity k+1 */
        uint32_t hn = 0;
        for (uint64_t i = 0; i < j->nrows; i++) {
            float* vec = (float*)malloc(sizeof(float) * (j->d ? j->d : 1)); /* Vec<f32> per row */
            c.scratch = vec;
            cand_t cd; score_row(&c, (const uint8_t*)j->db + (j->row0 + i) * rb, &cd);
            cd.idx = j->row0 + i;
            heap[hn] = cd; heap_sift_up(heap, hn); hn++;
            if (hn > j->k) { heap[0] = heap[hn - 1]; hn--; heap_sift_down(heap, hn); }
            free(vec);
        }
        /* final sort (:172-173) */
        for (uint32_t a = 1; a < hn; a++) {
            cand_t t = heap[a]; uint32_t b = a;
            while (b > 0 && cand_less(&t, &heap[b - 1])) { heap[b] = heap[b - 1]; b--; }
            heap[b] = t;
        }
        j->count = hn;
    } else {
        for (uint64_t i = 0; i < j->nrows; i++) {
            cand_t cd; score_row(&c, (const uint8_t*)j->db + (j->row0 + i) * rb, &cd);
            cd.idx = j->row0 + i;
            topk_push(j->best, &j->count, j->k, &cd);
        }
    }
    free(qbuf); free(scratch);
    return NULL;
}

double oracle_scan_timed(const void* db, int dtype, uint64_t n, uint32_t d, int metric,
                         const void* queries, int qdtype, uint32_t nq, uint32_t k, int nthreads,
                         int faithful, uint64_t* out_ids, float* out_scores) {
    if (nthreads < 1) nthreads = 1;
    if (faithful) nthreads = 1;
    size_t qb = (size_t)d * ((dtype == DT_F32 || dtype == DT_F16) ? elem_size(qdtype) : 1);
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) jobs[t].best = (cand_t*)malloc(sizeof(cand_t) * (k + 1));
    cand_t* merged = (cand_t*)malloc(sizeof(cand_t) * (k ? k : 1));
    uint32_t mcount = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t qi = 0; qi < nq; qi++) {
        uint64_t per = (n + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
        for (int t = 0; t < nthreads; t++) {
            uint64_t b = per * (uint64_t)t; if (b > n) b = n;
            uint64_t e = b + per; if (e > n) e = n;
            job_t* j = &jobs[t];
            j->db = db; j->dtype = dtype; j->row0 = b; j->nrows = e - b; j->d = d; j->metric = metric;
            j->query = (const uint8_t*)queries + qi * qb; j->qdtype = qdtype; j->k = k;
            j->faithful = faithful;
            if (nthreads == 1) scan_job(j); else pthread_create(&th[t], NULL, scan_job, j);
        }
        mcount = 0;
        for (int t = 0; t < nthreads; t++) {
            if (nthreads > 1) pthread_join(th[t], NULL);
            for (uint32_t r = 0; r < jobs[t].count; r++) topk_push(merged, &mcount, k, &jobs[t].best[r]);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (uint32_t r = 0; r < mcount; r++) {
        if (out_ids) out_ids[r] = merged[r].idx;
        if (out_scores) out_scores[r] = merged[r].fscore;
    }
    for (int t = 0; t < nthreads; t++) free(jobs[t].best);
    free(jobs); free(th); free(merged);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* multi-threaded synthetic fill so bench.py's CPU sample is generated quickly */
typedef struct { void* out; int dtype; uint64_t seed, row0, nrows; uint32_t d; } fill_job_t;
static void* fill_job(void* a) {
    fill_job_t* f = (fill_job_t*)a;
    oracle_synth_fill(f->out, f->dtype, f->seed, f->row0, f->nrows, f->d);
    return NULL;
}
void oracle_synth_fill_mt(void* out, int dtype, uint64_t seed, uint64_t row0, uint64_t nrows,
                          uint32_t d, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    size_t rb = (size_t)d * elem_size(dtype);
    fill_job_t* jobs = (fill_job_t*)calloc((size_t)nthreads, sizeof(fill_job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    uint64_t per = (nrows + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t b = per * (uint64_t)t; if (b > nrows) b = nrows;
        uint64_t e = b + per; if (e > nrows) e = nrows;
        jobs[t] = (fill_job_t){(uint8_t*)out + b * rb, dtype, seed, row0 + b, e - b, d};
        pthread_create(&th[t], NULL, fill_job, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
}
