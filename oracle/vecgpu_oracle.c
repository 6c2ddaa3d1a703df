/*
 * vecgpu_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement of the reference's distance-scoring path
 * (brianmacy/sqlite-vec-hnsw; all file:line citations are relative to the
 * reference tree).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's library; the
 * product (libvecgpu.so) never links, loads or calls it.
 *
 * PARITY STATUS
 *   - i8 L2, i8 L1, bit Hamming, f32 L1: PINNED BY ARITHMETIC.  They are exact
 *     integer sums (+ one correctly rounded IEEE sqrt / cast), or — f32 L1 — a
 *     strict left-to-right f32 sum written in the reference's own Rust
 *     (src/distance/scalar.rs:31-35, no reassociation), so this restatement is
 *     bit-identical to the reference by construction.
 *   - f32 L2, f32 cosine, i8 cosine: PARITY UNPINNED at working precision.
 *     The reference calls simsimd 6.5.16 (Cargo.toml:22, Cargo.lock:725-731)
 *     for these (src/distance/scalar.rs:17,48,94); that crate is neither
 *     vendored under the reference tree nor installed here, and there is no
 *     Rust toolchain, so its lane order and rsqrt approximation cannot be
 *     observed.  The reference's own tests pin these only to +-0.01
 *     (src/distance/scalar.rs:120-171, src/distance/mod.rs:165-188); all of
 *     those known answers are checked in tests/test_oracle_golden.py.  The
 *     accumulation order used here ("canonical order") is the
 *     hypothesised AVX-512 shape: 16 f32 FMA lanes, element i -> lane i%16,
 *     balanced tree reduction pairing lane i with i+8, +4, +2, +1.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp).
 */
#define _GNU_SOURCE
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_F32 0
#define ORC_I8 1
#define ORC_BIT 2
#define ORC_L2 0
#define ORC_L1 1
#define ORC_COSINE 2
#define ORC_HAMMING 3

#define ORC_OK 0
#define ORC_ERR_INVALID_PARAM 1
#define ORC_ERR_DIM_MISMATCH 2
#define ORC_ERR_UNSUPPORTED 3

/* Runtime ISA dispatch without changing results: every clone executes the
 * same IEEE operations in the same order. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORC_NO_CLONES)
#define ORC_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define ORC_CLONES
#endif

/* ------------------------------------------------------------------------ */
/* formats — src/vector.rs:223-242 (f32 LE / i8), :592-600 (bits LSB-first)    */

uint32_t orc_row_bytes(int elem, uint32_t dims) {
    switch (elem) {
        case ORC_F32: return dims * 4u;
        case ORC_I8: return dims;
        case ORC_BIT: return (dims + 7u) / 8u;
        default: return 0;
    }
}

/* the seven pairs of src/distance/mod.rs:70-83 */
int orc_metric_supported(int elem, int metric) {
    if (elem == ORC_F32 || elem == ORC_I8) return metric == ORC_L2 || metric == ORC_L1 || metric == ORC_COSINE;
    if (elem == ORC_BIT) return metric == ORC_HAMMING;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* canonical 16-lane f32 accumulation                          */

static inline float tree16(float* l) {
    for (int i = 0; i < 8; ++i) l[i] = l[i] + l[i + 8];
    for (int i = 0; i < 4; ++i) l[i] = l[i] + l[i + 4];
    for (int i = 0; i < 2; ++i) l[i] = l[i] + l[i + 2];
    return l[0] + l[1];
}

/* Hand-vectorised forms of the two canonical kernels for hosts with AVX-512 (the CPU arm of bench.py runs at
 * memory speed with them).  Lane j of the zmm accumulator IS canonical lane j: the same IEEE operations in the same
 * order, so the results are bit-identical to the plain C loops below (tests/test_oracle_golden.py checks this). */
#if defined(__x86_64__) && defined(__GNUC__)
#define ORC_HAVE_AVX512 1
__attribute__((target("avx512f"))) static float l2sq_f32_avx512(const float* a, const float* b, uint32_t d) {
    __m512 acc = _mm512_setzero_ps();
    uint32_t i = 0;
    for (; i + 16 <= d; i += 16) {
        __m512 t = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        acc = _mm512_fmadd_ps(t, t, acc);
    }
    if (i < d) {
        const __mmask16 m = (__mmask16)((1u << (d - i)) - 1u);
        __m512 t = _mm512_sub_ps(_mm512_maskz_loadu_ps(m, a + i), _mm512_maskz_loadu_ps(m, b + i));
        acc = _mm512_mask3_fmadd_ps(t, t, acc, m); /* lanes past the tail stay untouched */
    }
    float l[16];
    _mm512_storeu_ps(l, acc);
    return tree16(l);
}
__attribute__((target("avx512f"))) static void dot3_f32_avx512(const float* a, const float* b, uint32_t d, float* ab,
                                                               float* a2, float* b2) {
    __m512 vab = _mm512_setzero_ps(), va = _mm512_setzero_ps(), vb = _mm512_setzero_ps();
    uint32_t i = 0;
    for (; i + 16 <= d; i += 16) {
        __m512 x = _mm512_loadu_ps(a + i), y = _mm512_loadu_ps(b + i);
        vab = _mm512_fmadd_ps(x, y, vab);
        va = _mm512_fmadd_ps(x, x, va);
        vb = _mm512_fmadd_ps(y, y, vb);
    }
    if (i < d) {
        const __mmask16 m = (__mmask16)((1u << (d - i)) - 1u);
        __m512 x = _mm512_maskz_loadu_ps(m, a + i), y = _mm512_maskz_loadu_ps(m, b + i);
        vab = _mm512_mask3_fmadd_ps(x, y, vab, m);
        va = _mm512_mask3_fmadd_ps(x, x, va, m);
        vb = _mm512_mask3_fmadd_ps(y, y, vb, m);
    }
    float l[16];
    _mm512_storeu_ps(l, vab);
    *ab = tree16(l);
    _mm512_storeu_ps(l, va);
    *a2 = tree16(l);
    _mm512_storeu_ps(l, vb);
    *b2 = tree16(l);
}
static int orc_use_avx512 = -1; /* -1: not probed; orc_force_plain(1) pins the plain C loops (tests) */
static inline int use_avx512(void) {
    if (orc_use_avx512 < 0) orc_use_avx512 = __builtin_cpu_supports("avx512f") ? 1 : 0;
    return orc_use_avx512;
}
void orc_force_plain(int plain) { orc_use_avx512 = plain ? 0 : -1; }
int orc_has_avx512(void) { return __builtin_cpu_supports("avx512f") ? 1 : 0; }
#else
#define ORC_HAVE_AVX512 0
void orc_force_plain(int plain) { (void)plain; }
int orc_has_avx512(void) { return 0; }
#endif

/* sum (a_i - b_i)^2 — stands in for simsimd f32::sqeuclidean, src/distance/scalar.rs:17 */
ORC_CLONES
static float l2sq_f32_plain(const float* a, const float* b, uint32_t d) {
    float l[16];
    for (int j = 0; j < 16; ++j) l[j] = 0.0f;
    uint32_t i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; ++j) {
            float t = a[i + j] - b[i + j];
            l[j] = fmaf(t, t, l[j]);
        }
    for (int j = 0; i + j < d; ++j) {
        float t = a[i + j] - b[i + j];
        l[j] = fmaf(t, t, l[j]);
    }
    return tree16(l);
}

static inline float l2sq_f32(const float* a, const float* b, uint32_t d) {
#if ORC_HAVE_AVX512
    if (use_avx512()) return l2sq_f32_avx512(a, b, d);
#endif
    return l2sq_f32_plain(a, b, d);
}

/* ab, a2, b2 — stands in for the accumulation half of simsimd f32::cosine, scalar.rs:48 */
ORC_CLONES
static void dot3_f32_plain(const float* a, const float* b, uint32_t d, float* ab, float* a2, float* b2) {
    float lab[16], la[16], lb[16];
    for (int j = 0; j < 16; ++j) lab[j] = la[j] = lb[j] = 0.0f;
    uint32_t i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; ++j) {
            float x = a[i + j], y = b[i + j];
            lab[j] = fmaf(x, y, lab[j]);
            la[j] = fmaf(x, x, la[j]);
            lb[j] = fmaf(y, y, lb[j]);
        }
    for (int j = 0; i + j < d; ++j) {
        float x = a[i + j], y = b[i + j];
        lab[j] = fmaf(x, y, lab[j]);
        la[j] = fmaf(x, x, la[j]);
        lb[j] = fmaf(y, y, lb[j]);
    }
    *ab = tree16(lab);
    *a2 = tree16(la);
    *b2 = tree16(lb);
}

static inline void dot3_f32(const float* a, const float* b, uint32_t d, float* ab, float* a2, float* b2) {
#if ORC_HAVE_AVX512
    if (use_avx512()) {
        dot3_f32_avx512(a, b, d, ab, a2, b2);
        return;
    }
#endif
    dot3_f32_plain(a, b, d, ab, a2, b2);
}

/* the normalisation half of cosine, in f64 with IEEE sqrt/div:
 * a2==0 && b2==0 -> 0 ; ab==0 -> 1 ; else 1 - ab/(sqrt(a2)*sqrt(b2)), clamped at 0
 * (simsimd uses an rsqrt estimate here — unpinned, see header). */
static inline float cos_finish(double ab, double a2, double b2) {
    if (a2 == 0.0 && b2 == 0.0) return 0.0f;
    if (ab == 0.0) return 1.0f;
    double r = 1.0 - ab / (sqrt(a2) * sqrt(b2));
    return (float)(r > 0.0 ? r : 0.0);
}

/* ------------------------------------------------------------------------ */
/* the seven distance functions of src/distance/scalar.rs                      */

/* scalar.rs:12-21 — cast the f64 sum to f32, THEN f32 sqrt */
static float dist_l2_f32(const float* a, const float* b, uint32_t d) { return sqrtf(l2sq_f32(a, b, d)); }

/* scalar.rs:25-38 — strict left-to-right f32 sum of |a-b| */
static float dist_l1_f32(const float* a, const float* b, uint32_t d) {
    float s = 0.0f;
    for (uint32_t i = 0; i < d; ++i) s = s + fabsf(a[i] - b[i]);
    return s;
}

/* scalar.rs:42-52 */
static float dist_cos_f32(const float* a, const float* b, uint32_t d) {
    float ab, a2, b2;
    dot3_f32(a, b, d, &ab, &a2, &b2);
    return cos_finish((double)ab, (double)a2, (double)b2);
}

/* scalar.rs:56-66 — exact integer sum, f64 sqrt, THEN cast (opposite order to f32) */
static float dist_l2_i8(const int8_t* a, const int8_t* b, uint32_t d) {
    int64_t s = 0;
    for (uint32_t i0 = 0; i0 < d; i0 += 16384) { /* exact: 16384 * 255^2 < 2^31, blocks summed in i64 */
        const uint32_t i1 = d - i0 < 16384 ? d : i0 + 16384;
        int32_t sb = 0;
        for (uint32_t i = i0; i < i1; ++i) {
            int32_t t = (int32_t)a[i] - (int32_t)b[i];
            sb += t * t;
        }
        s += sb;
    }
    return (float)sqrt((double)s);
}

/* scalar.rs:70-83 — i32 sum of |a-b|, cast to f32 */
static float dist_l1_i8(const int8_t* a, const int8_t* b, uint32_t d) {
    int32_t s = 0;
    for (uint32_t i = 0; i < d; ++i) {
        int32_t t = (int32_t)a[i] - (int32_t)b[i];
        s += t < 0 ? -t : t;
    }
    return (float)s;
}

/* scalar.rs:88-98 — exact integer ab/a2/b2, FP finish */
static float dist_cos_i8(const int8_t* a, const int8_t* b, uint32_t d) {
    int64_t ab = 0, a2 = 0, b2 = 0;
    for (uint32_t i0 = 0; i0 < d; i0 += 16384) { /* exact: 16384 * 128^2 < 2^31, blocks summed in i64 */
        const uint32_t i1 = d - i0 < 16384 ? d : i0 + 16384;
        int32_t sab = 0, sa = 0, sb = 0;
        for (uint32_t i = i0; i < i1; ++i) {
            int32_t x = a[i], y = b[i];
            sab += x * y;
            sa += x * x;
            sb += y * y;
        }
        ab += sab, a2 += sa, b2 += sb;
    }
    return cos_finish((double)ab, (double)a2, (double)b2);
}

/* scalar.rs:102-112 — popcount(a xor b) over as_bytes(), padding bits included */
static float dist_hamming(const uint8_t* a, const uint8_t* b, uint32_t nbytes) {
    uint32_t s = 0;
    uint32_t i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        s += (uint32_t)__builtin_popcountll(x ^ y);
    }
    for (; i < nbytes; ++i) s += (uint32_t)__builtin_popcount((unsigned)(a[i] ^ b[i]));
    return (float)s;
}

/* dispatch — src/distance/mod.rs:52-84 (dims check, then the 7-way match) */
static inline float dist_dispatch(int elem, int metric, const void* a, const void* b, uint32_t dims) {
    if (elem == ORC_F32) {
        if (metric == ORC_L2) return dist_l2_f32(a, b, dims);
        if (metric == ORC_L1) return dist_l1_f32(a, b, dims);
        return dist_cos_f32(a, b, dims);
    }
    if (elem == ORC_I8) {
        if (metric == ORC_L2) return dist_l2_i8(a, b, dims);
        if (metric == ORC_L1) return dist_l1_i8(a, b, dims);
        return dist_cos_i8(a, b, dims);
    }
    return dist_hamming(a, b, (dims + 7u) / 8u);
}

int orc_distance(int elem, uint32_t dims_a, uint32_t dims_b, const void* a, const void* b, int metric, float* out) {
    if (elem < 0 || elem > 2 || metric < 0 || metric > 3) return ORC_ERR_INVALID_PARAM;
    if (dims_a != dims_b) return ORC_ERR_DIM_MISMATCH;       /* mod.rs:57-62 */
    if (!orc_metric_supported(elem, metric)) return ORC_ERR_UNSUPPORTED; /* mod.rs:78-82 */
    *out = dist_dispatch(elem, metric, a, b, dims_a);
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* total order used for ranking: (d_f32, position) ascending; NaN after +inf */
/* (the reference's comparator maps NaN to Equal, vtab.rs:2619 —  */
/* a non-total order; the deviation is documented in DESIGN.md).             */

static inline uint32_t order_bits(float d) {
    uint32_t u;
    memcpy(&u, &d, 4);
    if (d != d) return 0xFFFFFFFFu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct {
    uint32_t key; /* order_bits(distance) */
    uint32_t pos; /* row position == ascending rowid order */
} orc_rank_t;

/* Stable LSD radix sort on `key` (3 passes of 11 bits).  The input is in
 * ascending `pos` order, so stability gives exactly the reference's
 * stable sort_by(distance) over ascending rowids (src/vtab.rs:2619). */
static void rank_sort(orc_rank_t* a, orc_rank_t* tmp, uint64_t n) {
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass * 11;
        uint64_t cnt[2049];
        memset(cnt, 0, sizeof(cnt));
        for (uint64_t i = 0; i < n; ++i) cnt[((a[i].key >> shift) & 2047u) + 1]++;
        for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
        for (uint64_t i = 0; i < n; ++i) tmp[cnt[(a[i].key >> shift) & 2047u]++] = a[i];
        orc_rank_t* t = a;
        a = tmp;
        tmp = t;
    }
    /* 3 passes: the sorted data ended in the buffer that was `tmp` at entry */
}

/*
 * Exact scan — src/vtab.rs:2573-2623.
 *   rowids: ascending (src/shadow.rs:853-868) or NULL for 1..n
 *   skip:   optional n bytes, non-zero = row's blob is empty / absent / of the
 *           wrong length and is skipped (vtab.rs:2596-2613)
 * Distances for ALL rows are computed, then all (rowid, dist) pairs are sorted
 * (vtab.rs:2619) and truncated (vtab.rs:2620).  OpenMP over rows only.
 * Unused result slots: rowid -1, distance +inf.
 */
int orc_knn(int elem, uint32_t dims, const int64_t* rowids, const void* vectors, const uint8_t* skip, uint64_t n,
            const void* queries, uint32_t nq, uint32_t k, int metric, int64_t* out_rowids, float* out_dists,
            uint32_t* out_counts) {
    if (elem < 0 || elem > 2 || metric < 0 || metric > 3 || dims == 0) return ORC_ERR_INVALID_PARAM;
    if (!orc_metric_supported(elem, metric)) return ORC_ERR_UNSUPPORTED;
    const uint32_t rb = orc_row_bytes(elem, dims);
    if (n >= 0xFFFFFFFFull) return ORC_ERR_INVALID_PARAM;
    orc_rank_t* r = malloc((n ? n : 1) * sizeof(orc_rank_t));
    orc_rank_t* r2 = malloc((n ? n : 1) * sizeof(orc_rank_t));
    if (!r || !r2) return ORC_ERR_INVALID_PARAM;
    for (uint32_t q = 0; q < nq; ++q) {
        const char* qv = (const char*)queries + (size_t)q * rb;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            float d = dist_dispatch(elem, metric, qv, (const char*)vectors + (size_t)i * rb, dims);
            r[i].key = order_bits(d);
            r[i].pos = (uint32_t)i;
        }
        uint64_t m = 0;
        for (uint64_t i = 0; i < n; ++i)
            if (!skip || !skip[i]) r[m++] = r[i];
        rank_sort(r, r2, m);
        orc_rank_t* sorted = r2; /* odd number of passes */
        uint32_t cnt = (uint32_t)(m < k ? m : k);
        for (uint32_t j = 0; j < k; ++j) {
            if (j < cnt) {
                uint64_t pos = sorted[j].pos;
                out_rowids[(size_t)q * k + j] = rowids ? rowids[pos] : (int64_t)pos + 1;
                out_dists[(size_t)q * k + j] =
                    dist_dispatch(elem, metric, qv, (const char*)vectors + (size_t)pos * rb, dims);
            } else {
                out_rowids[(size_t)q * k + j] = -1;
                out_dists[(size_t)q * k + j] = INFINITY;
            }
        }
        if (out_counts) out_counts[q] = cnt;
    }
    free(r);
    free(r2);
    return ORC_OK;
}

/* All distances of one query against n rows (no ranking): the neighbour loop
 * of search_layer, src/hnsw/search.rs:501-513. */
int orc_distances(int elem, uint32_t dims, const void* vectors, uint64_t n, const void* query, int metric,
                  float* out) {
    if (!orc_metric_supported(elem, metric)) return ORC_ERR_UNSUPPORTED;
    const uint32_t rb = orc_row_bytes(elem, dims);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        out[i] = dist_dispatch(elem, metric, query, (const char*)vectors + (size_t)i * rb, dims);
    return ORC_OK;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* producers — src/vector.rs:444-608                                           */

/* vector.rs:444-466: magnitude = sqrt of a strict left-to-right f32 sum */
int orc_normalize_f32(const float* in, uint32_t d, float* out) {
    float s = 0.0f;
    for (uint32_t i = 0; i < d; ++i) s = s + in[i] * in[i];
    float m = sqrtf(s);
    if (m == 0.0f) return ORC_ERR_INVALID_PARAM;
    for (uint32_t i = 0; i < d; ++i) out[i] = in[i] / m;
    return ORC_OK;
}

/* Rust f32::round = half away from zero */
static inline float round_haz(float x) {
    /* == roundf(x) for |x| < 2^23 (all this file rounds): t = trunc(x); x - t is exact; step away from zero when the
     * fraction is >= 0.5.  Written without libm and without branches so the quantiser loops vectorise;
     * tests/test_oracle_golden.py compares it with roundf. */
    float t = (float)(int32_t)x;
    float f = x - t;
    t += (f >= 0.5f) ? 1.0f : 0.0f;
    t -= (f <= -0.5f) ? 1.0f : 0.0f;
    return t;
}
float orc_round_haz(float x) { return fabsf(x) < 8388608.0f ? round_haz(x) : x; }

/* vector.rs:514-545 */
void orc_quantize_int8(const float* restrict in, uint32_t d, int8_t* restrict out) {
    float mn = INFINITY, mx = -INFINITY;
    for (uint32_t i = 0; i < d; ++i) { /* f32::min / f32::max folds (NaN operands are ignored, as in Rust) */
        mn = in[i] < mn ? in[i] : mn;
        mx = in[i] > mx ? in[i] : mx;
    }
    if (mn == mx) {
        memset(out, 0, d);
        return;
    }
    float range = mx - mn;
    for (uint32_t i = 0; i < d; ++i) {
        float normalized = (in[i] - mn) / range;
        float scaled = normalized * 255.0f - 128.0f; /* in [-128, 127] up to rounding: |scaled| < 2^23 */
        float r = round_haz(scaled);
        r = r < -128.0f ? -128.0f : r;
        r = r > 127.0f ? 127.0f : r;
        out[i] = (int8_t)(int32_t)r;
    }
}

/* vector.rs:554-575 */
void orc_quantize_int8_for_index(const float* in, uint32_t d, int8_t* out) {
    for (uint32_t i = 0; i < d; ++i) {
        float c = in[i] < -1.0f ? -1.0f : (in[i] > 1.0f ? 1.0f : in[i]);
        out[i] = (int8_t)round_haz(c * 127.0f);
    }
}

/* vector.rs:579-608 */
void orc_quantize_binary(const float* in, uint32_t d, uint8_t* out) {
    float s = 0.0f;
    for (uint32_t i = 0; i < d; ++i) s = s + in[i];
    float mean = s / (float)d;
    memset(out, 0, (d + 7u) / 8u);
    for (uint32_t i = 0; i < d; ++i)
        if (in[i] >= mean) out[i / 8] |= (uint8_t)(1u << (i % 8));
}

/* src/hnsw/mod.rs:139-146 — cosine output conversion for normalised HNSW columns */
float orc_convert_cosine_output(float d_l2) { return (d_l2 * d_l2) / 2.0f; }

/* ------------------------------------------------------------------------ */
/* synthetic corpus generator value(seed, rowid, word) — SURVEY §8d.          */
/* Pure integer hashing + exact int->float conversions, so the CUDA copy in  */
/* csrc/synth.cuh produces identical bytes.                                  */

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t synth_word(uint64_t seed, int64_t rowid, uint32_t w) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)rowid);
    return mix64(h + 0xD1B54A32D192ED03ull * (uint64_t)(w + 1));
}

static inline float u24_to_unit(uint32_t u) { /* U[-1,1) on a 2^-23 grid */
    return (float)(u >> 8) * 0x1p-23f - 1.0f;
}

static inline float gauss4(uint64_t h) { /* Irwin-Hall(4): sum of four u16, centred, * 2^-15 */
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                (int32_t)((h >> 48) & 0xFFFF) - 131070;
    return (float)s * 0x1p-15f;
}

void orc_synth_f32_row(uint64_t seed, int64_t rowid, uint32_t dims, int kind, float* out) {
    /* synth_word(seed, rowid, w) with the row hash hoisted; plain loops so the compiler can vectorise the 64-bit mixes */
    const uint64_t hr = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)rowid);
    if (kind == 1) {
        for (uint32_t j = 0; j < dims; ++j) out[j] = gauss4(mix64(hr + 0xD1B54A32D192ED03ull * (uint64_t)(j + 1)));
    } else {
        const uint32_t pairs = dims / 2;
        for (uint32_t w = 0; w < pairs; ++w) {
            uint64_t h = mix64(hr + 0xD1B54A32D192ED03ull * (uint64_t)(w + 1));
            out[2 * w] = u24_to_unit((uint32_t)h);
            out[2 * w + 1] = u24_to_unit((uint32_t)(h >> 32));
        }
        if (dims & 1) out[dims - 1] = u24_to_unit((uint32_t)mix64(hr + 0xD1B54A32D192ED03ull * (uint64_t)(pairs + 1)));
    }
}

static void orc_synth_bit_row(uint64_t seed, int64_t rowid, uint32_t dims, uint8_t* dst) {
    const uint32_t rb = (dims + 7u) / 8u;
    for (uint32_t b0 = 0; b0 < rb; b0 += 8) {
        uint64_t h = synth_word(seed, rowid, b0 >> 3);
        for (uint32_t b = b0; b < rb && b < b0 + 8; ++b) {
            uint8_t v = (uint8_t)(h >> (8 * (b & 7)));
            uint32_t bits_left = dims - b * 8u;
            if (bits_left < 8u) v &= (uint8_t)((1u << bits_left) - 1u);
            dst[b] = v;
        }
    }
}

/* n rows with dense rowids first_rowid.. ; elem-specific encodings:
 *   f32: kind 0 uniform / 1 gauss4
 *   i8 : quantize_int8(uniform f32 row)   (config 3: vec_quantize_int8 corpus)
 *   bit: word w of the row = low/high half of synth_word(seed,rowid,w/2), LSB-first */
void orc_synth_rows(int elem, uint64_t seed, int64_t first_rowid, uint64_t n, uint32_t dims, int kind, void* out) {
    const uint32_t rb = orc_row_bytes(elem, dims);
#pragma omp parallel
    {
        float* tmp = malloc((size_t)dims * sizeof(float) + 16);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            int64_t rowid = first_rowid + i;
            char* dst = (char*)out + (size_t)i * rb;
            if (elem == ORC_F32) {
                orc_synth_f32_row(seed, rowid, dims, kind, (float*)dst);
            } else if (elem == ORC_I8) {
                orc_synth_f32_row(seed, rowid, dims, 0, tmp);
                orc_quantize_int8(tmp, dims, (int8_t*)dst);
            } else {
                orc_synth_bit_row(seed, rowid, dims, (uint8_t*)dst);
            }
        }
        free(tmp);
    }
}

/* ------------------------------------------------------------------------ */
/* Selection form of the exact scan, for sizes where keeping every (rowid, distance) pair and sorting all of them
 * (orc_knn above, the literal restatement of src/vtab.rs:2594-2620) is impractical: BASELINE.json's full configs
 * (10 M, 50 M, 500 M rows) and the CPU arm of bench.py.
 *
 * Same result by construction: a stable sort by d_f32 over ascending rowids followed by truncate(k) returns the k
 * smallest elements of the total order (order_bits(d_f32), position).  Every thread keeps the k smallest u64 keys
 * order_bits(d) << 32 | position of the rows it visited (a max-heap), the per-thread survivors are sorted and the
 * first k are the answer.  tests/test_oracle_golden.py checks orc_knn_select == orc_knn, ties and skips included. */

static inline float order_bits_inv(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float d;
    if (k == 0xFFFFFFFFu) return NAN;
    memcpy(&d, &u, 4);
    return d;
}

typedef struct {
    uint64_t* h; /* max-heap of at most k keys */
    uint32_t n, k;
} orc_heap_t;

static inline void heap_offer(orc_heap_t* hp, uint64_t key) {
    uint64_t* h = hp->h;
    if (hp->n < hp->k) {
        uint32_t i = hp->n++;
        while (i > 0) {
            uint32_t p = (i - 1) / 2;
            if (h[p] >= key) break;
            h[i] = h[p];
            i = p;
        }
        h[i] = key;
        return;
    }
    if (key >= h[0]) return;
    uint32_t i = 0;
    const uint32_t n = hp->n;
    for (;;) {
        uint32_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && h[c + 1] > h[c]) ++c;
        if (h[c] <= key) break;
        h[i] = h[c];
        i = c;
    }
    h[i] = key;
}

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* shared tail: merge the per-thread heaps of every query, decode */
static void select_finish(orc_heap_t* heaps, int nt, uint32_t nq, uint32_t k, const int64_t* rowids, int64_t first_rowid,
                          int64_t* out_rowids, float* out_dists, uint32_t* out_counts) {
    uint64_t* all = malloc(((size_t)nt * k + 1) * 8);
    for (uint32_t q = 0; q < nq; ++q) {
        size_t m = 0;
        for (int t = 0; t < nt; ++t) {
            orc_heap_t* hp = &heaps[(size_t)t * nq + q];
            memcpy(all + m, hp->h, (size_t)hp->n * 8);
            m += hp->n;
        }
        qsort(all, m, 8, cmp_u64);
        uint32_t cnt = (uint32_t)(m < k ? m : k);
        for (uint32_t j = 0; j < k; ++j) {
            if (j < cnt) {
                uint32_t pos = (uint32_t)all[j];
                out_rowids[(size_t)q * k + j] = rowids ? rowids[pos] : first_rowid + (int64_t)pos;
                out_dists[(size_t)q * k + j] = order_bits_inv((uint32_t)(all[j] >> 32));
            } else {
                out_rowids[(size_t)q * k + j] = -1;
                out_dists[(size_t)q * k + j] = INFINITY;
            }
        }
        if (out_counts) out_counts[q] = cnt;
    }
    free(all);
}

static orc_heap_t* heaps_alloc(int nt, uint32_t nq, uint32_t k) {
    orc_heap_t* heaps = calloc((size_t)nt * nq, sizeof(orc_heap_t));
    for (size_t i = 0; i < (size_t)nt * nq; ++i) {
        heaps[i].h = malloc(((size_t)k + 1) * 8);
        heaps[i].k = k;
    }
    return heaps;
}
static void heaps_free(orc_heap_t* heaps, int nt, uint32_t nq) {
    for (size_t i = 0; i < (size_t)nt * nq; ++i) free(heaps[i].h);
    free(heaps);
}

int orc_knn_select(int elem, uint32_t dims, const int64_t* rowids, const void* vectors, const uint8_t* skip, uint64_t n,
                   const void* queries, uint32_t nq, uint32_t k, int metric, int64_t* out_rowids, float* out_dists,
                   uint32_t* out_counts) {
    if (elem < 0 || elem > 2 || metric < 0 || metric > 3 || dims == 0) return ORC_ERR_INVALID_PARAM;
    if (!orc_metric_supported(elem, metric)) return ORC_ERR_UNSUPPORTED;
    if (n >= 0xFFFFFFFFull) return ORC_ERR_INVALID_PARAM;
    const uint32_t rb = orc_row_bytes(elem, dims);
    const int nt = orc_num_threads();
    orc_heap_t* heaps = heaps_alloc(nt, nq, k ? k : 1);
    if (k)
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            if (skip && skip[i]) continue;
            const char* row = (const char*)vectors + (size_t)i * rb;
            for (uint32_t q = 0; q < nq; ++q) {
                float d = dist_dispatch(elem, metric, (const char*)queries + (size_t)q * rb, row, dims);
                heap_offer(&heaps[(size_t)t * nq + q], ((uint64_t)order_bits(d) << 32) | (uint32_t)i);
            }
        }
    }
    if (k) select_finish(heaps, nt, nq, k, rowids, 1, out_rowids, out_dists, out_counts);
    else if (out_counts) memset(out_counts, 0, (size_t)nq * 4);
    heaps_free(heaps, nt, nq);
    return ORC_OK;
}

/* The same scan over the SYNTHETIC corpus value(seed, rowid, j) without materialising it: every thread regenerates
 * one row at a time into a private buffer and scores it against all nq queries.  Rows are first_rowid .. first_rowid +
 * n - 1 (dense).  This is how the full BASELINE sizes are checked on the CPU (10 M x 768 f32 is 30 GB, 50 M x 1024 i8
 * 51 GB, 500 M x 1024 bit 64 GB). */
int orc_knn_synth(int elem, uint32_t dims, uint64_t seed, int64_t first_rowid, uint64_t n, int kind, const void* queries,
                  uint32_t nq, uint32_t k, int metric, int64_t* out_rowids, float* out_dists, uint32_t* out_counts) {
    if (elem < 0 || elem > 2 || metric < 0 || metric > 3 || dims == 0 || k == 0) return ORC_ERR_INVALID_PARAM;
    if (!orc_metric_supported(elem, metric)) return ORC_ERR_UNSUPPORTED;
    if (n >= 0xFFFFFFFFull) return ORC_ERR_INVALID_PARAM;
    const uint32_t rb = orc_row_bytes(elem, dims);
    const int nt = orc_num_threads();
    orc_heap_t* heaps = heaps_alloc(nt, nq, k);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        float* tmp = malloc((size_t)dims * sizeof(float) + 64);
        char* row = malloc((size_t)rb + 64);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const int64_t rowid = first_rowid + i;
            if (elem == ORC_F32) {
                orc_synth_f32_row(seed, rowid, dims, kind, (float*)row);
            } else if (elem == ORC_I8) {
                orc_synth_f32_row(seed, rowid, dims, 0, tmp);
                orc_quantize_int8(tmp, dims, (int8_t*)row);
            } else {
                orc_synth_bit_row(seed, rowid, dims, (uint8_t*)row);
            }
            for (uint32_t q = 0; q < nq; ++q) {
                float d = dist_dispatch(elem, metric, (const char*)queries + (size_t)q * rb, row, dims);
                heap_offer(&heaps[(size_t)t * nq + q], ((uint64_t)order_bits(d) << 32) | (uint32_t)i);
            }
        }
        free(tmp);
        free(row);
    }
    select_finish(heaps, nt, nq, k, NULL, first_rowid, out_rowids, out_dists, out_counts);
    heaps_free(heaps, nt, nq);
    return ORC_OK;
}
