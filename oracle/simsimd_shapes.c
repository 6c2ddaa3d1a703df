/*
 * simsimd_shapes.c — CPU ORACLE, TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The reference computes f32 L2, f32 cosine and i8 cosine inside simsimd 6.5.16 (Cargo.toml:22,
 * Cargo.lock:725-731; call sites src/distance/scalar.rs:17, :48, :94), which is not vendored under the reference tree
 * and cannot be built here.  The canonical order of vecgpu_oracle.c is ONE hypothesis about that crate's arithmetic.
 * This file restates the OTHER accumulation / finishing shapes SimSIMD v6 publishes for its back ends, as far as
 * they are known (recalled from upstream include/simsimd/spatial.h — unverifiable here, hence "shapes"), so that
 * tests/test_simsimd_gap.py can bound how far the choice of back end can move a distance or flip a returned rowid:
 *
 *   accumulation shapes (f32 L2^2 and the three cosine sums)
 *     0  canonical       16 f32 FMA lanes, i -> lane i%16, tree i+8,+4,+2,+1        (vecgpu_oracle.c, "skylake"-like)
 *     1  serial          one f32 accumulator, d2 += (a-b)*(a-b), no FMA             (simsimd_*_f32_serial)
 *     2  serial_fma      the same loop contracted to FMA by the compiler
 *     3  lanes8_f64red   8 f32 FMA lanes, horizontal sum carried out in f64         (haswell: _simsimd_reduce_f32x8_haswell)
 *     4  lanes16_hadd    16 f32 FMA lanes, reduce 512->256->128 then two hadd steps (skylake: _simsimd_reduce_f32x16_skylake)
 *     5  lanes4          4 f32 FMA lanes, pairwise horizontal add                   (neon: vaddvq_f32)
 *     6  f64             every product and sum in f64                               (simsimd_*_f32_accurate)
 *   cosine finish shapes
 *     0  ieee_f64        1 - ab / (sqrt(a2) * sqrt(b2)) in f64                      (vecgpu_oracle.c)
 *     1  ieee_f32        the same in f32 with sqrtf
 *     2  rsqrt12_nr_f64  rsqrt estimate (_mm_rsqrt_ps, 12 bit) + one Newton step in f64 (haswell _simsimd_cos_normalize_f64_haswell)
 *     3  rsqrt14_nr_f64  _mm_rsqrt14_pd + one Newton step in f64                     (skylake)
 *     4  rsqrt12_nr_f32  _mm_rsqrt_ps + one Newton step in f32
 *   Finish shapes 2-4 use the host's own estimate instructions (available only on x86 with SSE / AVX-512VL);
 *   orc_shape_supported() says which exist on the machine the test runs on.
 *
 * All shapes share the zero rules (a2 == 0 && b2 == 0 -> 0; ab == 0 -> 1) and the clamp at 0.
 */
#define _GNU_SOURCE
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_ACC 7
#define N_FIN 5

static void acc3(int shape, const float* a, const float* b, uint32_t d, int l2, double* o_ab, double* o_a2, double* o_b2) {
    /* l2 != 0: o_ab = sum (a-b)^2 only */
    const int L = shape == 3 ? 8 : shape == 4 ? 16 : shape == 5 ? 4 : shape == 0 ? 16 : 1;
    if (shape == 6) {
        double ab = 0, a2 = 0, b2 = 0;
        for (uint32_t i = 0; i < d; ++i) {
            double x = a[i], y = b[i];
            if (l2) {
                double t = x - y;
                ab += t * t;
            } else {
                ab += x * y;
                a2 += x * x;
                b2 += y * y;
            }
        }
        *o_ab = ab, *o_a2 = a2, *o_b2 = b2;
        return;
    }
    if (shape == 1 || shape == 2) {
        volatile float vab = 0, va2 = 0, vb2 = 0; /* volatile: keep the compiler from re-associating or contracting */
        float ab = 0, a2 = 0, b2 = 0;
        for (uint32_t i = 0; i < d; ++i) {
            float x = a[i], y = b[i];
            if (l2) {
                float t = x - y;
                if (shape == 2) ab = fmaf(t, t, ab);
                else { vab = t * t; ab = ab + vab; }
            } else if (shape == 2) {
                ab = fmaf(x, y, ab);
                a2 = fmaf(x, x, a2);
                b2 = fmaf(y, y, b2);
            } else {
                vab = x * y; ab = ab + vab;
                va2 = x * x; a2 = a2 + va2;
                vb2 = y * y; b2 = b2 + vb2;
            }
        }
        *o_ab = ab, *o_a2 = a2, *o_b2 = b2;
        return;
    }
    float lab[16] = {0}, la[16] = {0}, lb[16] = {0};
    for (uint32_t i = 0; i < d; ++i) {
        const int j = (int)(i % (uint32_t)L);
        float x = a[i], y = b[i];
        if (l2) {
            float t = x - y;
            lab[j] = fmaf(t, t, lab[j]);
        } else {
            lab[j] = fmaf(x, y, lab[j]);
            la[j] = fmaf(x, x, la[j]);
            lb[j] = fmaf(y, y, lb[j]);
        }
    }
    float* v[3] = {lab, la, lb};
    double out[3];
    for (int s = 0; s < 3; ++s) {
        float* l = v[s];
        if (shape == 0) {
            for (int i = 0; i < 8; ++i) l[i] += l[i + 8];
            for (int i = 0; i < 4; ++i) l[i] += l[i + 4];
            for (int i = 0; i < 2; ++i) l[i] += l[i + 2];
            out[s] = (double)(l[0] + l[1]);
        } else if (shape == 3) {
            /* low 128-bit half + high half in f64, then (s0+s2)+(s1+s3)-style f64 reduce */
            double s0 = (double)l[0] + (double)l[4], s1 = (double)l[1] + (double)l[5];
            double s2 = (double)l[2] + (double)l[6], s3 = (double)l[3] + (double)l[7];
            out[s] = (s0 + s2) + (s1 + s3);
        } else if (shape == 4) {
            for (int i = 0; i < 8; ++i) l[i] += l[i + 8];
            for (int i = 0; i < 4; ++i) l[i] += l[i + 4];
            out[s] = (double)((l[0] + l[1]) + (l[2] + l[3])); /* two hadd steps */
        } else { /* 4 lanes, pairwise */
            out[s] = (double)((l[0] + l[1]) + (l[2] + l[3]));
        }
    }
    *o_ab = out[0], *o_a2 = out[1], *o_b2 = out[2];
}

#if defined(__x86_64__) && defined(__GNUC__)
static float rsqrt12(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
__attribute__((target("avx512f,avx512vl"))) static double rsqrt14(double x) {
    return _mm_cvtsd_f64(_mm_rsqrt14_pd(_mm_set1_pd(x)));
}
static int have_r14(void) { return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl"); }
#define HAVE_X86 1
#else
#define HAVE_X86 0
#endif

int orc_shape_supported(int acc, int fin) {
    if (acc < 0 || acc >= N_ACC || fin < 0 || fin >= N_FIN) return 0;
#if HAVE_X86
    if (fin == 3) return have_r14();
    return 1;
#else
    return fin <= 1;
#endif
}

static float finish(int fin, double ab, double a2, double b2) {
    if (a2 == 0.0 && b2 == 0.0) return 0.0f;
    if (ab == 0.0) return 1.0f;
    double r;
    switch (fin) {
        default:
        case 0: r = 1.0 - ab / (sqrt(a2) * sqrt(b2)); break;
        case 1: {
            float f = 1.0f - (float)ab / (sqrtf((float)a2) * sqrtf((float)b2));
            r = f;
            break;
        }
#if HAVE_X86
        case 2:
        case 3: {
            double ra = fin == 2 ? (double)rsqrt12((float)a2) : rsqrt14(a2);
            double rb = fin == 2 ? (double)rsqrt12((float)b2) : rsqrt14(b2);
            ra = 1.5 * ra + (a2 * -0.5 * ra) * (ra * ra);
            rb = 1.5 * rb + (b2 * -0.5 * rb) * (rb * rb);
            r = 1.0 - ab * ra * rb;
            break;
        }
        case 4: {
            float fa = (float)a2, fb = (float)b2;
            float ra = rsqrt12(fa), rb = rsqrt12(fb);
            ra = ra * (1.5f - 0.5f * fa * ra * ra);
            rb = rb * (1.5f - 0.5f * fb * rb * rb);
            r = 1.0f - (float)ab * ra * rb;
            break;
        }
#endif
    }
    return (float)(r > 0.0 ? r : 0.0);
}

/* f32 L2 under an accumulation shape: sqrtf((float)sum) as src/distance/scalar.rs:17-20 */
/* f32 cosine under (accumulation, finish) shapes: (float)result as scalar.rs:48-51 */
/* metric: 0 = L2, 2 = cosine.  out[i] for the n rows of `vectors`. */
int orc_shape_distances_f32(int acc, int fin, uint32_t dims, const float* vectors, uint64_t n, const float* query, int metric,
                            float* out) {
    if (!orc_shape_supported(acc, metric == 2 ? fin : 0)) return 3;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double ab, a2, b2;
        const float* row = vectors + (size_t)i * dims;
        if (metric == 0) {
            acc3(acc, query, row, dims, 1, &ab, &a2, &b2);
            out[i] = sqrtf((float)ab);
        } else {
            acc3(acc, query, row, dims, 0, &ab, &a2, &b2);
            out[i] = finish(fin, ab, a2, b2);
        }
    }
    return 0;
}

/* i8 cosine: exact integer sums (as every SimSIMD back end), finish shape varies — scalar.rs:94-97 */
int orc_shape_distances_i8cos(int fin, uint32_t dims, const int8_t* vectors, uint64_t n, const int8_t* query, float* out) {
    if (!orc_shape_supported(0, fin)) return 3;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const int8_t* row = vectors + (size_t)i * dims;
        int64_t ab = 0, a2 = 0, b2 = 0;
        for (uint32_t j = 0; j < dims; ++j) {
            int32_t x = query[j], y = row[j];
            ab += x * y, a2 += x * x, b2 += y * y;
        }
        out[i] = finish(fin, (double)ab, (double)a2, (double)b2);
    }
    return 0;
}
