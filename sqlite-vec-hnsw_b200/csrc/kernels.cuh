// kernels.cuh — sm_100a device code of libvecgpu.so.
//
// Replaces, on the GPU, the arithmetic of src/distance/scalar.rs:12-112 (7
// distance functions), the scan + sort + truncate of brute_force_search
// (src/vtab.rs:2594-2620) and the neighbour-scoring loop of search_layer
// (src/hnsw/search.rs:501-513) of the reference.  Not a translation: the
// reference scores one pair per call through simsimd behind one SQLite lookup
// per row; here a persistent grid streams the HBM-resident slab through shared
// memory with bulk async copies (cp.async.bulk -> SASS UBLKCP) completing on
// mbarriers, scores QB queries per pass in a fixed ("canonical") accumulation
// order and keeps a fused per-warp top-k, so no distance array is written.
#pragma once
#include "xpush.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace vg {

// ---------------------------------------------------------------------------
// ranking keys: (order_bits(d_f32) << 32) | row position.  u64 '<' on keys ==
// (distance_f32, rowid) ascending because slab rows are stored in ascending
// rowid order (src/shadow.rs:856 order; stable sort of src/vtab.rs:2619).
// NaN ranks after +inf.
// ---------------------------------------------------------------------------
static constexpr uint64_t KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint32_t order_bits(float d) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(d);
#else
    uint32_t u;
    memcpy(&u, &d, 4);
#endif
    if (d != d) return 0xFFFFFFFFu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float order_bits_inv(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    if (k == 0xFFFFFFFFu) u = 0x7FC00000u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t pos) {
    return ((uint64_t)order_bits(d) << 32) | (uint64_t)pos;
}

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, no tensor map).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------------
// metric traits.  A "unit" is 16 bytes of a row.  LPR lanes cooperate on one
// row; lane g of the group handles units g, g+LPR, g+2*LPR, ... in increasing
// order.  For f32 L2/cosine LPR=4 and the lane's 4 accumulators are canonical
// lanes 4g..4g+3 of SURVEY §A.4 (element i -> lane i%16), reduced with the
// i<->i+8, i<->i+4, i<->i+2, i<->i+1 tree; for f32 L1 LPR=1 (strict order,
// src/distance/scalar.rs:31-35); integer metrics are order-free.
// QC = per-query constant computed once (|q|^2).
// ---------------------------------------------------------------------------
// canonical tree over the 16 lanes held as 4 lanes x 4 accumulators
__device__ __forceinline__ float canon_tree(float a0, float a1, float a2, float a3) {
    a0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, 2));  // l[i] + l[i+8]
    a1 = __fadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, 2));
    a2 = __fadd_rn(a2, __shfl_xor_sync(0xffffffffu, a2, 2));
    a3 = __fadd_rn(a3, __shfl_xor_sync(0xffffffffu, a3, 2));
    a0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, 1));  // + l[i+4]
    a1 = __fadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, 1));
    a2 = __fadd_rn(a2, __shfl_xor_sync(0xffffffffu, a2, 1));
    a3 = __fadd_rn(a3, __shfl_xor_sync(0xffffffffu, a3, 1));
    return __fadd_rn(__fadd_rn(a0, a2), __fadd_rn(a1, a3));    // (l0+l2)+(l1+l3)
}
__device__ __forceinline__ int group4_sum_i(int v) {
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// cosine finish in f64 with IEEE sqrt/div
__device__ __forceinline__ float cos_finish(double ab, double a2, double b2) {
    if (a2 == 0.0 && b2 == 0.0) return 0.0f;
    if (ab == 0.0) return 1.0f;
    double r = __dsub_rn(1.0, __ddiv_rn(ab, __dmul_rn(__dsqrt_rn(a2), __dsqrt_rn(b2))));
    return __double2float_rn(r > 0.0 ? r : 0.0);
}

template <int QB>
struct F32L2 {
    static constexpr int LPR = 4;
    struct Acc {
        float s[QB][4];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) a.s[q][e] = 0.0f;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
        const float xf[4] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const float qf[4] = {__uint_as_float(q[i].x), __uint_as_float(q[i].y), __uint_as_float(q[i].z),
                                 __uint_as_float(q[i].w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = __fsub_rn(qf[e], xf[e]);
                a.s[i][e] = __fmaf_rn(t, t, a.s[i][e]);
            }
        }
    }
    // returns the distance of query i; all lanes of the group get the same value
    __device__ static float finish(const Acc& a, int i, const float*) {
        float s = canon_tree(a.s[i][0], a.s[i][1], a.s[i][2], a.s[i][3]);
        return __fsqrt_rn(s);  // src/distance/scalar.rs:20: (s as f32).sqrt()
    }
    static constexpr bool HAS_QC = false;
    static constexpr bool ORDER_FREE = false;
    static constexpr bool HAS_BOUND = false;
};

template <int QB>
struct F32Cos {
    static constexpr int LPR = 4;
    struct Acc {
        float b2[4];
        float ab[QB][4];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int e = 0; e < 4; ++e) a.b2[e] = 0.0f;
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) a.ab[q][e] = 0.0f;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
        const float xf[4] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) a.b2[e] = __fmaf_rn(xf[e], xf[e], a.b2[e]);
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const float qf[4] = {__uint_as_float(q[i].x), __uint_as_float(q[i].y), __uint_as_float(q[i].z),
                                 __uint_as_float(q[i].w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) a.ab[i][e] = __fmaf_rn(qf[e], xf[e], a.ab[i][e]);
        }
    }
    struct Red {
        float ab, b2;
    };
    __device__ static Red reduce(const Acc& a, int i) {
        Red r;
        r.b2 = canon_tree(a.b2[0], a.b2[1], a.b2[2], a.b2[3]);
        r.ab = canon_tree(a.ab[i][0], a.ab[i][1], a.ab[i][2], a.ab[i][3]);
        return r;
    }
    // cheap f32 value guaranteed <= the exact f64-finished distance (|error| of the f32 path <= 6e-7 for
    // |cos| <= 1; margin 2e-6).  -inf when a2*b2 leaves the safe range: the caller then takes the exact path.
    __device__ static float bound(const Red& r, float a2) {
        const float p = __fmul_rn(a2, r.b2);
        if (!(p > 1e-30f && p < 1e30f)) return __int_as_float(0xFF800000);
        const float c = __fmul_rn(r.ab, rsqrtf(p));
        return __fsub_rn(__fsub_rn(1.0f, c), 2e-6f);
    }
    __device__ static float exact(const Red& r, float a2) { return cos_finish((double)r.ab, (double)a2, (double)r.b2); }
    __device__ static float finish(const Acc& a, int i, const float* qc) { return exact(reduce(a, i), qc[i]); }
    static constexpr bool HAS_QC = true;
    static constexpr bool ORDER_FREE = false;
    static constexpr bool HAS_BOUND = true;
};

template <int QB>
struct F32L1 {
    static constexpr int LPR = 1;
    struct Acc {
        float s[QB];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int q = 0; q < QB; ++q) a.s[q] = 0.0f;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
        const float xf[4] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const float qf[4] = {__uint_as_float(q[i].x), __uint_as_float(q[i].y), __uint_as_float(q[i].z),
                                 __uint_as_float(q[i].w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) a.s[i] = __fadd_rn(a.s[i], fabsf(__fsub_rn(qf[e], xf[e])));
        }
    }
    __device__ static float finish(const Acc& a, int i, const float*) { return a.s[i]; }
    static constexpr bool HAS_QC = false;
    static constexpr bool ORDER_FREE = false;
    static constexpr bool HAS_BOUND = false;
};

// int8: exact int32 partial sums by dp4a; |q|^2 arrives as a float-encoded pair in qc (hi/lo split not
// needed: we pass it as int via __float_as_int).
template <int QB, bool COS>
struct I8Dot {
    static constexpr int LPR = 4;
    struct Acc {  // four independent partial sums per quantity: IDP.4A chains stay short
        int bb[4];
        int ab[QB][4];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int e = 0; e < 4; ++e) a.bb[e] = 0;
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) a.ab[q][e] = 0;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
        const int xw[4] = {(int)x.x, (int)x.y, (int)x.z, (int)x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) a.bb[e] = __dp4a(xw[e], xw[e], a.bb[e]);
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const int qw[4] = {(int)q[i].x, (int)q[i].y, (int)q[i].z, (int)q[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) a.ab[i][e] = __dp4a(qw[e], xw[e], a.ab[i][e]);
        }
    }
    struct Red {
        int bb, ab;
    };
    __device__ static Red reduce(const Acc& a, int i) {
        Red r;
        r.bb = group4_sum_i((a.bb[0] + a.bb[1]) + (a.bb[2] + a.bb[3]));
        r.ab = group4_sum_i((a.ab[i][0] + a.ab[i][1]) + (a.ab[i][2] + a.ab[i][3]));
        return r;
    }
    // cheap f32 lower bound of the exact distance (see F32Cos::bound); qa2 carries |q|^2 as int bits
    __device__ static float bound(const Red& r, float qa2) {
        const long long a2 = (long long)__float_as_int(qa2);
        if (COS) {
            const float p = __fmul_rn((float)a2, (float)r.bb);
            if (!(p > 1e-30f && p < 1e30f)) return __int_as_float(0xFF800000);
            const float c = __fmul_rn((float)r.ab, rsqrtf(p));
            return __fsub_rn(__fsub_rn(1.0f, c), 3e-6f);
        }
        const long long s = a2 + (long long)r.bb - 2ll * (long long)r.ab;
        return __fmul_rn(__fsqrt_rn((float)s), 0.9999995f);  // sqrtf((float)s) is within 2^-23 of sqrt(s)
    }
    __device__ static float exact(const Red& r, float qa2) {
        const long long a2 = (long long)__float_as_int(qa2);
        if (COS) return cos_finish((double)r.ab, (double)a2, (double)r.bb);
        const long long s = a2 + (long long)r.bb - 2ll * (long long)r.ab;  // == sum (a-b)^2 exactly
        // src/distance/scalar.rs:65: distance.sqrt() as f32  (f64 sqrt, then cast)
        return __double2float_rn(__dsqrt_rn((double)s));
    }
    __device__ static float finish(const Acc& a, int i, const float* qc) { return exact(reduce(a, i), qc[i]); }
    static constexpr bool HAS_QC = true;
    static constexpr bool ORDER_FREE = true;
    static constexpr bool HAS_BOUND = true;
};

__device__ __forceinline__ unsigned absdiff_s8x4(unsigned a, unsigned b) {
    // per-byte |a-b| for signed bytes, as unsigned bytes 0..255
    unsigned mx = __vmaxs4(a, b), mn = __vmins4(a, b);
    return __vsub4(mx, mn);
}

template <int QB>
struct I8L1 {
    static constexpr int LPR = 4;
    struct Acc {
        int s[QB][4];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) a.s[q][e] = 0;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
        const unsigned xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const unsigned qw[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                a.s[i][e] = (int)__dp4a(absdiff_s8x4(qw[e], xw[e]), 0x01010101u, (unsigned)a.s[i][e]);
        }
    }
    __device__ static float finish(const Acc& a, int i, const float*) {
        return (float)group4_sum_i((a.s[i][0] + a.s[i][1]) + (a.s[i][2] + a.s[i][3]));
    }
    static constexpr bool HAS_QC = false;
    static constexpr bool ORDER_FREE = true;
    static constexpr bool HAS_BOUND = false;
};

// Hamming: one thread per row (rows are short: bit[1024] = 128 B), exact popcount
template <int QB>
struct BitHamming {
    static constexpr int LPR = 1;
    struct Acc {  // two independent add chains
        int s[QB][2];
    };
    __device__ static void init(Acc& a) {
#pragma unroll
        for (int q = 0; q < QB; ++q) a.s[q][0] = a.s[q][1] = 0;
    }
    __device__ static void step(Acc& a, uint4 x, const uint4 (&q)[QB]) {
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            a.s[i][0] += __popc(x.x ^ q[i].x) + __popc(x.y ^ q[i].y);
            a.s[i][1] += __popc(x.z ^ q[i].z) + __popc(x.w ^ q[i].w);
        }
    }
    __device__ static float finish(const Acc& a, int i, const float*) { return (float)(a.s[i][0] + a.s[i][1]); }
    static constexpr bool HAS_QC = false;
    static constexpr bool ORDER_FREE = true;
    static constexpr bool HAS_BOUND = false;
};

// |q|^2 of one query in the representation finish() expects; executed by one
// 4-lane group (lanes 0..3 of a warp) over a zero-padded query in shared or
// global memory.  kind: 0 = f32 canonical sum of squares, 1 = int8 exact.
__device__ __forceinline__ float query_const(const uint4* q, uint32_t units, int g, int kind) {
    if (kind == 0) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (uint32_t u = g; u < units; u += 4) {
            uint4 v = q[u];
            const float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = __fmaf_rn(f[e], f[e], a[e]);
        }
        return canon_tree(a[0], a[1], a[2], a[3]);
    } else {
        int s = 0;
        for (uint32_t u = g; u < units; u += 4) {
            uint4 v = q[u];
            s = __dp4a((int)v.x, (int)v.x, s);
            s = __dp4a((int)v.y, (int)v.y, s);
            s = __dp4a((int)v.z, (int)v.z, s);
            s = __dp4a((int)v.w, (int)v.w, s);
        }
        return __int_as_float(group4_sum_i(s));
    }
}

// ---------------------------------------------------------------------------
// per-warp top-k list in shared memory: unsorted buffer of K keys + tracked
// maximum.  All lanes call with warp-uniform arguments.
// ---------------------------------------------------------------------------
struct ListHdr {
    uint64_t tau;     // current admission bound: max key in list when full, else KEY_NONE
    uint32_t maxpos;  // position of tau in list when full
    uint32_t cnt;
};

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void list_refresh_max(uint64_t* list, ListHdr* hdr, uint32_t K, int lane) {
    uint64_t best = 0;
    uint32_t bpos = 0;
    for (uint32_t i = lane; i < K; i += 32) {
        uint64_t v = list[i];
        if (v >= best) {
            best = v;
            bpos = i;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        uint64_t ob = shfl_xor_u64(best, m);
        uint32_t op = __shfl_xor_sync(0xffffffffu, bpos, m);
        if (ob > best || (ob == best && op > bpos)) {
            best = ob;
            bpos = op;
        }
    }
    if (lane == 0) {
        hdr->tau = best;
        hdr->maxpos = bpos;
    }
    __syncwarp();
}

// insert a key known to be < hdr->tau (or list not yet full)
__device__ __forceinline__ void list_insert(uint64_t* list, ListHdr* hdr, uint32_t K, uint64_t key, int lane) {
    uint32_t cnt = hdr->cnt;
    __syncwarp();
    if (cnt < K) {
        if (lane == 0) {
            list[cnt] = key;
            hdr->cnt = cnt + 1;
        }
        __syncwarp();
        if (cnt + 1 == K) list_refresh_max(list, hdr, K, lane);
    } else {
        if (lane == 0) list[hdr->maxpos] = key;
        __syncwarp();
        list_refresh_max(list, hdr, K, lane);
    }
}

// offer the keys held by the lanes in `mask` (one key per lane) to the list
__device__ __forceinline__ void list_offer(uint64_t* list, ListHdr* hdr, uint32_t K, uint64_t key, bool want, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, want && key < hdr->tau);
    while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        uint64_t kq = shfl_u64(key, src);
        if (kq < hdr->tau) list_insert(list, hdr, K, kq, lane);
    }
}

// ---------------------------------------------------------------------------
// Warp-wide sort of 256 u64 keys held 8 per lane (element index = lane * 8 + r), ascending: a bitonic network that
// lives entirely in registers — strides below 8 are register-to-register compare-exchanges, larger strides are lane
// shuffles; the "flip" form of the network (first stage of every size compares i with i ^ (size - 1)) needs no
// direction flags.  ~1.2 k instructions per lane.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t u64max(uint64_t a, uint64_t b) { return a < b ? b : a; }

__device__ __forceinline__ void warp_sort256(uint64_t (&v)[8], int lane) {
#pragma unroll
    for (int size = 2; size <= 256; size <<= 1) {
        if (size <= 8) {  // element i against i ^ (size - 1), both in this lane
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int pr = r ^ (size - 1);
                if (pr > r) {
                    const uint64_t a = v[r], b = v[pr];
                    v[r] = u64min(a, b);
                    v[pr] = u64max(a, b);
                }
            }
        } else {          // partner lane = lane ^ (size/8 - 1), partner register = 7 - r
            const bool keep_min = (lane & (size / 16)) == 0;
            uint64_t o[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) o[r] = shfl_xor_u64(v[7 - r], size / 8 - 1);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = keep_min ? u64min(v[r], o[r]) : u64max(v[r], o[r]);
        }
#pragma unroll
        for (int stride = size / 4; stride >= 1; stride >>= 1) {  // element i against i ^ stride
            if (stride >= 8) {
                const bool keep_min = (lane & (stride / 8)) == 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint64_t o = shfl_xor_u64(v[r], stride / 8);
                    v[r] = keep_min ? u64min(v[r], o) : u64max(v[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if ((r & stride) == 0) {
                        const uint64_t a = v[r], b = v[r ^ stride];
                        v[r] = u64min(a, b);
                        v[r ^ stride] = u64max(a, b);
                    }
            }
        }
    }
}

// in-place ascending bitonic sort of n (power of two) u64 keys in shared memory by the whole CTA
// ONE out-of-line copy of the register sorter for the scan's tails (per-CTA list merge, fused final merge, merge kernels):
// every CTA runs it on its own lists just before the last CTA needs it for the final merge, so its ~1.2 k instructions are
// already in the instruction caches there.  (Measured with the globaltimer stamps of tools/scan_timeline.py: with a
// separately inlined copy the final merge took ~35 us after a 3.8 GB scan — cold code fetched behind a thrashed L2 —
// against ~12 us on a table that fits L2.)
__device__ __noinline__ void warp_sort256_shared(uint64_t* v, int lane) {
    uint64_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i];
    warp_sort256(r, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r[i];
}

__device__ __forceinline__ void block_bitonic_sort(uint64_t* keys, uint32_t n) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// K1/K3/K4 (+K2a when QB>1): the streaming scan.
//
// Persistent CTAs (one per SM) of C warps.  The slab is cut into tiles of RS
// rows; tile t belongs to CTA t % gridDim.x and the j-th tile of a CTA to warp
// j % C.  Every warp owns a private ring of D shared-memory stages and is its
// own producer: after it has finished reading a stage it immediately issues
// the bulk copy of the tile D steps ahead into that stage.  Hence there is one
// mbarrier per stage (full[s], completed by the copy's transaction bytes), no
// "empty" barrier, no producer warp and no cross-warp protocol; a barrier is
// only ever waited on, in phase order, by the warp that armed it.
//   contig mode : ONE cp.async.bulk per stage (RS*row_stride contiguous bytes,
//                 >= 8 KB, so the fixed per-copy cost of the copy engine — about
//                 70 cycles, measured — is amortised).  Bank conflicts are
//                 avoided without padding by rotating the order in which a lane
//                 reads the 16-byte units of a 128-byte segment (legal for the
//                 order-free integer metrics); the canonical-order f32 metrics
//                 accept a 2-way conflict (shared memory has >2x headroom).
//   per-row mode: one bulk copy per row chunk (>= 1.5 KB) into a padded stage;
//                 used for f32 L1 (strict order, thread-per-row needs odd unit
//                 strides) and for rows too wide for RPW whole rows per stage.
// History: a shared round-robin ring with a producer warp failed sporadically —
// with copies completing out of order a warp could observe the parity of an
// older, still incomplete phase (mbarrier parity waits alias every two phases).
// ---------------------------------------------------------------------------
struct MergeParams {
    const uint64_t* keys;  // [nq][n_cand]
    uint64_t n_cand;
    uint32_t k;
    uint32_t kp2;             // next power of two >= k
    const int64_t* rowids;    // position -> rowid, or nullptr when dense
    int64_t first_rowid;      // dense: rowid = first_rowid + position
    int64_t* out_rowids;      // [nq][k]
    float* out_dists;         // [nq][k]
    uint32_t* out_counts;     // [nq] or nullptr
    int64_t pad_rowid;        // value for unused slots (-1 host API, INT64_MAX device API)
};


// one result slot: key -> (rowid, distance) or padding; returns 1 for a real entry
__device__ __forceinline__ uint32_t merge_emit(const MergeParams& p, uint32_t q, uint32_t j, uint64_t key) {
    const size_t o = (size_t)q * p.k + j;
    if (key == KEY_NONE) {
        p.out_rowids[o] = p.pad_rowid;
        p.out_dists[o] = __int_as_float(0x7F800000);
        return 0;
    }
    const uint32_t pos = (uint32_t)key;
    p.out_rowids[o] = p.rowids ? p.rowids[pos] : p.first_rowid + (int64_t)pos;
    p.out_dists[o] = order_bits_inv((uint32_t)(key >> 32));
    return 1;
}

// Final selection for query q by ONE CTA (any block size that is a multiple of 32): the n_cand partial keys -> k smallest,
// ascending, decoded.  `scratch` is shared memory: 256 keys are enough for the register path (k <= 32 and at most 2048
// candidates: eight 256-key register sorts, then one more over their 8 x 32 survivors), np2 keys for the bitonic path.
// Used by the merge kernels and by the fused tail of scan_kernel (last CTA done).  Ends with a __syncthreads().
__device__ __forceinline__ void final_merge_cta(const MergeParams& p, uint32_t q, uint64_t* scratch, uint32_t np2, bool allow_small) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint64_t* src = p.keys + (size_t)q * p.n_cand;
    if (allow_small && p.k <= 32 && p.n_cand <= 2048) {
        for (int w = warp; w < 8; w += nwarps) {
            uint64_t v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t j = (uint32_t)w * 256 + (uint32_t)lane * 8 + r;
                v[r] = j < p.n_cand ? __ldcg(src + j) : KEY_NONE;
            }
            warp_sort256_shared(v, lane);
            if (lane < 4) {
#pragma unroll
                for (int r = 0; r < 8; ++r) scratch[w * 32 + lane * 8 + r] = v[r];
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint64_t v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = scratch[lane * 8 + r];
            warp_sort256_shared(v, lane);
            uint32_t cnt = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t j = (uint32_t)lane * 8 + r;
                if (j < p.k) cnt += merge_emit(p, q, j, v[r]);
            }
            if (p.out_counts) {
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
                if (lane == 0) p.out_counts[q] = cnt;
            }
        }
        __syncthreads();
        return;
    }
    for (uint32_t j = threadIdx.x; j < np2; j += blockDim.x) scratch[j] = j < p.n_cand ? __ldcg(src + j) : KEY_NONE;
    __syncthreads();
    block_bitonic_sort(scratch, np2);
    __shared__ uint32_t fm_total;
    if (threadIdx.x == 0) fm_total = 0;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) cnt += merge_emit(p, q, j, j < np2 ? scratch[j] : KEY_NONE);
    if (cnt) atomicAdd(&fm_total, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && p.out_counts) p.out_counts[q] = fm_total;
    __syncthreads();
}

// Fused tail of the streaming scan: when `counter` is set, the CTA that finishes LAST (atomic ticket per query pass) runs
// the final merge of its pass's queries itself — no second launch — and, for a sharded slab, pushes the result into the
// peers' gather buffers (xpush.cuh).  The counter re-arms itself.
struct ScanTail {
    uint32_t* counter;   // [gridDim.y] zero-initialised tickets (nullptr: no tail at all); the last CTA of a pass re-arms them
    uint32_t* dyn;       // [gridDim.y] zero-initialised tile counters: tiles are handed out dynamically (nullptr: static round robin)
    uint32_t static_rounds;  // with dyn: every warp first takes this many tiles of the static round robin, the rest is dynamic
    uint32_t do_merge;   // 1: the last CTA runs the final merge (no separate merge launch)
    uint32_t np2;        // bitonic size for the shared-memory path
    MergeParams mp;
    XPushParams push;    // push.tab == nullptr: nothing to push
};

static constexpr uint32_t SCAN_INLINE_Q_MAX = 3328;  // bytes of one padded query row carried in the parameters (f32[768] = 3072)
struct ScanParams {
    const uint8_t* vectors;  // slab rows, row_stride bytes apart, zero padded
    const uint8_t* skip;     // per-row flags (non-zero = skipped by scans) or nullptr
    const uint8_t* queries;  // nq_total rows of row_stride bytes, zero padded (device)
    uint64_t* out_keys;      // TOPK: [nq_total][gridDim.x][k] ; EMIT: [nq_total][n_rows]
    uint64_t n_rows;
    uint32_t nq_total;
    uint32_t k;
    uint32_t row_stride;       // bytes, multiple of 16
    uint32_t chunk_bytes;      // per-row mode: multiple of 64; contig: == row_stride
    uint32_t n_chunks;
    uint32_t smem_row_stride;  // bytes between rows of a stage in shared memory
    uint32_t rows_per_stage;   // RS = RPW * m   (m == 1 when n_chunks > 1); one consumer warp per stage
    uint32_t n_stages;         // D: ring depth per consumer warp (C*D stages in total)
    uint32_t contig;           // 1: a stage is one contiguous bulk copy (smem_row_stride == row_stride)
    uint32_t n_consumers;      // C warps per CTA (blockDim = 32*C)
    uint32_t qc_kind;          // 0 f32 sum of squares, 1 int8
    uint32_t list_stride;      // keys reserved per query for the C per-warp lists: pow2 >= C*k
    ScanTail tail;
    unsigned long long* dbg;   // optional [gridDim.x][4] globaltimer stamps: CTA start, pipeline primed, rows done, CTA end (tools/scan_timeline.py)
    // A single host query travels INSIDE the launch (kernel parameter space) instead of through a host->device copy that the
    // kernel would have to wait for: one stream operation less per query (~6 us of a 600 us sharded query).
    uint32_t inline_q_bytes;   // 0: queries are at `queries` (device memory)
    alignas(16) uint8_t inline_q[SCAN_INLINE_Q_MAX];
};
static_assert(sizeof(ScanParams) <= 4096, "kernel parameters are limited to 4 KB");

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// physical unit read by a lane for logical unit u (see "rotating" above)
template <class T>
__device__ __forceinline__ uint32_t phys_unit(uint32_t u, uint32_t units8, uint32_t rot) {
    if (!T::ORDER_FREE) return u;
    if (T::LPR == 4) return u < units8 ? (u ^ rot) : u;
    return u < units8 ? ((u & ~7u) | ((u + rot) & 7u)) : u;
}

template <class T, int QB, bool EMIT>
__global__ void __launch_bounds__(512, 1) scan_kernel(const __grid_constant__ ScanParams p) {
    constexpr int LPR = T::LPR;
    constexpr int RPW = 32 / LPR;  // rows a warp scores at once
    extern __shared__ __align__(128) uint8_t smem[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t C = p.n_consumers, D = p.n_stages, S = C * D, RS = p.rows_per_stage;
    const uint32_t stage_bytes = RS * p.smem_row_stride;
    // layout: [stages][queries QB*row_stride][qc QB floats, padded to 64][hdr C*QB][lists C*QB*k][barriers 2*S]
    uint8_t* s_stage = smem;
    uint8_t* s_query = s_stage + (size_t)S * stage_bytes;
    float* s_qc = (float*)(s_query + (size_t)QB * p.row_stride);
    ListHdr* s_hdr = (ListHdr*)((uint8_t*)s_qc + 64);
    uint64_t* s_list = (uint64_t*)(s_hdr + C * QB);
    uint64_t* s_bar = s_list + (EMIT ? 0 : (size_t)QB * p.list_stride);  // lists are query-major: [QB][list_stride]
    const uint32_t bar_full = smem_u32(s_bar);
    uint32_t* s_tile = (uint32_t*)(s_bar + S);  // [S] tile held by each ring slot (dynamic tile scheduling)

    const uint32_t q0 = blockIdx.y * QB;  // first query of this pass
    const uint32_t nq_here = min((uint32_t)QB, p.nq_total - q0);
    if (p.dbg && threadIdx.x == 0 && blockIdx.y == 0) p.dbg[blockIdx.x * 4 + 0] = globaltimer_ns();

    // Dynamic tile scheduling state (see the comment at `dyn` below): the CTA draws BLOCKS of 16 consecutive tiles from the
    // global counter (one global atomic per 16 tiles: per-tile global atomics on one address saturate at ~350 M/s and cost
    // 14 % of the bandwidth), its warps take single tiles out of the current block with a shared-memory ticket.
    __shared__ uint32_t s_dynL, s_dynProd, s_dynDone, s_dynBase[8];
    const bool dyn = !EMIT && p.tail.dyn != nullptr;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) mbar_init(bar_full + 8 * s, 1);
        mbar_fence_init();
        if (dyn) {
            const uint32_t r = atomicAdd(p.tail.dyn + blockIdx.y, 32u);  // blocks 0 and 1 of this CTA
            s_dynBase[0] = r;
            s_dynBase[1] = r + 16;
            s_dynProd = 2;
            s_dynL = 0;
            s_dynDone = 0;
        }
    }
    // stage queries (zero-fill the slots past nq_here so the arithmetic stays finite)
    {
        const uint32_t qunits = p.row_stride / 16;
        const uint4* gq = p.inline_q_bytes ? (const uint4*)p.inline_q : (const uint4*)(p.queries + (size_t)q0 * p.row_stride);
        uint4* sq = (uint4*)s_query;
        for (uint32_t i = threadIdx.x; i < QB * qunits; i += blockDim.x)
            sq[i] = (i / qunits) < nq_here ? gq[i] : make_uint4(0, 0, 0, 0);
        if (!EMIT)
            for (uint32_t i = threadIdx.x; i < C * QB; i += blockDim.x) {
                s_hdr[i].tau = KEY_NONE;
                s_hdr[i].maxpos = 0;
                s_hdr[i].cnt = 0;
            }
    }
    __syncthreads();
    if (T::HAS_QC) {
        // warp w computes |q|^2 for queries w, w+nwarps, ... with its lanes 0..3
        const uint32_t qunits = p.row_stride / 16;
        for (uint32_t i = warp; i < QB; i += (blockDim.x >> 5)) {
            float v = query_const((const uint4*)(s_query + (size_t)i * p.row_stride), qunits, lane & 3, p.qc_kind);
            if (lane == 0) s_qc[i] = v;
        }
        __syncthreads();
    }

    const uint64_t n_tiles = (p.n_rows + RS - 1) / RS;

    {
        // ================= every warp: consumer + producer of its own ring =================
        const int g = lane % LPR;
        const int rl = lane / LPR;
        const uint32_t m_steps = RS / RPW;
        const uint32_t rot = !p.contig ? 0u : (LPR == 4 ? 4u * (rl & 1) : (uint32_t)(lane & 7));
        uint64_t* my_list = EMIT ? nullptr : s_list + (size_t)warp * p.k;  // + i*list_stride for query i
        ListHdr* my_hdr = s_hdr + warp * QB;
        const uint32_t q_base = smem_u32(s_query);
        const uint32_t my_bar = bar_full + 8 * warp * D;
        uint8_t* my_stage = s_stage + (size_t)warp * D * stage_bytes;
        // number of tiles of this warp: tiles blockIdx.x + (warp + i*C)*gridDim.x < n_tiles
        const uint64_t first_tile = blockIdx.x + (uint64_t)warp * gridDim.x;
        const uint64_t tile_step = (uint64_t)C * gridDim.x;
        const uint32_t my_tiles = first_tile < n_tiles ? (uint32_t)((n_tiles - first_tile + tile_step - 1) / tile_step) : 0u;
        const uint32_t my_iters = my_tiles * p.n_chunks;

        // Dynamic tile scheduling (single-chunk rows): instead of the fixed round robin, a warp takes the NEXT tile of the
        // CTA's current 16-tile block whenever it refills a ring slot, and the CTA draws the next block from a global
        // counter.  Tiles are still handed out in address order (the CTAs keep reading neighbouring memory), but a CTA that
        // streams a little slower simply takes fewer blocks: with the static split the first CTA was done after 508 us
        // and the last after 551 us on a 3.84 GB shard (tools/scan_timeline.py) — 8 % of the launch spent waiting for
        // stragglers.
        uint32_t* my_tile_slots = s_tile + warp * D;
        // ring depth in use: the prologue must not let the first warps hoard a small table (each warp starts with its fair
        // share of tiles at most), so a table with fewer tiles than D per warp runs with a shallower ring
        const uint32_t Dr = dyn ? (uint32_t)max((uint64_t)1, min((uint64_t)D, (n_tiles + (uint64_t)C * gridDim.x - 1) / ((uint64_t)C * gridDim.x))) : D;
        bool reserve_next = false;
        // Two phases: the first `sr` tiles of every warp come from the static round robin (tile -> SM affinity as before: a
        // fully dynamic hand-out, per tile or per 16-tile block, streamed 13 % SLOWER in steady state — measured three ways,
        // profiles/README.md), the remaining ~10 % of the slab is handed out dynamically and absorbs the stragglers.
        const uint32_t sr = dyn ? p.tail.static_rounds : 0u;
        const uint32_t n_static = sr * C * gridDim.x;
        uint32_t my_ticket = 0;  // lane 0: this warp's shared-memory ticket for its NEXT dynamic refill (drawn one refill ahead)
        if (dyn && sr == 0 && lane == 0) my_ticket = atomicAdd(&s_dynL, 1u);
        // ticket -> tile: block = ticket / 16 (its base was published by warp 0 at least one block ago), offset = ticket % 16
        auto resolve_ticket = [&]() -> uint32_t {
            const uint32_t b = my_ticket >> 4, o = my_ticket & 15u;
            while (*(volatile uint32_t*)&s_dynProd <= b) {
                if (*(volatile uint32_t*)&s_dynDone) return 0xFFFFFFFFu;  // the slab is exhausted: later blocks are never published
                __nanosleep(32);
            }
            const uint32_t base = *(volatile uint32_t*)&s_dynBase[b & 7u];
            return base >= 0xFFFFFFF0u ? 0xFFFFFFFFu : base + o;
        };
        // warp 0 keeps the block ring one block ahead of the CTA's consumption
        auto top_up = [&]() {
            uint32_t prod = *(volatile uint32_t*)&s_dynProd;
            const uint32_t cons = (*(volatile uint32_t*)&s_dynL) >> 4;
            while ((int32_t)(prod - cons) < 2) {
                const uint32_t r = atomicAdd(p.tail.dyn + blockIdx.y, 16u);
                *(volatile uint32_t*)&s_dynBase[prod & 7u] = r;
                __threadfence_block();
                *(volatile uint32_t*)&s_dynProd = ++prod;
                if ((uint64_t)r + n_static >= n_tiles) {
                    *(volatile uint32_t*)&s_dynDone = 1u;
                    break;
                }
            }
        };
        // issue the copies of ring iteration `lit` (tile lit / n_chunks, chunk lit % n_chunks) into slot lit % D
        auto issue = [&](uint32_t lit) {
            uint32_t jl, c;
            uint64_t row0;
            if (dyn) {
                uint32_t t = 0xFFFFFFFFu;
                if (lane == 0) {
                    if (lit < sr) {
                        t = (uint32_t)(first_tile + (uint64_t)lit * tile_step);  // static phase: the fixed round robin
                    } else {
                        t = resolve_ticket();
                        if (t != 0xFFFFFFFFu) t += n_static;
                        if ((uint64_t)t >= n_tiles) {
                            t = 0xFFFFFFFFu;
                            if (warp == 0) *(volatile uint32_t*)&s_dynDone = 1u;  // warp 0 stops topping up: nobody may wait for it
                        }
                    }
                    my_tile_slots[lit % Dr] = t;
                }
                t = __shfl_sync(0xffffffffu, t, 0);
                __syncwarp();
                if (t == 0xFFFFFFFFu) return;
                reserve_next = lit + 1 >= sr;  // the next refill is a dynamic one: draw its ticket now
                jl = lit;
                c = 0;
                row0 = (uint64_t)t * RS;
            } else {
                if (lit >= my_iters) return;
                jl = lit / p.n_chunks;
                c = lit - jl * p.n_chunks;
                row0 = (first_tile + (uint64_t)jl * tile_step) * RS;
            }
            const uint32_t valid = (uint32_t)min((uint64_t)RS, p.n_rows - row0);
            const uint32_t slot = lit % Dr;
            const uint32_t bar = my_bar + 8 * slot;
            const uint32_t dst0 = smem_u32(my_stage + (size_t)slot * stage_bytes);
            fence_proxy_async();  // the stage was read through the generic proxy; the copy writes through the async proxy
            if (p.contig) {
                if (lane == 0) {
                    const uint32_t bytes = valid * p.row_stride;  // <= ~200 KB: fits tx-count and copy size
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(dst0, p.vectors + row0 * p.row_stride, bytes, bar);
                }
            } else {
                const uint32_t off = c * p.chunk_bytes;
                const uint32_t len = min(p.chunk_bytes, p.row_stride - off);
                if (lane == 0) mbar_expect_tx(bar, valid * len);
                __syncwarp();
                for (uint32_t r = lane; r < valid; r += 32)
                    bulk_g2s(dst0 + r * p.smem_row_stride, p.vectors + (row0 + r) * p.row_stride + off, len, bar);
            }
            if (reserve_next) {
                if (lane == 0) {
                    my_ticket = atomicAdd(&s_dynL, 1u);  // consumed by the NEXT refill: the latency overlaps a tile of work
                    if (warp == 0) top_up();
                }
                reserve_next = false;
            }
            __syncwarp();
        };
        for (uint32_t lit = 0; lit < Dr; ++lit) issue(lit);  // prologue: fill the ring
        if (p.dbg && threadIdx.x == 0 && blockIdx.y == 0) p.dbg[blockIdx.x * 4 + 1] = globaltimer_ns();

        // thread-per-row integer path (Hamming): byte offsets of the 8 units of a 128-byte segment in this
        // lane's rotated order, and the query in registers when a row is exactly one segment (bit[1024])
        uint32_t offs[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) offs[t] = ((uint32_t)(t + rot) & 7u) * 16u;
        const bool qreg_ok = (LPR == 1) && T::ORDER_FREE && (QB == 1) && p.contig && p.row_stride == 128;
        uint4 qreg[8];
        if (qreg_ok) {
#pragma unroll
            for (int t = 0; t < 8; ++t) qreg[t] = lds128(q_base + offs[t]);
        }

        for (uint32_t jl = 0; dyn || jl < my_tiles; ++jl) {
            uint64_t row0 = (first_tile + (uint64_t)jl * tile_step) * RS;
            if (dyn) {
                const uint32_t t = my_tile_slots[jl % Dr];  // written by lane 0 in issue(), __syncwarp()ed there
                if (t == 0xFFFFFFFFu) break;               // tiles are handed out in order: nothing is left for this warp
                row0 = (uint64_t)t * RS;
            }
            const uint32_t it0 = jl * p.n_chunks;
            for (uint32_t ms = 0; ms < m_steps; ++ms) {
                const uint32_t r_in_stage = ms * RPW + rl;
                typename T::Acc acc;
                T::init(acc);
                for (uint32_t c = 0; c < p.n_chunks; ++c) {
                    const uint32_t lit = it0 + c;
                    const uint32_t s = lit % Dr, ph = (lit / Dr) & 1;
                    if (ms == 0) mbar_wait(my_bar + 8 * s, ph);
                    const uint32_t off = p.contig ? 0 : c * p.chunk_bytes;
                    const uint32_t len = p.contig ? p.row_stride : min(p.chunk_bytes, p.row_stride - off);
                    const uint32_t units = len / 16, units8 = units & ~7u;
                    const uint32_t xb = smem_u32(my_stage + (size_t)s * stage_bytes) + r_in_stage * p.smem_row_stride;
                    const uint32_t qb = q_base + off;
                    uint32_t u = g;
                    if constexpr (LPR == 1 && T::ORDER_FREE) {
                        const uint32_t nseg = units >> 3;
                        for (uint32_t sg = 0; sg < nseg; ++sg) {
                            const uint32_t xs = xb + sg * 128, qs = qb + sg * 128;
                            uint4 xv[8];
#pragma unroll
                            for (int t = 0; t < 8; ++t) xv[t] = lds128(xs + offs[t]);
                            if (qreg_ok) {
#pragma unroll
                                for (int t = 0; t < 8; ++t) {
                                    uint4 qv[QB];
                                    qv[0] = qreg[t];
#pragma unroll
                                    for (int i = 1; i < QB; ++i) qv[i] = qreg[t];
                                    T::step(acc, xv[t], qv);
                                }
                            } else {
#pragma unroll
                                for (int t = 0; t < 8; ++t) {
                                    uint4 qv[QB];
#pragma unroll
                                    for (int i = 0; i < QB; ++i) qv[i] = lds128(qs + i * p.row_stride + offs[t]);
                                    T::step(acc, xv[t], qv);
                                }
                            }
                        }
                        u = nseg * 8;  // the (< 8) tail units below are read unrotated
                    }
                    if constexpr (LPR == 4 && T::ORDER_FREE) {
                        // predicate-free walk over full 128-byte segments: this lane's two units of a segment sit at
                        // oA / oB (swapped for odd rows, which is what keeps the 4-lane groups off each other's banks)
                        const uint32_t nseg = units >> 3;
                        const uint32_t oA = (uint32_t)(g + rot) * 16u, oB = (uint32_t)(g + 4 - rot) * 16u;
                        uint32_t sg = 0;
                        for (; sg + 2 <= nseg; sg += 2) {
                            const uint32_t xs = xb + sg * 128, qs = qb + sg * 128;
                            uint4 x0 = lds128(xs + oA), x1 = lds128(xs + oB), x2 = lds128(xs + 128 + oA),
                                  x3 = lds128(xs + 128 + oB);
                            uint4 q0v[QB], q1v[QB], q2v[QB], q3v[QB];
#pragma unroll
                            for (int i = 0; i < QB; ++i) {
                                q0v[i] = lds128(qs + i * p.row_stride + oA);
                                q1v[i] = lds128(qs + i * p.row_stride + oB);
                                q2v[i] = lds128(qs + i * p.row_stride + 128 + oA);
                                q3v[i] = lds128(qs + i * p.row_stride + 128 + oB);
                            }
                            T::step(acc, x0, q0v);
                            T::step(acc, x1, q1v);
                            T::step(acc, x2, q2v);
                            T::step(acc, x3, q3v);
                        }
                        if (sg < nseg) {
                            const uint32_t xs = xb + sg * 128, qs = qb + sg * 128;
                            uint4 x0 = lds128(xs + oA), x1 = lds128(xs + oB);
                            uint4 q0v[QB], q1v[QB];
#pragma unroll
                            for (int i = 0; i < QB; ++i) {
                                q0v[i] = lds128(qs + i * p.row_stride + oA);
                                q1v[i] = lds128(qs + i * p.row_stride + oB);
                            }
                            T::step(acc, x0, q0v);
                            T::step(acc, x1, q1v);
                        }
                        u = g + nseg * 8;  // the (< 8) tail units below are read unrotated
                    }
                    // 4 units in flight per lane
                    for (; u + 3 * LPR < units; u += 4 * LPR) {
                        const uint32_t o0 = phys_unit<T>(u, units8, rot) * 16, o1 = phys_unit<T>(u + LPR, units8, rot) * 16,
                                       o2 = phys_unit<T>(u + 2 * LPR, units8, rot) * 16,
                                       o3 = phys_unit<T>(u + 3 * LPR, units8, rot) * 16;
                        uint4 x0 = lds128(xb + o0), x1 = lds128(xb + o1), x2 = lds128(xb + o2), x3 = lds128(xb + o3);
                        if constexpr (QB <= 2) {
                            // all eight (x, q) loads in flight before the first FMA
                            uint4 qa[QB], qb2[QB], qc2[QB], qd[QB];
#pragma unroll
                            for (int i = 0; i < QB; ++i) {
                                qa[i] = lds128(qb + i * p.row_stride + o0);
                                qb2[i] = lds128(qb + i * p.row_stride + o1);
                                qc2[i] = lds128(qb + i * p.row_stride + o2);
                                qd[i] = lds128(qb + i * p.row_stride + o3);
                            }
                            T::step(acc, x0, qa);
                            T::step(acc, x1, qb2);
                            T::step(acc, x2, qc2);
                            T::step(acc, x3, qd);
                        } else {
                            uint4 qv[QB];
#pragma unroll
                            for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.row_stride + o0);
                            T::step(acc, x0, qv);
#pragma unroll
                            for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.row_stride + o1);
                            T::step(acc, x1, qv);
#pragma unroll
                            for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.row_stride + o2);
                            T::step(acc, x2, qv);
#pragma unroll
                            for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.row_stride + o3);
                            T::step(acc, x3, qv);
                        }
                    }
                    for (; u < units; u += LPR) {
                        const uint32_t o = phys_unit<T>(u, units8, rot) * 16;
                        uint4 x0 = lds128(xb + o);
                        uint4 qv[QB];
#pragma unroll
                        for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.row_stride + o);
                        T::step(acc, x0, qv);
                    }
                    if (ms == m_steps - 1) {
                        __syncwarp();     // every lane has consumed its loads of this stage
                        issue(lit + Dr);  // refill it with the tile Dr steps ahead
                    }
                }
                // ---- distance -> key -> fused top-k (or emit) ----
                const uint64_t row = row0 + r_in_stage;
                bool live = row < p.n_rows;
                if (live && p.skip != nullptr && g == 0) live = p.skip[row] == 0;
#pragma unroll
                for (int i = 0; i < QB; ++i) {
                    if constexpr (EMIT) {
                        float d = T::finish(acc, i, s_qc);
                        if (g == 0 && row < p.n_rows && (uint32_t)i < nq_here)
                            p.out_keys[(size_t)(q0 + i) * p.n_rows + row] = live ? make_key(d, (uint32_t)row) : KEY_NONE;
                    } else if constexpr (T::HAS_BOUND) {
                        // cheap f32 lower bound first; the f64 finish only when some row of the warp may enter the list
                        const typename T::Red red = T::reduce(acc, i);
                        const float tau_d = order_bits_inv((uint32_t)(my_hdr[i].tau >> 32));  // NaN while the list fills
                        const bool cand = live && g == 0 && !(T::bound(red, s_qc[i]) > tau_d);
                        if (__any_sync(0xffffffffu, cand)) {
                            const float d = T::exact(red, s_qc[i]);
                            list_offer(my_list + (size_t)i * p.list_stride, my_hdr + i, p.k, make_key(d, (uint32_t)row), cand, lane);
                        }
                    } else {
                        const float d = T::finish(acc, i, s_qc);
                        list_offer(my_list + (size_t)i * p.list_stride, my_hdr + i, p.k, make_key(d, (uint32_t)row), live && g == 0, lane);
                    }
                }
            }
        }
    }
    if (EMIT) return;
    __syncthreads();
    if (p.dbg && threadIdx.x == 0 && blockIdx.y == 0) p.dbg[blockIdx.x * 4 + 2] = globaltimer_ns();
    // ---- CTA merge: the C per-warp lists of a query are contiguous; blank the unused slots, bitonic-sort the
    //      list_stride keys with the whole CTA and emit the k smallest as this CTA's partial result ----
    if (p.list_stride <= 256) {  // the usual case (C * k <= 256): one warp per query sorts that query's lists in registers,
                                 // the QB queries of a multi-query pass in parallel instead of QB block-wide sorts in a row
        for (uint32_t i = (uint32_t)warp; i < nq_here; i += blockDim.x >> 5) {
            const uint64_t* base = s_list + (size_t)i * p.list_stride;
            uint64_t v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t j = (uint32_t)lane * 8 + r;
                uint64_t key = KEY_NONE;
                if (j < p.list_stride) {
                    const uint32_t w = j / p.k, e = j - w * p.k;
                    if (w < C && e < s_hdr[w * QB + i].cnt) key = base[j];
                }
                v[r] = key;
            }
            warp_sort256_shared(v, lane);
            uint64_t* out = p.out_keys + ((size_t)(q0 + i) * gridDim.x + blockIdx.x) * p.k;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t j = (uint32_t)lane * 8 + r;
                if (j < p.k) out[j] = v[r];
            }
        }
    } else {
        for (uint32_t i = 0; i < nq_here; ++i) {
            uint64_t* base = s_list + (size_t)i * p.list_stride;
            for (uint32_t j = threadIdx.x; j < p.list_stride; j += blockDim.x) {
                const uint32_t w = j / p.k, e = j - w * p.k;
                if (w >= C || e >= s_hdr[w * QB + i].cnt) base[j] = KEY_NONE;
            }
            __syncthreads();
            block_bitonic_sort(base, p.list_stride);
            uint64_t* out = p.out_keys + ((size_t)(q0 + i) * gridDim.x + blockIdx.x) * p.k;
            for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = base[j];
        }
    }
    // ---- fused tail: the last CTA of this query pass merges the gridDim.x partial lists (and pushes to the peers) ----
    if (p.dbg && blockIdx.y == 0) {
        __syncthreads();
        if (threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 3] = globaltimer_ns();
    }
    if (p.tail.counter == nullptr) return;
    __shared__ uint32_t s_last;
    __threadfence();  // this CTA's partial lists are visible device-wide before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(p.tail.counter + blockIdx.y, 1u);
        s_last = prev + 1 == gridDim.x ? 1u : 0u;
        if (s_last) {  // re-arm for the next launch
            p.tail.counter[blockIdx.y] = 0;
            if (p.tail.dyn) p.tail.dyn[blockIdx.y] = 0;
        }
    }
    __syncthreads();
    if (!s_last || !p.tail.do_merge) return;
    __threadfence();
    uint64_t* scratch = (uint64_t*)s_stage;  // every bulk copy into the stages has been consumed
    for (uint32_t i = 0; i < nq_here; ++i) final_merge_cta(p.tail.mp, q0 + i, scratch, p.tail.np2, true);
    if (p.tail.push.tab != nullptr) {
        for (uint32_t i = 0; i < nq_here; ++i) xpush_query(p.tail.push, q0 + i);
        xpush_publish(p.tail.push, gridDim.y);
    }
    if (p.dbg && threadIdx.x == 0 && blockIdx.y == 0) p.dbg[gridDim.x * 4] = globaltimer_ns();  // end of the fused merge
}

// ---------------------------------------------------------------------------
// K4 batched: Hamming for many queries over short bit rows (row_stride <= 128 bytes), lane = query.
// The multi-query scan (QB = 8) re-reads the query words from shared memory for every row (5.4 instructions per
// XOR/POPC pair, 9 k instructions of unrolled code, 8 queries per pass over the data).  Here a warp owns 32 queries:
// each lane keeps ITS query in registers (<= 32 words), every row word is one broadcast LDS for the whole warp, and
// each lane ranks the rows for its own query in a private k-entry list (unsorted, tracked maximum; entry e of lane l
// at list[e * 32 + l]: conflict-free).  Rows reach shared memory through the same per-warp self-refilled bulk-copy
// rings as scan_kernel.  Exact: integer popcounts, keys (order_bits((float)d) << 32 | row) as everywhere else.
// ---------------------------------------------------------------------------
struct HamBatchParams {
    const uint8_t* vectors;
    const uint8_t* skip;
    const uint8_t* queries;   // [nq][row_stride], zero padded
    uint64_t* out_keys;       // [nq][gridDim.x][k]
    uint64_t n_rows;
    uint32_t nq, k;
    uint32_t row_stride;      // multiple of 16, <= 128
    uint32_t rows_per_tile;   // tile = rows_per_tile * row_stride bytes, one bulk copy
    uint32_t n_stages;        // ring depth per warp
    uint32_t n_warps;         // C
};

template <int W>  // 32-bit words per row
__global__ void __launch_bounds__(512) ham_batch_kernel(const HamBatchParams p) {
    extern __shared__ __align__(128) uint8_t hb_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t C = p.n_warps, D = p.n_stages, tile_bytes = p.rows_per_tile * p.row_stride;
    uint8_t* s_tiles = hb_smem;                                                  // [C][D][tile_bytes]
    uint64_t* s_list = (uint64_t*)(hb_smem + (size_t)C * D * tile_bytes);        // [C][k][32]
    uint64_t* s_bar = s_list + (size_t)C * p.k * 32;                             // [C][D]
    const uint32_t my_bar = smem_u32(s_bar + (size_t)warp * D);
    uint8_t* my_tiles = s_tiles + (size_t)warp * D * tile_bytes;
    uint64_t* my_list = s_list + (size_t)warp * p.k * 32;
    if (lane == 0)
        for (uint32_t d = 0; d < D; ++d) mbar_init(my_bar + 8 * d, 1);
    mbar_fence_init();
    __syncwarp();

    const uint32_t q = blockIdx.y * 32 + lane;
    const bool q_ok = q < p.nq;
    uint32_t qw[W];
    {
        const uint4* qp = (const uint4*)(p.queries + (size_t)(q_ok ? q : 0) * p.row_stride);
#pragma unroll
        for (int w4 = 0; w4 < W / 4; ++w4) {
            const uint4 v = q_ok ? qp[w4] : make_uint4(0, 0, 0, 0);
            qw[4 * w4 + 0] = v.x; qw[4 * w4 + 1] = v.y; qw[4 * w4 + 2] = v.z; qw[4 * w4 + 3] = v.w;
        }
    }
    // tiles of this warp: tile t -> CTA t % gridDim.x, the j-th tile of a CTA -> warp j % C
    const uint64_t n_tiles = (p.n_rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const uint64_t first = blockIdx.x + (uint64_t)warp * gridDim.x, step = (uint64_t)C * gridDim.x;
    const uint64_t my_n = first < n_tiles ? (n_tiles - first + step - 1) / step : 0;
    auto issue = [&](uint64_t j) {  // lane 0: bulk copy of my j-th tile into ring slot j % D
        const uint64_t row0 = (first + j * step) * p.rows_per_tile;
        const uint32_t rows = (uint32_t)min((uint64_t)p.rows_per_tile, p.n_rows - row0);
        const uint32_t bytes = rows * p.row_stride, slot = (uint32_t)(j % D);
        mbar_expect_tx(my_bar + 8 * slot, bytes);
        bulk_g2s(smem_u32(my_tiles + (size_t)slot * tile_bytes), p.vectors + row0 * p.row_stride, bytes, my_bar + 8 * slot);
    };
    if (lane == 0)
        for (uint64_t j = 0; j < min((uint64_t)D, my_n); ++j) issue(j);

    uint32_t cnt = 0, maxpos = 0;
    uint64_t tau_key = q_ok ? KEY_NONE : 0ull;   // a padding lane never admits anything
    for (uint64_t j = 0; j < my_n; ++j) {
        const uint32_t slot = (uint32_t)(j % D);
        mbar_wait(my_bar + 8 * slot, (uint32_t)((j / D) & 1));
        const uint64_t row0 = (first + j * step) * p.rows_per_tile;
        const uint32_t rows = (uint32_t)min((uint64_t)p.rows_per_tile, p.n_rows - row0);
        const uint32_t tb = smem_u32(my_tiles + (size_t)slot * tile_bytes);
#pragma unroll 2
        for (uint32_t r = 0; r < rows; ++r) {
            // (a Harley-Seal carry-save tree that trades 32 POPC for 60 LOP3 + 6 POPC was measured: 10.3 vs 9.8 ms, not kept)
            int d0 = 0, d1 = 0;
#pragma unroll
            for (int w4 = 0; w4 < W / 4; ++w4) {
                const uint4 x = lds128(tb + r * p.row_stride + w4 * 16);  // same address in every lane: broadcast
                d0 += __popc(x.x ^ qw[4 * w4 + 0]) + __popc(x.y ^ qw[4 * w4 + 1]);
                d1 += __popc(x.z ^ qw[4 * w4 + 2]) + __popc(x.w ^ qw[4 * w4 + 3]);
            }
            const uint64_t key = make_key((float)(d0 + d1), (uint32_t)(row0 + r));
            if (key < tau_key && !(p.skip && p.skip[row0 + r])) {   // rare after the first few hundred rows
                if (cnt < p.k) {
                    my_list[cnt * 32 + lane] = key;
                    ++cnt;
                } else {
                    my_list[maxpos * 32 + lane] = key;
                }
                if (cnt == p.k) {  // full: the bound is the largest kept key
                    uint64_t m = my_list[lane];
                    uint32_t mp = 0;
                    for (uint32_t e = 1; e < p.k; ++e) {
                        const uint64_t v = my_list[e * 32 + lane];
                        if (v > m) {
                            m = v;
                            mp = e;
                        }
                    }
                    tau_key = m;
                    maxpos = mp;
                }
            }
        }
        __syncwarp();         // every lane is done with the slot before it is refilled
        fence_proxy_async();  // the slot was read through the generic proxy; the copy writes through the async proxy
        if (lane == 0 && j + D < my_n) issue(j + D);
    }
    // blank the unused entries, then one warp per query merges the C lists (C * k <= 256 keys) in registers
    for (uint32_t e = cnt; e < p.k; ++e) my_list[e * 32 + lane] = KEY_NONE;
    __syncthreads();
    for (uint32_t ql = (uint32_t)warp; ql < 32; ql += C) {
        const uint32_t qq = blockIdx.y * 32 + ql;
        if (qq >= p.nq) continue;
        uint64_t v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t i = (uint32_t)lane * 8 + r;   // i -> (warp i / k, entry i % k)
            v[r] = i < C * p.k ? s_list[((size_t)(i / p.k) * p.k + i % p.k) * 32 + ql] : KEY_NONE;
        }
        warp_sort256(v, lane);
        uint64_t* out = p.out_keys + ((size_t)qq * gridDim.x + blockIdx.x) * p.k;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t i = (uint32_t)lane * 8 + r;
            if (i < p.k) out[i] = v[r];
        }
    }
}

// ---------------------------------------------------------------------------
// K6: final selection + sort + decode.  One CTA per query reads `n_cand` keys
// (per-CTA partial lists of the scan), keeps the k smallest, sorts them
// ascending == (distance_f32, rowid) order, translates positions to rowids.
// ---------------------------------------------------------------------------
// small candidate sets (n_cand <= 16384, e.g. 148 lists x k<=110): load everything, bitonic-sort, keep the first k.
__global__ void __launch_bounds__(1024) merge_sort_kernel(const MergeParams p, uint32_t np2) {
    extern __shared__ __align__(128) uint8_t smem[];
    final_merge_cta(p, blockIdx.x, (uint64_t*)smem, np2, false);
}

// Small-table fast path of the final merge (k <= 32, at most 2048 candidates per query: e.g. 148 CTAs x k = 10):
// eight warps each sort 256 candidates in registers and keep their 32 smallest, warp 0 sorts those 256 and emits the
// first k.  Two register sorts and one __syncthreads instead of 66 block-wide bitonic stages (22 us -> a few us, which
// is most of a single query's latency on a 10 k-row table).
__global__ void __launch_bounds__(256) merge_small_kernel(const MergeParams p) {
    __shared__ uint64_t part[256];
    final_merge_cta(p, blockIdx.x, part, 0, true);
}

// large candidate sets: after a segmented radix sort of [nq][n_cand] keys, decode the first k of each segment
__global__ void decode_segments_kernel(const uint64_t* keys, uint64_t n_cand, uint32_t k, const int64_t* rowids,
                                       int64_t first_rowid, int64_t pad_rowid, int64_t* out_rowids, float* out_dists,
                                       uint32_t* out_counts) {
    const uint64_t* seg = keys + (size_t)blockIdx.x * n_cand;
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const uint64_t key = j < n_cand ? seg[j] : KEY_NONE;
        const size_t o = (size_t)blockIdx.x * k + j;
        if (key == KEY_NONE) {
            out_rowids[o] = pad_rowid;
            out_dists[o] = __int_as_float(0x7F800000);
        } else {
            const uint32_t pos = (uint32_t)key;
            out_rowids[o] = rowids ? rowids[pos] : first_rowid + (int64_t)pos;
            out_dists[o] = order_bits_inv((uint32_t)(key >> 32));
            ++cnt;
        }
    }
    if (cnt) atomicAdd(&total, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && out_counts) out_counts[blockIdx.x] = total;
}

// large-k path (k > fused limit): after a full radix sort of the emitted keys
// the first k keys of each query are decoded.
__global__ void decode_sorted_kernel(const uint64_t* keys, uint64_t n_rows, uint32_t k, const int64_t* rowids,
                                     int64_t first_rowid, int64_t pad_rowid, int64_t* out_rowids, float* out_dists,
                                     uint32_t* out_count) {
    uint32_t cnt = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
        const uint64_t key = j < n_rows ? keys[j] : KEY_NONE;
        if (key == KEY_NONE) {
            out_rowids[j] = pad_rowid;
            out_dists[j] = __int_as_float(0x7F800000);
        } else {
            const uint32_t pos = (uint32_t)key;
            out_rowids[j] = rowids ? rowids[pos] : first_rowid + (int64_t)pos;
            out_dists[j] = order_bits_inv((uint32_t)(key >> 32));
            ++cnt;
        }
    }
    if (out_count) atomicAdd(out_count, cnt);
}

// cross-shard merge input: (dist, rowid) records from G lists -> keys are not
// usable (positions are shard-local), so rank on (order_bits(dist), rowid).
struct XMergeParams {
    const int64_t* rowids;  // [nlists][nq][k]
    const float* dists;     // [nlists][nq][k]
    uint32_t nlists, nq, k, kp2;
    int64_t* out_rowids;  // [nq][k]
    float* out_dists;
};
__global__ void __launch_bounds__(256) xmerge_kernel(const XMergeParams p) {
    // candidates per query: nlists*k (small: G<=8 shards).  Sort all of them
    // with a bitonic network on (order_bits, rowid) and keep the first k.
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* kb = (uint32_t*)smem;          // [np2]
    int64_t* rid = (int64_t*)(kb + p.kp2);   // [np2]  (kp2 here = pow2 >= nlists*k, 8-byte aligned by construction)
    const uint32_t q = blockIdx.x, n = p.nlists * p.k, np2 = p.kp2;
    for (uint32_t j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < n) {
            const uint32_t l = j / p.k, i = j % p.k;
            const size_t o = ((size_t)l * p.nq + q) * p.k + i;
            int64_t r = p.rowids[o];
            kb[j] = (r == INT64_MAX) ? 0xFFFFFFFFu : order_bits(p.dists[o]);
            rid[j] = r;
        } else {
            kb[j] = 0xFFFFFFFFu;
            rid[j] = INT64_MAX;
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= np2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2 / 2; t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool up = (lo & size) == 0;
                uint32_t ka = kb[lo], kc = kb[hi];
                int64_t ra = rid[lo], rc = rid[hi];
                bool gt = ka > kc || (ka == kc && ra > rc);
                if (gt == up) {
                    kb[lo] = kc;
                    kb[hi] = ka;
                    rid[lo] = rc;
                    rid[hi] = ra;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
        const size_t o = (size_t)q * p.k + j;
        p.out_rowids[o] = rid[j];
        p.out_dists[o] = rid[j] == INT64_MAX ? __int_as_float(0x7F800000) : order_bits_inv(kb[j]);
    }
}

// ---------------------------------------------------------------------------
// K5: gathered pair scoring.  One LPR-lane group per (a, b) pair, operands read
// straight from global memory with 128-bit loads.  Used by vecgpu_score (a =
// query of the candidate, b = slab row found by rowid) and
// vecgpu_distance_pairs (a[i], b[i]).
// ---------------------------------------------------------------------------
struct PairParams {
    const uint8_t* a_base;
    const uint8_t* b_base;
    uint32_t a_stride, b_stride;  // bytes between rows (multiples of 16, zero padded)
    uint32_t units;               // 16-byte units per row
    const uint32_t* a_index;      // pair -> row of a (nullptr: pair index)
    const int64_t* b_index;       // pair -> row of b, -1 = missing (nullptr: pair index)
    uint64_t n_pairs;
    float* out;
    uint32_t qc_kind;
};

template <class T>
__global__ void __launch_bounds__(256) pair_kernel(const PairParams p) {
    constexpr int LPR = T::LPR;
    constexpr int GPB = 256 / LPR;  // groups per block
    const int lane = threadIdx.x & 31;
    const int g = lane % LPR;
    const uint64_t n_iter = (p.n_pairs + GPB - 1) / GPB;
    for (uint64_t itn = blockIdx.x; itn < n_iter; itn += gridDim.x) {
        const uint64_t pair = itn * GPB + threadIdx.x / LPR;
        const bool in = pair < p.n_pairs;
        int64_t bi = in ? (p.b_index ? p.b_index[pair] : (int64_t)pair) : -1;
        const uint64_t ai = in ? (p.a_index ? p.a_index[pair] : pair) : 0;
        const uint4* a = (const uint4*)(p.a_base + ai * p.a_stride);
        const uint4* b = (const uint4*)(p.b_base + (bi < 0 ? 0 : bi) * (uint64_t)p.b_stride);
        typename T::Acc acc;
        T::init(acc);
        float qc = 0.f;
        if (T::HAS_QC) {
            // all 32 lanes participate in the shuffles; groups of 4 stay aligned because LPR==4 here
            qc = query_const(a, (in && bi >= 0) ? p.units : 0, g, p.qc_kind);  // empty candidate slots cost nothing
        }
        if (in && bi >= 0) {
            for (uint32_t u = g; u < p.units; u += LPR) {
                uint4 x = __ldg(b + u);
                uint4 qv[1] = {__ldg(a + u)};
                T::step(acc, x, qv);
            }
        }
        float d = T::finish(acc, 0, &qc);
        if (in && g == 0) p.out[pair] = bi >= 0 ? d : __int_as_float(0x7FC00000);
    }
}

// compaction: row j of the new slab = row keep[j] of the old one (one warp per row, 16-byte units)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* src, const uint32_t* keep, uint64_t n, uint32_t units, uint8_t* dst) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = w; j < n; j += nw) {
        const uint4* a = (const uint4*)(src + (uint64_t)keep[j] * units * 16);
        uint4* b = (uint4*)(dst + j * units * 16);
        for (uint32_t u = lane; u < units; u += 32) b[u] = a[u];
    }
}

// rowid -> position (binary search over the ascending rowid array, or dense
// arithmetic) and pair -> query index from the CSR offsets; skipped rows -> -1.
__global__ void resolve_kernel(const int64_t* cand_rowids, uint64_t n_pairs, const uint32_t* offsets, uint32_t nq,
                               const int64_t* rowids, uint64_t n_rows, int64_t first_rowid, const uint8_t* skip,
                               uint32_t* out_q, int64_t* out_pos) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        // query of pair i: last q with offsets[q] <= i
        uint32_t lo = 0, hi = nq;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= i) lo = mid; else hi = mid;
        }
        out_q[i] = lo;
        const int64_t r = cand_rowids[i];
        int64_t pos = -1;
        if (rowids == nullptr) {
            // dense: guard the subtraction against overflow
            if (r >= first_rowid && (uint64_t)(r - first_rowid) < n_rows) pos = r - first_rowid;
        } else {
            uint64_t a = 0, b = n_rows;
            while (a < b) {
                uint64_t mid = (a + b) >> 1;
                if (rowids[mid] < r) a = mid + 1; else b = mid;
            }
            if (a < n_rows && rowids[a] == r) pos = (int64_t)a;
        }
        if (pos >= 0 && skip && skip[pos]) pos = -1;
        out_pos[i] = pos;
    }
}

// K5s: ONE launch for a small scoring call (one search_layer expansion: a query and its <= 64 neighbour rowids).  The input
// block — [rowids np x 8][offsets (nq + 1) x 4, padded to 16][queries nq x stride] — sits in pinned host memory and is read
// by the kernel directly (zero-copy), the distances are stored straight into pinned host memory: no copy commands, no
// separate rowid -> position pass.  Every block stages the whole block of input in shared memory (<= 16 KB).
struct ScoreSmallParams {
    const uint8_t* in;         // pinned host (or device) input block
    uint32_t in_bytes;         // multiple of 16
    uint32_t np, nq, q_off;    // pairs, queries, byte offset of the queries in `in`
    const int64_t* rowids;     // slab rowids (nullptr: dense)
    uint64_t n_rows;
    int64_t first_rowid;
    const uint8_t* skip;       // or nullptr
    const uint8_t* b_base;
    uint32_t stride, units, qc_kind;
    float* out;                // [np], pinned host (or device)
};

template <class T>
__global__ void __launch_bounds__(256) score_small_kernel(const ScoreSmallParams p) {
    constexpr int LPR = T::LPR;
    constexpr int GPB = 256 / LPR;
    extern __shared__ __align__(16) uint8_t ss_smem[];
    for (uint32_t u = threadIdx.x; u < p.in_bytes / 16; u += 256) ((uint4*)ss_smem)[u] = ((const uint4*)p.in)[u];
    __syncthreads();
    const int64_t* cand = (const int64_t*)ss_smem;
    const uint32_t* offsets = (const uint32_t*)(ss_smem + (size_t)p.np * 8);
    const int lane = threadIdx.x & 31, g = lane % LPR;
    const uint32_t pair = blockIdx.x * GPB + threadIdx.x / LPR;
    const bool in = pair < p.np;
    uint32_t qi = 0;
    int64_t pos = -1;
    if (in) {
        uint32_t lo = 0, hi = p.nq;  // query of the pair: last q with offsets[q] <= pair
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= pair) lo = mid;
            else hi = mid;
        }
        qi = lo;
        const int64_t r = cand[pair];
        if (p.rowids == nullptr) {
            if (r >= p.first_rowid && (uint64_t)(r - p.first_rowid) < p.n_rows) pos = r - p.first_rowid;
        } else {
            uint64_t a = 0, b = p.n_rows;
            while (a < b) {
                const uint64_t mid = (a + b) >> 1;
                if (p.rowids[mid] < r) a = mid + 1;
                else b = mid;
            }
            if (a < p.n_rows && p.rowids[a] == r) pos = (int64_t)a;
        }
        if (pos >= 0 && p.skip && p.skip[pos]) pos = -1;
    }
    const uint4* a = (const uint4*)(ss_smem + p.q_off + (size_t)qi * p.stride);
    const uint4* b = (const uint4*)(p.b_base + (uint64_t)(pos < 0 ? 0 : pos) * p.stride);
    typename T::Acc acc;
    T::init(acc);
    float qc = 0.f;
    if (T::HAS_QC) qc = query_const(a, (in && pos >= 0) ? p.units : 0, g, p.qc_kind);
    if (in && pos >= 0) {
#pragma unroll 1
        for (uint32_t u0 = g; u0 < p.units; u0 += LPR * 4) {  // four row pieces in flight per lane
            uint4 xv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t u = u0 + (uint32_t)j * LPR;
                if (u < p.units) xv[j] = __ldg(b + u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t u = u0 + (uint32_t)j * LPR;
                if (u < p.units) {
                    const uint4 qv[1] = {a[u]};
                    T::step(acc, xv[j], qv);
                }
            }
        }
    }
    const float d = T::finish(acc, 0, &qc);
    if (in && g == 0) p.out[pair] = pos >= 0 ? d : __int_as_float(0x7FC00000);
}

// ---------------------------------------------------------------------------
// K7 producers — src/vector.rs:444-608, bit-for-bit (IEEE ops, no contraction)
// ---------------------------------------------------------------------------
// normalize (vector.rs:444-466): strict left-to-right f32 sum of x*x.  One thread per row.
__global__ void normalize_kernel(const float* in, uint64_t n, uint32_t d, float* out, int* zero_flag) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const float* v = in + r * d;
        float s = 0.f;
        for (uint32_t i = 0; i < d; ++i) s = __fadd_rn(s, __fmul_rn(v[i], v[i]));
        float m = __fsqrt_rn(s);
        if (m == 0.f) {
            atomicExch(zero_flag, 1);
            for (uint32_t i = 0; i < d; ++i) out[r * d + i] = v[i];
        } else {
            for (uint32_t i = 0; i < d; ++i) out[r * d + i] = __fdiv_rn(v[i], m);
        }
    }
}

__device__ __forceinline__ int8_t quant_i8(float v, float mn, float range) {
    float normalized = __fdiv_rn(__fsub_rn(v, mn), range);
    float scaled = __fsub_rn(__fmul_rn(normalized, 255.0f), 128.0f);
    float r = roundf(scaled);  // half away from zero == Rust f32::round
    r = r < -128.0f ? -128.0f : (r > 127.0f ? 127.0f : r);
    return (int8_t)r;
}

// quantize_int8 (vector.rs:514-545): one warp per row
__global__ void quantize_int8_kernel(const float* in, uint64_t n, uint32_t d, int8_t* out, uint32_t out_stride) {
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = w0; r < n; r += nw) {
        const float* v = in + r * d;
        float mn = __int_as_float(0x7F800000), mx = __int_as_float(0xFF800000);
        for (uint32_t i = lane; i < d; i += 32) {
            mn = fminf(mn, v[i]);
            mx = fmaxf(mx, v[i]);
        }
        for (int m = 16; m >= 1; m >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        }
        int8_t* o = out + r * out_stride;
        if (mn == mx) {
            for (uint32_t i = lane; i < d; i += 32) o[i] = 0;
        } else {
            const float range = __fsub_rn(mx, mn);
            for (uint32_t i = lane; i < d; i += 32) o[i] = quant_i8(v[i], mn, range);
        }
    }
}

// quantize_int8_for_index (vector.rs:554-575)
__global__ void quantize_index_kernel(const float* in, uint64_t total, int8_t* out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        float v = in[i];
        float c = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
        out[i] = (int8_t)roundf(__fmul_rn(c, 127.0f));
    }
}

// HNSW node vectors as the reference STORES them (src/hnsw/insert.rs:300-322): a float32 column row is normalised when the
// column's metric is cosine (Vector::normalize, vector.rs:444-466: strict-order sum of squares, IEEE sqrt and divisions) and
// then, with index_quantization=int8, quantised by quantize_int8_for_index (vector.rs:554-575).  One thread per row (the
// normalisation is a strict left-to-right sum), source and destination are slab rows (strided, zero padded).  A row that
// cannot be normalised (zero magnitude: the reference's insert fails, vector.rs:451-455) is flagged in skip_out.
__global__ void hnsw_stored_rows_kernel(const uint8_t* src, uint32_t src_stride, uint64_t n, uint32_t d, int do_norm, int do_q8,
                                        uint8_t* dst, uint32_t dst_stride, const uint8_t* skip_in, uint8_t* skip_out) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const float* v = (const float*)(src + r * src_stride);
        uint8_t* o = dst + r * dst_stride;
        bool dead = skip_in != nullptr && skip_in[r] != 0;
        float m = 1.f;
        if (do_norm && !dead) {
            float s = 0.f;
            for (uint32_t i = 0; i < d; ++i) s = __fadd_rn(s, __fmul_rn(v[i], v[i]));
            m = __fsqrt_rn(s);
            if (m == 0.f) dead = true;
        }
        for (uint32_t i = 0; i < d; ++i) {
            float x = dead ? 0.f : (do_norm ? __fdiv_rn(v[i], m) : v[i]);
            if (do_q8) {
                const float c = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
                ((int8_t*)o)[i] = (int8_t)roundf(__fmul_rn(c, 127.0f));
            } else {
                ((float*)o)[i] = x;
            }
        }
        const uint32_t used = do_q8 ? d : d * 4u;
        for (uint32_t i = used; i < dst_stride; ++i) o[i] = 0;
        if (skip_out) skip_out[r] = dead ? 1 : 0;
    }
}

// quantize_binary (vector.rs:579-608): strict-order mean, one thread per row
__global__ void quantize_binary_kernel(const float* in, uint64_t n, uint32_t d, uint8_t* out) {
    const uint32_t nb = (d + 7) / 8;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const float* v = in + r * d;
        float s = 0.f;
        for (uint32_t i = 0; i < d; ++i) s = __fadd_rn(s, v[i]);
        const float mean = __fdiv_rn(s, (float)d);
        for (uint32_t b = 0; b < nb; ++b) {
            uint32_t byte = 0;
            for (uint32_t j = 0; j < 8 && b * 8 + j < d; ++j)
                if (v[b * 8 + j] >= mean) byte |= 1u << j;
            out[r * nb + b] = (uint8_t)byte;
        }
    }
}

// ---------------------------------------------------------------------------
// synthetic corpus generator value(seed, rowid, word) of SURVEY §8d (the CPU checker restates it)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_word(uint64_t seed, int64_t rowid, uint32_t w) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)rowid);
    return mix64(h + 0xD1B54A32D192ED03ull * (uint64_t)(w + 1));
}
__device__ __forceinline__ float synth_f32(uint64_t seed, int64_t rowid, uint32_t j, int kind) {
    if (kind == 1) {
        uint64_t h = synth_word(seed, rowid, j);
        int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)((h >> 48) & 0xFFFF) -
                131070;
        return __fmul_rn((float)s, 0x1p-15f);
    }
    uint64_t h = synth_word(seed, rowid, j >> 1);
    uint32_t u = (j & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
    return __fsub_rn(__fmul_rn((float)(u >> 8), 0x1p-23f), 1.0f);
}

// f32 rows: one thread per element, row_stride/4 floats per stored row (padding written as 0)
__global__ void synth_f32_kernel(float* out, uint32_t stride_f, uint32_t d, uint64_t n, uint64_t seed,
                                 int64_t first_rowid, int kind) {
    const uint64_t total = n * stride_f;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / stride_f;
        const uint32_t j = (uint32_t)(i - r * stride_f);
        out[i] = j < d ? synth_f32(seed, first_rowid + (int64_t)r, j, kind) : 0.f;
    }
}

// i8 rows: quantize_int8(uniform f32 row); one warp per row
__global__ void synth_i8_kernel(int8_t* out, uint32_t stride, uint32_t d, uint64_t n, uint64_t seed, int64_t first_rowid) {
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = w0; r < n; r += nw) {
        const int64_t rowid = first_rowid + (int64_t)r;
        float mn = __int_as_float(0x7F800000), mx = __int_as_float(0xFF800000);
        for (uint32_t i = lane; i < d; i += 32) {
            float v = synth_f32(seed, rowid, i, 0);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        for (int m = 16; m >= 1; m >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        }
        int8_t* o = out + r * stride;
        const float range = __fsub_rn(mx, mn);
        for (uint32_t i = lane; i < stride; i += 32) {
            int8_t q = 0;
            if (i < d && mn != mx) q = quant_i8(synth_f32(seed, rowid, i, 0), mn, range);
            o[i] = q;
        }
    }
}

// bit rows: one thread per 8-byte word of the stored row
__global__ void synth_bit_kernel(uint8_t* out, uint32_t stride, uint32_t d, uint64_t n, uint64_t seed, int64_t first_rowid) {
    const uint32_t wpr = stride / 8;  // stride is a multiple of 16
    const uint64_t total = n * wpr;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / wpr;
        const uint32_t w = (uint32_t)(i - r * wpr);
        uint64_t h = synth_word(seed, first_rowid + (int64_t)r, w);
        const uint64_t bit0 = (uint64_t)w * 64;
        if (bit0 >= d) h = 0;
        else if (d - bit0 < 64) h &= (1ull << (d - bit0)) - 1ull;
        ((uint64_t*)out)[i] = h;
    }
}

// pad-copy: host-layout rows (src_bytes each) -> slab rows (row_stride each, zero padded)
__global__ void pad_rows_kernel(const uint8_t* src, uint32_t src_bytes, uint8_t* dst, uint32_t dst_stride, uint64_t n) {
    const uint64_t total = n * dst_stride;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / dst_stride;
        const uint32_t b = (uint32_t)(i - r * dst_stride);
        dst[i] = b < src_bytes ? src[r * src_bytes + b] : 0;
    }
}

}  // namespace vg
