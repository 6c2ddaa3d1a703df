// tc_batch.cuh — K2 / K3 batched: many queries against the slab as a tensor-core contraction (tcgen05, TMEM, TMA).
//
// Replaces, for nq >> 1 queries, the per-pair simsimd calls of the exact scan
// (src/vtab.rs:2594-2616 -> src/distance/scalar.rs:17,48,61) by S = Q * X^T on the 5th-gen tensor cores:
//     ||q - x||^2 = ||q||^2 + ||x||^2 - 2 q.x          cos(q,x) = q.x / (||q|| ||x||)
//
// tc_scan_kernel (float32, kind::tf32).  The product only SELECTS candidates: every query keeps its k+32 best
// approximate scores per CTA, a rigorous error bound turns them into a candidate superset of the exact top-k
// (tc_collect_kernel), and the candidates are re-scored by pair_kernel in the canonical fp32 order, so the final rowids
// and distances are bit-identical to the single-query scan.  Queries whose bound cannot be certified (massive ties)
// fall back to the exact scan.  Two ways to form q.x (TcParams::terms):
//   1 (default)  one TF32 pass: the tensor core TRUNCATES fp32 operands to tf32 (measured, tools/umma_test.cu), the
//                error |q.x - tf32(q).tf32(x)| <= 2^-9 |q||x| (Cauchy-Schwarz) goes into the certified bound;
//   3            3xTF32: hi = x as the hardware sees it, lo = x - trunc_tf32(x) (exact in fp32),
//                q.x ~= hi_q.hi_x + lo_q.hi_x + hi_q.lo_x (lo.lo ~ 2^-22 dropped); the lo tiles are made in shared memory.
// One CTA per SM, 12 warps, warp-specialised:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) of a 128x32 query chunk + 256x32 row chunk
//               (optionally multicast to a thread-block cluster), soft lockstep with the CTAs that stream the same rows
//   warp 1      tcgen05.mma issuer (one lane); owns the TMEM allocation (512 columns = 2 accumulators)
//   warps 4-7   epilogue: tcgen05.ld of the 128x256 fp32 accumulator, score -> pass mask -> append buffer -> compaction
//   warps 8-11  (terms = 3 only) transform: lo = x - trunc(x) for both operands, written to the twin "lo" tiles
// Pipelines: full_raw/full_lo/empty per smem stage, tmem_full/tmem_empty per accumulator.
//
// tci8_scan_kernel (int8 L2, kind::i8): the int32 accumulator is exact, so the epilogue emits the FINAL keys; see the
// second half of this file.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace vg {

static constexpr uint32_t TC_M = 128;      // queries per CTA tile (UMMA M)
static constexpr uint32_t TC_N = 256;      // slab rows per tile (UMMA N)
// f32 kernel: k-chunks of 32 floats = 128-byte rows (SWIZZLE_128B), two 96 KB stages.  (16-float chunks with
// SWIZZLE_64B and four 48 KB stages were measured 10 % slower: twice the barrier round trips per K.)
static constexpr uint32_t TC_KC = 32;
static constexpr uint32_t TC_ROW_BYTES = TC_KC * 4;
static constexpr uint32_t TC_KSTEPS = TC_ROW_BYTES / 32;  // tf32 MMAs (K=8) per chunk
static constexpr uint32_t TC_STAGES = 2;
static constexpr uint32_t TC_A_BYTES = TC_M * TC_ROW_BYTES, TC_B_BYTES = TC_N * TC_ROW_BYTES;
static constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // raw + lo of both operands = 96 KB
static constexpr uint32_t TC_THREADS = 384;            // warps 0-1 TMA/MMA, 4-7 epilogue, 8-11 lo-split transform
static constexpr uint32_t TC_XFORM_THREADS = 128;     // (8 transform warps were measured: no gain)
// int8 kernel: 128-byte rows (SWIZZLE_128B)
static constexpr uint32_t TCI_A_BYTES = TC_M * 128, TCI_B_BYTES = TC_N * 128;
// kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=128
static constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

// A squared norm outside this range (zero, denormal-ish, huge, inf, NaN) makes the approximate score of the row
// meaningless (overflow / cancellation), so the row bypasses the contraction and is always re-ranked exactly.
static constexpr uint32_t TC_MAX_UNSAFE = 64;
static constexpr uint32_t TC_BUF_CAP = 256;     // per-thread append buffer of the epilogue (kp + 32 <= TC_BUF_CAP)
__host__ __device__ __forceinline__ bool tc_norm_safe(float x2) { return x2 > 1e-30f && x2 < 1e30f; }

struct TcParams {
    uint64_t n_rows;
    uint32_t nq, nk;       // queries, k-chunks (ceil(dims/32); TMA zero-fills the ragged tail)
    uint32_t kp;           // entries kept per (CTA, query) = k + margin
    uint32_t cosine;       // 0: L2 (v = x2 - 2s), 1: cosine (v = -s / |x|)
    uint32_t terms;        // 3: 3xTF32 (hi.hi + lo.hi + hi.lo), 1: one TF32 pass (wider certified bound)
    uint32_t debug;        // timing experiments (results invalid): 1 = no epilogue work, 2 = common path only, 4 = no norm loads
    uint32_t QT, G;        // query tiles, row-tile groups; grid = QT*G, CTA c -> (qt = c % QT, g = c / QT)
    uint32_t* lockstep;    // [G][32] tile counters of the query-tile CTAs of each row group (zeroed before the launch), or NULL
    uint32_t lock_slack;   // how many tiles a peer may be behind
    uint32_t cs;           // cluster size (1, 2, 4, 8; divides QT): the CTAs of a cluster share every row tile through TMA multicast
    uint32_t pair;         // 1 (with cs == 2, terms == 1): the two CTAs of a cluster form a tcgen05 cta_group::2 PAIR — M = 256
                           // queries x N = 256 rows per instruction, each CTA stages its own 128 queries and HALF of the row tile
    const float* norms;    // [rows] canonical sum of squares of each slab row
    const uint8_t* skip;   // per-row skip flags or nullptr
    uint64_t* buf_keys;    // [grid][128][TC_BUF_CAP] per-thread append buffers
    float* cand_v;         // [grid][128][kp] approximate scores kept
    uint32_t* cand_r;      // [grid][128][kp] row positions
    uint32_t* cand_cnt;    // [grid][128]
    float* cand_tau;       // [grid][128] largest kept score when the list is full, else +inf
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// shared-memory matrix descriptors, K-major: 8-row groups are SBO bytes apart; layout 2 = SWIZZLE_128B (128-byte rows),
// layout 4 = SWIZZLE_64B (64-byte rows).  Both verified on the device (tools/umma_test.cu; the SWIZZLE_64B variant differs only in the layout field and the 64-byte row pitch).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// cluster variants: a row-tile slice is loaded once and written into the same shared-memory offset of every CTA in `mask`
// (each destination's mbarrier at the same offset receives the bytes); the commit arrives on the barrier of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
// ---- cta_group::2 (CTA pair) forms, validated by tools/umma2_test.cu ----
// Both CTAs issue their own TMA loads; the bytes complete on the LEADER's mbarrier (same offset, peer bit of the
// shared::cluster address cleared), whose owner then issues one M = 256 MMA for the pair.
static constexpr uint32_t TC_PEER_BIT_MASK = 0xFEFFFFFFu;
static constexpr uint32_t TC_IDESC_PAIR = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_N >> 3) << 17) | ((256u >> 4) << 24);  // M = 256
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(leader_bar & TC_PEER_BIT_MASK)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC_PAIR), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}\n" ::"r"(bar),
                 "r"(rank)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier wait that traps instead of spinning forever (a protocol bug must not hang the GPU)
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; spins < (1u << 28); ++spins)
        if (mbar_test(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t tf32_lo(uint32_t xbits) {
    const float x = __uint_as_float(xbits);
    return __float_as_uint(__fsub_rn(x, __uint_as_float(xbits & 0xFFFFE000u)));  // exact
}

// Soft lockstep of the QT CTAs that stream the same row tiles (one per query tile): before loading tile ti the producer warp
// publishes its tile index and waits until its peers have reached the same tile, so the QT requests for a tile hit L2
// within a few microseconds of each other and the slab is read from HBM once, not once per query tile (measured: 58 %
// of the row-tile requests missed L2 without it; the window in which a line survives the streaming is about one tile).
// Purely a performance hint: after a bounded number of polls (a peer CTA is not resident, ...) the CTA stops waiting for
// the rest of the kernel.  Returns false once lockstep has been abandoned.
__device__ __forceinline__ bool tile_lockstep(uint32_t* ctr, uint32_t qt, uint32_t QT, uint32_t ti, uint32_t slack, int lane) {
    volatile uint32_t* c = ctr;
    if (lane == 0) c[qt] = ti + 1;
    if (ti < slack) return true;
    for (uint32_t polls = 0; polls < 2048; ++polls) {
        const uint32_t v = (uint32_t)lane < QT ? c[lane] : 0xFFFFFFFFu;
        if (__all_sync(0xffffffffu, v + slack >= ti + 1)) return true;  // no peer is more than `slack` tiles behind
    }
    return false;
}

// (float)sqrt((double)s) for an exact non-negative integer s (src/distance/scalar.rs:65).  Below 2^24 the integer is a
// float, and the correctly rounded float square root equals the double-rounded value (double rounding is innocuous for
// sqrt when the wide format has >= 2*24+2 bits), so no FP64 instruction is needed; tests/test_oracle_golden.py checks
// the identity exhaustively on the CPU.
__device__ __forceinline__ float exact_sqrt_int(int s) {
    if (s < (1 << 24)) return __fsqrt_rn((float)s);
    return __double2float_rn(__dsqrt_rn((double)s));
}

// Integer bound that goes with a key bound: any squared distance s beyond (next float after the key's distance)^2
// has a strictly larger f32 distance, so it can never beat the key.
__device__ __forceinline__ int tci_tau_s(uint64_t key) {
    const float dr = order_bits_inv((uint32_t)(key >> 32));
    const double dn = (double)__uint_as_float(__float_as_uint(dr) + 1u);
    const double lim = dn * dn;
    return lim < 2147483000.0 ? (int)lim : 0x7FFFFFFE;
}

// Warp-cooperative compaction of one thread's append buffer (at most 256 keys): the keys are loaded 8 per lane,
// sorted ascending across the warp with a bitonic network that lives entirely in registers (strides below 8 are
// register-to-register compare-exchanges, larger strides are lane shuffles; the "flip" form of the network needs no
// direction flags), and the k smallest are written back to the front of the buffer (and to `out` when given).
// Returns (to every lane) the k-th smallest key, KEY_NONE when fewer than k keys exist.  ~1.2k instructions per
// lane: cheap enough that a compaction no longer stalls the accumulator pipeline.
__device__ __noinline__ uint64_t warp_compact(uint64_t* buf, uint32_t cnt, uint32_t k, uint64_t* out, int lane) {
    __syncwarp();  // the owner lane's appends become visible to the helping lanes
    uint64_t v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t i = (uint32_t)lane * 8 + r;
        v[r] = i < cnt ? buf[i] : KEY_NONE;
    }
    warp_sort256(v, lane);
    uint64_t mine = v[0];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t i = (uint32_t)lane * 8 + r;
        if (i < k) {
            buf[i] = v[r];
            if (out) out[i] = v[r];
        }
        if (((k - 1) & 7u) == (uint32_t)r) mine = v[r];
    }
    const uint64_t kth = shfl_u64(mine, (int)((k - 1) >> 3));
    __syncwarp();
    return kth;
}

// TMEM -> registers, 32 lanes x 32 columns; the wait names the registers so that no use is scheduled before it
#define TCI_LD32(v, addr)                                                                                                  \
    asm volatile(                                                                                                          \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),    \
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),       \
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),       \
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                               \
        : "r"(addr))
#define TCI_WAIT32(v)                                                                                                      \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                          \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),      \
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),           \
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),          \
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])                        \
                 :                                                                                                         \
                 : "memory")

// PAIR = true is the cta_group::2 instantiation: it must be launched with a cluster of exactly two CTAs (a kernel that contains
// cta_group::2 instructions cannot be launched without a cluster at all, so the two forms are separate kernels).
template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapX, const TcParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries: align by hand (the launch reserves 1 KB of slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout: [stage][A_raw | A_lo | B_raw | B_lo] ... [colA[4][256] | colB[4][256]] [barriers] [tmem slot] [stage[4][2 KB]]
    float* s_colA = (float*)(smem + TC_STAGES * TC_STAGE_BYTES);   // [4 epilogue warps][256] per-column scale
    float* s_colB = s_colA + 4 * TC_N;                             // [4 epilogue warps][256] per-column offset
    uint64_t* s_bar = (uint64_t*)(s_colB + 4 * TC_N);
    uint32_t* s_tmem = (uint32_t*)(s_bar + 16);
    uint8_t* s_stage = (uint8_t*)(s_bar + 32);                     // [4 epilogue warps][2 KB] staging of half a chunk (rare path)
    // terms == 3: two 96 KB stages [A_raw | A_lo | B_raw | B_lo]; terms == 1: four 48 KB stages [A | B], no lo-split;
    // pair: six 32 KB stages [A (this CTA's 128 queries) | B half (this CTA's 128 rows of the tile)]
    constexpr bool pair = PAIR;
    const uint32_t n_stages = pair ? 6u : (p.terms == 1 ? 2 * TC_STAGES : TC_STAGES);
    const uint32_t stage_bytes = pair ? TC_A_BYTES + TC_B_BYTES / 2 : (p.terms == 1 ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES);
    const uint32_t b_off = p.terms == 1 ? TC_A_BYTES : 2 * TC_A_BYTES;
    // 16 barrier slots: full_raw[n] empty[n] tfull[2] tempty[2] full_lo[n] (full_lo only exists for terms == 3, n = 2)
    const uint32_t bar_full_raw = smem_u32(s_bar), bar_empty = smem_u32(s_bar + n_stages), bar_tfull = smem_u32(s_bar + 2 * n_stages),
                   bar_tempty = smem_u32(s_bar + 2 * n_stages + 2), bar_full_lo = smem_u32(s_bar + 2 * n_stages + 4);

    if (warp == 1) {
        if constexpr (pair) {  // both CTAs of the pair allocate, same warp, same slot address
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(bar_full_raw + 8 * s, 1);
            if (p.terms != 1) mbar_init(bar_full_lo + 8 * s, TC_XFORM_THREADS / 32);   // one arrival per transform warp
            mbar_init(bar_empty + 8 * s, pair ? 1u : p.cs);  // tcgen05.commit of every CTA that receives the multicast row tiles (pair: ONE multicast commit)
        }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);     // tcgen05.commit
            mbar_init(bar_tempty + 8 * a, pair ? 8u : 4u);    // one arrival per epilogue warp (pair: of BOTH CTAs, on the leader's barrier)
        }
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.cs > 1) cluster_sync_all();  // every CTA's barriers exist before the first multicast can land
    const uint32_t tmem_base = *s_tmem;
    uint32_t crank = blockIdx.x % p.cs;                          // rank in the cluster (consecutive query tiles of one row group)
    if (p.cs > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
    const uint32_t slice_rows = TC_N / p.cs;                     // rows of every tile this CTA fetches for the whole cluster

    const uint32_t qt = blockIdx.x % p.QT, g = blockIdx.x / p.QT;
    const uint64_t n_xt = (p.n_rows + TC_N - 1) / TC_N;
    const uint32_t my_tiles = g < n_xt ? (uint32_t)((n_xt - g + p.G - 1) / p.G) : 0u;

    if (warp == 0) {
        // ===== TMA producer =====
        uint32_t it = 0;
        bool lock = p.lockstep != nullptr;
        for (uint32_t ti = 0; ti < my_tiles; ++ti) {
            if (lock) lock = tile_lockstep(p.lockstep + (size_t)g * 32, qt, p.QT, ti, p.lock_slack, lane);
            if (lane == 0) {
                const int row0 = (int)(((uint64_t)g + (uint64_t)ti * p.G) * TC_N);
                for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                    mbar_wait_b(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t base = smem_u32(smem + s * stage_bytes);
                    if constexpr (pair) {
                        // this CTA's 128 queries and its half of the row tile; all four loads of the pair complete on the
                        // leader's barrier, which expects the 64 KB
                        if (crank == 0) mbar_expect_tx(bar_full_raw + 8 * s, 2 * (TC_A_BYTES + TC_B_BYTES / 2));
                        tma_load_2d_pair(base, &mapQ, (int)(kc * TC_KC), (int)(qt * TC_M), bar_full_raw + 8 * s);
                        tma_load_2d_pair(base + b_off, &mapX, (int)(kc * TC_KC), row0 + (int)(crank * (TC_N / 2)), bar_full_raw + 8 * s);
                        continue;
                    }
                    mbar_expect_tx(bar_full_raw + 8 * s, TC_A_BYTES + TC_B_BYTES);
                    tma_load_2d(base, &mapQ, (int)(kc * TC_KC), (int)(qt * TC_M), bar_full_raw + 8 * s);
                    if (p.cs > 1)
                        tma_load_2d_mc(base + b_off + crank * slice_rows * TC_ROW_BYTES, &mapX, (int)(kc * TC_KC),
                                       row0 + (int)(crank * slice_rows), bar_full_raw + 8 * s, cmask);
                    else
                        tma_load_2d(base + b_off, &mapX, (int)(kc * TC_KC), row0, bar_full_raw + 8 * s);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The hi.hi MMAs of a chunk only need the TMA data, the two lo terms also need the lo-split.  While the issuer
        // waits for the lo-split of chunk `it` it opportunistically issues hi.hi of chunk it+1 as soon as that chunk's
        // data has landed (non-blocking probes), so the tensor pipe has work during the transform.
        if constexpr (pair) {
            // CTA pair: only the leader issues; one instruction covers the pair's 256 queries x the tile's 256 rows, reading
            // each CTA's A tile and B half from that CTA's shared memory — per SM and K-step 8 KB are written by TMA and 8 KB
            // read by the tensor core instead of 12 + 12 (the single-CTA form is bound by exactly that shared-memory traffic)
            if (lane == 0 && crank == 0) {
                uint32_t it = 0;
                for (uint32_t ti = 0; ti < my_tiles; ++ti) {
                    const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
                    mbar_wait_b(bar_tempty + 8 * acc, aph ^ 1);  // the epilogue warps of BOTH CTAs have drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TC_N;
                    for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                        const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                        const uint32_t a = smem_u32(smem + s * stage_bytes), b = a + b_off;
                        mbar_wait_b(bar_full_raw + 8 * s, ph);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t k = 0; k < TC_KSTEPS; ++k) umma2_tf32(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), (kc | k) != 0);
                        umma2_commit_mc(bar_empty + 8 * s, 0b11);  // both producers may refill the stage
                    }
                    umma2_commit_mc(bar_tfull + 8 * acc, 0b11);   // both CTAs' epilogues may read their half of the accumulator
                }
            }
        } else if (lane == 0 && p.terms == 1) {
            // single TF32 pass: the operands are used as they are (the tensor core truncates them to TF32); the wider
            // error bound is paid for by a slightly larger certified candidate set, not by two more MMAs
            uint32_t it = 0;
            for (uint32_t ti = 0; ti < my_tiles; ++ti) {
                const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
                mbar_wait_b(bar_tempty + 8 * acc, aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_N;
                for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                    const uint32_t a = smem_u32(smem + s * stage_bytes), b = a + b_off;
                    mbar_wait_b(bar_full_raw + 8 * s, ph);
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 0; k < TC_KSTEPS; ++k) umma_tf32(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), (kc | k) != 0);
                    if (p.cs > 1) umma_commit_mc(bar_empty + 8 * s, cmask);
                    else umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        } else if (lane == 0) {
            const uint32_t total = my_tiles * p.nk;
            uint32_t next_hihi = 0;  // first chunk whose hi.hi MMAs have not been issued yet
            auto hihi_ready = [&](uint32_t it) {
                const uint32_t ti = it / p.nk, kc = it - ti * p.nk;
                if (kc == 0 && !mbar_test(bar_tempty + 8 * (ti & 1), ((ti >> 1) & 1) ^ 1)) return false;
                return mbar_test(bar_full_raw + 8 * (it % TC_STAGES), (it / TC_STAGES) & 1);
            };
            auto issue_hihi = [&](uint32_t it) {  // blocking
                const uint32_t ti = it / p.nk, kc = it - ti * p.nk;
                const uint32_t acc = ti & 1, s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                if (kc == 0) mbar_wait_b(bar_tempty + 8 * acc, ((ti >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                const uint32_t a_raw = smem_u32(smem + s * TC_STAGE_BYTES), b_raw = a_raw + 2 * TC_A_BYTES;
                mbar_wait_b(bar_full_raw + 8 * s, ph);
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < TC_KSTEPS; ++k)
                    umma_tf32(tmem_base + acc * TC_N, umma_desc(a_raw + k * 32), umma_desc(b_raw + k * 32), (kc | k) != 0);
            };
            for (uint32_t it = 0; it < total; ++it) {
                if (next_hihi <= it) {
                    issue_hihi(it);
                    next_hihi = it + 1;
                }
                const uint32_t ti = it / p.nk, kc = it - ti * p.nk;
                const uint32_t acc = ti & 1, s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                while (!mbar_test(bar_full_lo + 8 * s, ph)) {
                    if (next_hihi == it + 1 && next_hihi < total && hihi_ready(next_hihi)) {
                        issue_hihi(next_hihi);
                        ++next_hihi;
                    }
                }
                tc_fence_after();
                const uint32_t a_raw = smem_u32(smem + s * TC_STAGE_BYTES), a_lo = a_raw + TC_A_BYTES,
                               b_raw = a_raw + 2 * TC_A_BYTES, b_lo = b_raw + TC_B_BYTES;
                const uint32_t d_tmem = tmem_base + acc * TC_N;
#pragma unroll
                for (uint32_t k = 0; k < TC_KSTEPS; ++k)  // lo_q.hi_x
                    umma_tf32(d_tmem, umma_desc(a_lo + k * 32), umma_desc(b_raw + k * 32), 1);
#pragma unroll
                for (uint32_t k = 0; k < TC_KSTEPS; ++k)  // hi_q.lo_x
                    umma_tf32(d_tmem, umma_desc(a_raw + k * 32), umma_desc(b_lo + k * 32), 1);
                if (p.cs > 1) umma_commit_mc(bar_empty + 8 * s, cmask);  // stage reusable once these MMAs have read it
                else umma_commit(bar_empty + 8 * s);
                if (kc == p.nk - 1) umma_commit(bar_tfull + 8 * acc);  // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ===== transform warps: lo = x - trunc_tf32(x), same (swizzled) positions in the twin tile =====
        const uint32_t t = threadIdx.x - 256;  // 0..TC_XFORM_THREADS-1
        uint32_t it = 0;
        for (uint32_t ti = 0; ti < (p.terms == 1 ? 0u : my_tiles); ++ti) {
            for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                const uint32_t a_raw = smem_u32(smem + s * TC_STAGE_BYTES), b_raw = a_raw + 2 * TC_A_BYTES;
                mbar_wait_b(bar_full_raw + 8 * s, ph);
#pragma unroll 4
                for (uint32_t u = t; u < TC_A_BYTES / 16; u += TC_XFORM_THREADS) {
                    uint4 v = lds128(a_raw + u * 16);
                    v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
                    sts128(a_raw + TC_A_BYTES + u * 16, v);
                }
#pragma unroll 4
                for (uint32_t u = t; u < TC_B_BYTES / 16; u += TC_XFORM_THREADS) {
                    uint4 v = lds128(b_raw + u * 16);
                    v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
                    sts128(b_raw + TC_B_BYTES + u * 16, v);
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full_lo + 8 * s);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue warps: thread = query (TMEM lane), columns = slab rows of the tile =====
        // Same scheme as tci8_scan_kernel: a branch-free common path (score + pass mask for 32 columns, the next TMEM chunk
        // already in flight), a compact rare path (staged chunk, indexed loop) that APPENDS (order_bits(score) << 32 | row)
        // to the thread's private buffer, and a register bitonic compaction by the whole warp when a buffer fills up
        // (keeps the kp best, tightens the bound).  The per-column coefficients live in a per-warp copy fetched one tile
        // ahead, so the four warps never synchronise with each other.
        const uint32_t ew = (uint32_t)(warp - 4);
        const uint32_t lane_base = ew * 32;
        const uint32_t e = lane_base + lane;         // 0..127 == TMEM lane == query within the tile
        const bool q_ok = qt * TC_M + e < p.nq;      // tile rows past nq hold zero-filled queries: they never admit anything
        const size_t lidx = (size_t)blockIdx.x * TC_M + e;
        uint64_t* buf = p.buf_keys + lidx * TC_BUF_CAP;
        float* my_colA = s_colA + ew * TC_N;
        float* my_colB = s_colB + ew * TC_N;
        int* stage = (int*)(s_stage + ew * 2048);    // [16 columns][32 lanes]
        uint32_t cnt = 0;
        uint64_t tau_key = KEY_NONE;                 // kp-th best key as of the last compaction
        float tau_f = q_ok ? __int_as_float(0x7F800000) : __int_as_float(0xFF800000);  // its score: every better key has score <= tau_f
        float na[TC_N / 32], nb[TC_N / 32];
        auto fetch_coef = [&](uint32_t ti) {          // per-column coefficients: v = s * a + b
            const uint64_t r0 = ((uint64_t)g + (uint64_t)ti * p.G) * TC_N;
#pragma unroll
            for (int i = 0; i < (int)TC_N / 32; ++i) {
                const uint64_t row = r0 + (uint64_t)(i * 32 + lane);
                bool ok = row < p.n_rows && !(p.skip && p.skip[row]);
                const float x2 = ok ? p.norms[row] : 0.f;
                ok = ok && tc_norm_safe(x2);  // unsafe rows are handled by the exact re-rank alone
                if (p.cosine) {
                    na[i] = (ok && x2 > 0.f) ? -rsqrtf(x2) : 0.f;
                    nb[i] = ok ? 0.f : __int_as_float(0x7F800000);
                } else {
                    na[i] = -2.f;
                    nb[i] = ok ? x2 : __int_as_float(0x7F800000);
                }
            }
        };
        if (my_tiles) fetch_coef(0);
        for (uint32_t ti = 0; ti < my_tiles; ++ti) {
            const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
            const uint64_t row0 = ((uint64_t)g + (uint64_t)ti * p.G) * TC_N;
#pragma unroll
            for (int i = 0; i < (int)TC_N / 32; ++i) {
                my_colA[i * 32 + lane] = na[i];
                my_colB[i * 32 + lane] = nb[i];
            }
            __syncwarp();
            if (ti + 1 < my_tiles) fetch_coef(ti + 1);
            mbar_wait_b(bar_tfull + 8 * acc, aph);
            tc_fence_after();
            auto process = [&](uint32_t (&v)[32], const uint32_t c) {
                const float4* ca4 = (const float4*)(my_colA + c * 32);
                const float4* cb4 = (const float4*)(my_colB + c * 32);
                float sc[32];
                uint32_t mask = 0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 a = ca4[j4], b = cb4[j4];
                    sc[4 * j4 + 0] = __fmaf_rn(__uint_as_float(v[4 * j4 + 0]), a.x, b.x);
                    sc[4 * j4 + 1] = __fmaf_rn(__uint_as_float(v[4 * j4 + 1]), a.y, b.y);
                    sc[4 * j4 + 2] = __fmaf_rn(__uint_as_float(v[4 * j4 + 2]), a.z, b.z);
                    sc[4 * j4 + 3] = __fmaf_rn(__uint_as_float(v[4 * j4 + 3]), a.w, b.w);
                }
                // NaN and +inf scores (rows that must not be kept) fail the comparison
#pragma unroll
                for (int j = 0; j < 32; ++j) mask |= ((sc[j] <= tau_f && sc[j] < __int_as_float(0x7F800000)) ? 1u : 0u) << j;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t m16 = (mask >> (16 * hh)) & 0xFFFFu;
                    if (__any_sync(0xffffffffu, m16 != 0)) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) stage[j * 32 + lane] = __float_as_int(sc[16 * hh + j]);
                        while (m16) {
                            const int j = __ffs(m16) - 1;
                            m16 &= m16 - 1;
                            const uint64_t key = make_key(__int_as_float(stage[j * 32 + lane]), (uint32_t)(row0 + c * 32 + 16 * hh + (uint32_t)j));
                            if (key < tau_key) buf[cnt++] = key;
                        }
                        __syncwarp();
                    }
                }
                // a buffer that could overflow during the next 32 columns is compacted now (warp-uniform loop)
                unsigned need = __ballot_sync(0xffffffffu, cnt + 32 > TC_BUF_CAP);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const uint64_t bp = shfl_u64((uint64_t)(uintptr_t)buf, src);
                    const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
                    const uint64_t kth = warp_compact((uint64_t*)(uintptr_t)bp, bc, p.kp, nullptr, lane);
                    if (lane == src) {
                        cnt = min(bc, p.kp);
                        if (kth != KEY_NONE) {
                            tau_key = kth;
                            tau_f = order_bits_inv((uint32_t)(kth >> 32));
                        }
                    }
                }
            };
            if (!(p.debug & 1)) {
                uint32_t va[32], vb[32];
                const uint32_t tbase = tmem_base + (lane_base << 16) + acc * TC_N;
                TCI_LD32(va, tbase);
#pragma unroll 1
                for (uint32_t c = 0; c < TC_N / 32; c += 2) {
                    TCI_WAIT32(va);
                    TCI_LD32(vb, tbase + (c + 1) * 32);
                    process(va, c);
                    TCI_WAIT32(vb);
                    if (c + 2 < TC_N / 32) TCI_LD32(va, tbase + (c + 2) * 32);
                    process(vb, c + 1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (pair) {
                    if (crank != 0) mbar_arrive_remote(bar_tempty + 8 * acc, 0);  // the leader's MMA warp owns the pair's accumulators
                    else mbar_arrive(bar_tempty + 8 * acc);
                } else {
                    mbar_arrive(bar_tempty + 8 * acc);
                }
            }
        }
        // final compaction of every lane's buffer -> the kp best (score, row) of this (CTA, query) in the layout
        // tc_collect_kernel reads; a list that holds kp entries reports its largest kept score as its drop bound
        for (int src = 0; src < 32; ++src) {
            const uint64_t bp = shfl_u64((uint64_t)(uintptr_t)buf, src);
            const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
            const size_t li = (size_t)blockIdx.x * TC_M + lane_base + (uint32_t)src;
            const uint64_t kth = warp_compact((uint64_t*)(uintptr_t)bp, bc, p.kp, nullptr, lane);
            const uint32_t keep = min(bc, p.kp);
            const uint64_t* sorted = (const uint64_t*)(uintptr_t)bp;
            for (uint32_t i = lane; i < keep; i += 32) {
                const uint64_t key = sorted[i];
                p.cand_v[li * p.kp + i] = order_bits_inv((uint32_t)(key >> 32));
                p.cand_r[li * p.kp + i] = (uint32_t)key;
            }
            if (lane == 0) {
                p.cand_cnt[li] = keep;
                p.cand_tau[li] = kth != KEY_NONE ? order_bits_inv((uint32_t)(kth >> 32)) : __int_as_float(0x7F800000);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cs > 1) cluster_sync_all();  // peers may still arrive on this CTA's barriers until they are done too
    if (warp == 1) {
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// canonical sum of squares of every row (the b2 / a2 of F32Cos), the maximum over the safe rows, and the list
// of unsafe rows (unsafe[0] = count, unsafe[1..] = positions of the first TC_MAX_UNSAFE)
// `base` / `norms` point at the first of the n rows; row_base is that row's position in the slab (unsafe[] holds positions)
__global__ void __launch_bounds__(256) row_norms_kernel(const uint8_t* base, uint32_t stride, uint32_t units, uint64_t n, float* norms,
                                                        uint32_t* max_bits, uint32_t* unsafe, uint32_t row_base) {
    const int g = threadIdx.x & 3;
    const uint64_t n_iter = (n + 63) / 64;
    float mx = 0.f;
    for (uint64_t itn = blockIdx.x; itn < n_iter; itn += gridDim.x) {
        const uint64_t row = itn * 64 + (threadIdx.x >> 2);
        const uint4* r = (const uint4*)(base + (row < n ? row : 0) * (uint64_t)stride);
        const float v = query_const(r, row < n ? units : 0, g, 0);
        if (row < n && g == 0) {
            norms[row] = v;
            if (tc_norm_safe(v)) mx = fmaxf(mx, v);
            else if (unsafe) {
                const uint32_t slot = atomicAdd(unsafe, 1u);
                if (slot < TC_MAX_UNSAFE) unsafe[1 + slot] = row_base + (uint32_t)row;
            }
        }
    }
    for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if ((threadIdx.x & 31) == 0 && max_bits) atomicMax(max_bits, __float_as_uint(mx));
}

// an out-of-order insert moved the rows from `first_moved` on up by one position: the listed positions follow
__global__ void renumber_positions_kernel(uint32_t* list, uint32_t max_items, uint32_t first_moved) {
    const uint32_t n = min(list[0], max_items);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (list[1 + i] >= first_moved) list[1 + i] += 1;
}

// Per query: gather the kept approximate scores of all its CTAs, sort, derive the certified candidate set.
struct TcCollectParams {
    TcParams t;
    uint32_t k, cap;            // cap: power of two >= G*kp, candidate slots per query
    uint32_t dims;
    const float* qnorm;         // [nq] canonical |q|^2
    const uint32_t* x2max_bits; // max row norm
    uint64_t n_live;
    uint32_t* pair_q;           // [nq*cap] query index of each candidate slot
    int64_t* pair_pos;          // [nq*cap] row position or -1
    uint8_t* fallback;          // [nq] 1 = bound not certified, use the exact scan
    const uint32_t* unsafe;     // [1 + TC_MAX_UNSAFE] rows that bypass the contraction
};
__global__ void __launch_bounds__(512) tc_collect_kernel(const TcCollectParams c) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* keys = (uint64_t*)smem;  // (order_bits(v) << 32) | slot
    const uint32_t q = blockIdx.x, qt = q / TC_M, e = q % TC_M;
    const TcParams& p = c.t;
    __shared__ uint32_t s_bad, s_total;
    if (threadIdx.x == 0) { s_bad = 0; s_total = 0; }
    __syncthreads();
    const uint32_t total_slots = p.G * p.kp;
    for (uint32_t j = threadIdx.x; j < c.cap; j += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (j < total_slots) {
            const uint32_t gg = j / p.kp, i = j - gg * p.kp;
            const size_t cta = (size_t)gg * p.QT + qt;
            if (i < p.cand_cnt[cta * TC_M + e]) {
                key = ((uint64_t)order_bits(p.cand_v[(cta * TC_M + e) * p.kp + i]) << 32) | j;
                atomicAdd(&s_total, 1u);
            }
        }
        keys[j] = key;
    }
    __syncthreads();
    block_bitonic_sort(keys, c.cap);
    const uint32_t total = s_total;
    // error bound of the approximate score (worst case over D fp32 accumulations of 3xTF32 partial products)
    const float q2 = c.qnorm[q], x2max = __uint_as_float(*c.x2max_bits);
    // (D+32) * 2^-21 covers the fp32 accumulation of the partial products; a single TF32 pass adds the truncation of both
    // operands to 10 mantissa bits: |q.x - tf32(q).tf32(x)| <= (2^-10 + 2^-10) |q||x| by Cauchy-Schwarz (+1 % slack)
    const float unit = (float)(c.dims + 32) * 4.76837158e-7f + (p.terms == 1 ? 1.97265625e-3f : 0.f);
    const float eps = p.cosine ? unit * sqrtf(q2) : unit * (q2 + x2max);
    float bound = __int_as_float(0x7F800000);
    if (total >= c.k) bound = order_bits_inv((uint32_t)(keys[c.k - 1] >> 32)) + 2.f * eps;
    // every full per-CTA list must have dropped only scores above the bound
    for (uint32_t gg = threadIdx.x; gg < p.G; gg += blockDim.x) {
        const size_t cta = (size_t)gg * p.QT + qt;
        if (p.cand_cnt[cta * TC_M + e] == p.kp && !(p.cand_tau[cta * TC_M + e] > bound)) atomicExch(&s_bad, 1u);
    }
    if (threadIdx.x == 0 && total < c.k && (uint64_t)total < c.n_live) s_bad = 1;  // non-finite scores hid rows
    if (threadIdx.x == 0 && !tc_norm_safe(q2)) s_bad = 1;                           // the query itself is unsafe
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < c.cap; j += blockDim.x) {
        const uint64_t key = keys[j];
        int64_t pos = -1;
        if (key != KEY_NONE && !(order_bits_inv((uint32_t)(key >> 32)) > bound)) {
            const uint32_t slot = (uint32_t)key, gg = slot / p.kp, i = slot - gg * p.kp;
            const size_t cta = (size_t)gg * p.QT + qt;
            pos = (int64_t)p.cand_r[(cta * TC_M + e) * p.kp + i];
        }
        // the last TC_MAX_UNSAFE slots are always free (cap >= G*kp + TC_MAX_UNSAFE): unsafe rows go there
        const uint32_t n_unsafe = min(c.unsafe[0], TC_MAX_UNSAFE);
        if (j >= c.cap - n_unsafe) {
            const uint32_t row = c.unsafe[1 + (c.cap - 1 - j)];
            pos = (p.skip && p.skip[row]) ? -1 : (int64_t)row;
        }
        c.pair_q[(size_t)q * c.cap + j] = q;
        c.pair_pos[(size_t)q * c.cap + j] = pos;
    }
    if (threadIdx.x == 0) c.fallback[q] = (uint8_t)s_bad;
}

// exact distances of the candidate pairs -> ranking keys (KEY_NONE for empty slots)
__global__ void tc_keys_kernel(const float* dist, const int64_t* pos, uint64_t n, uint64_t* keys) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        keys[i] = pos[i] >= 0 ? make_key(dist[i], (uint32_t)pos[i]) : KEY_NONE;
}

}  // namespace vg

// =====================================================================================================
// K3 (batched): int8 L2 on the tensor cores — tcgen05 kind::i8, exact int32 accumulation.
//   ||q - x||^2 = |q|^2 + |x|^2 - 2 q.x  is an exact integer, so the epilogue produces the FINAL ranking keys
//   ((float)sqrt((double)s), row) itself — no re-rank, no error bound.  Same pipeline as tc_scan_kernel minus
//   the lo-split: warp 0 TMA producer, warp 1 MMA issuer, warps 4-7 and 8-11 epilogue (one 128-column half of every
//   tile each); 4 smem stages of 48 KB.  Each epilogue thread (= one query, one column half) appends the keys that beat
//   its bound to a private buffer; full buffers are compacted by the whole warp (register bitonic sort) and the final
//   k best land in the [query][part][k] layout the final merge reads.  A first launch over a sampled prefix of the slab
//   provides every query's admission bound for the main launch (TciParams::tau_init).
// =====================================================================================================
namespace vg {

static constexpr uint32_t TCI_STAGES = 4;
static constexpr uint32_t TCI_EPI_WARPS = 8;             // two groups of four: each ranks one 128-column half of every tile
static constexpr uint32_t TCI_HALF = TC_N / 2;
static constexpr uint32_t TCI_THREADS = (4 + TCI_EPI_WARPS) * 32;
static constexpr int TCI_INEL = 0x40000000;  // |x|^2 stand-in of a row that must not be returned (valid values < 2^30 for dims <= 16384)
static constexpr uint32_t TCI_STAGE_BYTES = TCI_A_BYTES + TCI_B_BYTES;  // 48 KB
// kind::i8 instruction descriptor: D=S32, A=B=S8, both K-major, N=256, M=128
static constexpr uint32_t TCI_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

struct TciParams {
    uint64_t n_rows;
    uint32_t nq, nk, k;    // nk = ceil(row bytes / 128)
    uint32_t QT, G;
    const int* norms;      // [rows] exact |x|^2
    const int* qnorms;     // [nq] exact |q|^2
    const uint8_t* skip;
    uint64_t* out_keys;    // [nq][2G][k]   final per-(query, CTA, column half) results, KEY_NONE padded
    uint64_t* buf_keys;    // [nq][2G][cap] per-thread append buffers (cap: power of two >= k + 64)
    uint32_t cap;
    uint32_t* lockstep;    // see TcParams
    uint32_t lock_slack;
    uint64_t tile_begin, tile_end;  // 256-row tiles this launch scans
    uint32_t part_base, parts_total; // this launch writes parts part_base .. part_base + 2G - 1 of out_keys[q][parts_total][k]
    const uint64_t* tau_init;        // [nq] or NULL: a key every result must beat (k-th best of an earlier launch), KEY_NONE = none
    uint32_t debug;        // timing experiments (results invalid): 1 = no epilogue work, 2 = common path only, 4 = no norm loads
};

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TCI_IDESC), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(TCI_THREADS, 1)
tci8_scan_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapX, const TciParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* s_colB = (int*)(smem + TCI_STAGES * TCI_STAGE_BYTES);      // [8 epilogue warps][128] |x|^2 of the warp's half tile
    uint64_t* s_bar = (uint64_t*)(s_colB + TCI_EPI_WARPS * TCI_HALF);
    uint32_t* s_tmem = (uint32_t*)(s_bar + 16);
    uint64_t* s_scratch = s_bar + 32;                                // [8 epilogue warps] staging of half a chunk (rare path)
    const size_t scratch_bytes = 2048;
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + 4), bar_tfull = smem_u32(s_bar + 8),
                   bar_tempty = smem_u32(s_bar + 10);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < TCI_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, TCI_EPI_WARPS);
        }
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t qt = blockIdx.x % p.QT, g = blockIdx.x / p.QT;
    const uint64_t n_xt = p.tile_end - p.tile_begin;
    const uint32_t my_tiles = g < n_xt ? (uint32_t)((n_xt - g + p.G - 1) / p.G) : 0u;
    const uint64_t tile0 = p.tile_begin + g;

    if (warp == 0) {
        uint32_t it = 0;
        bool lock = p.lockstep != nullptr;
        for (uint32_t ti = 0; ti < my_tiles; ++ti) {
            if (lock) lock = tile_lockstep(p.lockstep + (size_t)g * 32, qt, p.QT, ti, p.lock_slack, lane);
            if (lane == 0) {
                const int row0 = (int)((tile0 + (uint64_t)ti * p.G) * TC_N);
                for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                    const uint32_t s = it % TCI_STAGES, ph = (it / TCI_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t base = smem_u32(smem + s * TCI_STAGE_BYTES);
                    mbar_expect_tx(bar_full + 8 * s, TCI_STAGE_BYTES);
                    tma_load_2d(base, &mapQ, (int)(kc * 128), (int)(qt * TC_M), bar_full + 8 * s);
                    tma_load_2d(base + TCI_A_BYTES, &mapX, (int)(kc * 128), row0, bar_full + 8 * s);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t ti = 0; ti < my_tiles; ++ti) {
                const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_N;
                for (uint32_t kc = 0; kc < p.nk; ++kc, ++it) {
                    const uint32_t s = it % TCI_STAGES, ph = (it / TCI_STAGES) & 1;
                    const uint32_t a = smem_u32(smem + s * TCI_STAGE_BYTES), b = a + TCI_A_BYTES;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)  // 4 x K=32 int8 per 128-byte chunk
                        umma_i8(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), (kc | k) != 0);
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // two epilogue groups: warps 4-7 rank columns 0..127 of every tile, warps 8-11 columns 128..255; warp w may only
        // touch TMEM lanes 32 (w % 4) .. +31, so both groups see all 128 queries
        const uint32_t ew = (uint32_t)(warp - 4);          // 0..7
        const uint32_t half = ew >> 2;                     // column half of the tile
        const uint32_t lane_base = (ew & 3) * 32;
        const uint32_t e = lane_base + lane;               // TMEM lane == query within the tile
        const uint32_t col0 = half * TCI_HALF;
        const uint32_t lpart = g * 2 + half;               // this thread's partial list of its query within this launch
        const uint32_t q = qt * TC_M + e;
        const bool q_ok = q < p.nq;
        // per-thread append buffer: keys that beat the (lazily refreshed) threshold are simply appended; when a
        // buffer is nearly full the warp compacts it cooperatively and the thread's threshold tightens
        uint64_t* buf = p.buf_keys + ((size_t)(q_ok ? q : 0) * (p.G * 2) + lpart) * p.cap;
        int* stage = (int*)((uint8_t*)s_scratch + (size_t)ew * scratch_bytes);  // [16 columns][32 lanes]
        int* my_colB = s_colB + ew * TCI_HALF; // this warp's copy of its half tile's |x|^2 (TCI_INEL = row not eligible)
        const int a2 = q_ok ? p.qnorms[q] : 0;
        uint32_t cnt = 0;
        uint64_t tau_key = KEY_NONE;   // k-th best key as of the last compaction
        int tau_s = 0x7FFFFFFE;        // every s above this is certainly not better than tau_key
        if (q_ok && p.tau_init && p.tau_init[q] != KEY_NONE) {  // k better rows are already known from a sample of the slab
            tau_key = p.tau_init[q];
            tau_s = tci_tau_s(tau_key);
        }
        int tau_a = q_ok ? tau_s - a2 : (int)0x80000000;  // same bound on t = |x|^2 - 2 q.x; nothing passes for a padding query
        // |x|^2 of the next tile is fetched one tile ahead (registers), so the loads never sit in front of the MMA wait
        int nb[TCI_HALF / 32];
        auto fetch_norms = [&](uint32_t ti) {
            const uint64_t r0 = (tile0 + (uint64_t)ti * p.G) * TC_N + col0;
#pragma unroll
            for (int i = 0; i < (int)TCI_HALF / 32; ++i) {
                const uint64_t row = r0 + (uint64_t)(i * 32 + lane);
                const bool ok = row < p.n_rows && !(p.skip && p.skip[row]);
                nb[i] = (p.debug & 4) ? 1 : ok ? p.norms[row] : TCI_INEL;
            }
        };
        if (my_tiles) fetch_norms(0);
        for (uint32_t ti = 0; ti < my_tiles; ++ti) {
            const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
            const uint64_t row0 = (tile0 + (uint64_t)ti * p.G) * TC_N + col0;
#pragma unroll
            for (int i = 0; i < (int)TCI_HALF / 32; ++i) my_colB[i * 32 + lane] = nb[i];
            __syncwarp();
            if (ti + 1 < my_tiles) fetch_norms(ti + 1);
            mbar_wait(bar_tfull + 8 * acc, aph);
            tc_fence_after();
            // accumulators come out of TMEM 32 columns at a time, the next chunk's load in flight while this one is ranked
            auto process = [&](uint32_t (&v)[32], const uint32_t c) {
                // common case, branch-free: t = |x|^2 - 2 q.x for the 32 columns and a pass mask against tau_a
                // (s = |q|^2 + t is the exact squared distance; ineligible rows carry a huge |x|^2)
                const int4* cb4 = (const int4*)(my_colB + c * 32);
                int tv[32];
                uint32_t mask = 0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int4 b = cb4[j4];
                    tv[4 * j4 + 0] = b.x - 2 * (int)v[4 * j4 + 0];
                    tv[4 * j4 + 1] = b.y - 2 * (int)v[4 * j4 + 1];
                    tv[4 * j4 + 2] = b.z - 2 * (int)v[4 * j4 + 2];
                    tv[4 * j4 + 3] = b.w - 2 * (int)v[4 * j4 + 3];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) mask |= (tv[j] <= tau_a ? 1u : 0u) << j;
                if (p.debug & 2) mask = mask == 0x12345678u ? 1u : 0u;  // timing experiments: common path only
                // rare path (warp-uniform entry): the chunk is staged in shared memory so that the few passing columns
                // can be picked by index in a compact loop; the exact key is (float)sqrt((double)s)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t m16 = (mask >> (16 * hh)) & 0xFFFFu;
                    if (__any_sync(0xffffffffu, m16 != 0)) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) stage[j * 32 + lane] = tv[16 * hh + j];
                        while (m16) {
                            const int j = __ffs(m16) - 1;
                            m16 &= m16 - 1;
                            const int s = a2 + stage[j * 32 + lane];
                            const uint32_t col = c * 32 + 16 * hh + (uint32_t)j;
                            if (my_colB[col] < TCI_INEL) {
                                // src/distance/scalar.rs:65: f64 sqrt, then cast — the same final value as the scan
                                const float d = exact_sqrt_int(s);
                                const uint64_t key = make_key(d, (uint32_t)(row0 + col));
                                if (key < tau_key) buf[cnt++] = key;
                            }
                        }
                        __syncwarp();
                    }
                }
                // a buffer that could overflow during the next 32 columns is compacted now (warp-uniform loop)
                unsigned need = __ballot_sync(0xffffffffu, cnt + 32 > p.cap);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const uint64_t bp = shfl_u64((uint64_t)(uintptr_t)buf, src);
                    const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
                    const uint64_t kth = warp_compact((uint64_t*)(uintptr_t)bp, bc, p.k, nullptr, lane);
                    if (lane == src) {
                        cnt = min(bc, p.k);
                        tau_key = kth;
                        if (kth != KEY_NONE) {
                            tau_s = tci_tau_s(kth);
                            tau_a = tau_s - a2;
                        }
                    }
                }
                        };
            if (!(p.debug & 1)) {
                uint32_t va[32], vb[32];
                const uint32_t tbase = tmem_base + (lane_base << 16) + acc * TC_N + col0;
                TCI_LD32(va, tbase);
#pragma unroll 1
                for (uint32_t c = 0; c < TCI_HALF / 32; c += 2) {
                    TCI_WAIT32(va);
                    TCI_LD32(vb, tbase + (c + 1) * 32);
                    process(va, c);
                    TCI_WAIT32(vb);
                    if (c + 2 < TCI_HALF / 32) TCI_LD32(va, tbase + (c + 2) * 32);
                    process(vb, c + 1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        // final compaction of every lane's buffer -> the k best keys of this (query, CTA), KEY_NONE padded
        for (int src = 0; src < 32; ++src) {
            const uint64_t bp = shfl_u64((uint64_t)(uintptr_t)buf, src);
            const uint32_t bc = __shfl_sync(0xffffffffu, cnt, src);
            const uint32_t qs = __shfl_sync(0xffffffffu, q, src);
            if (qs >= p.nq) continue;  // warp-uniform
            warp_compact((uint64_t*)(uintptr_t)bp, bc, p.k, p.out_keys + ((size_t)qs * p.parts_total + p.part_base + lpart) * p.k, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// k-th smallest key per query over the partial lists of a sample launch -> the admission bound of the main launch
__global__ void __launch_bounds__(1024) tci8_tau_kernel(const uint64_t* keys, uint32_t n_keys, uint64_t q_stride, uint32_t k, uint32_t np2,
                                                        uint64_t* tau) {
    extern __shared__ __align__(16) uint8_t tau_smem[];
    uint64_t* sk = (uint64_t*)tau_smem;
    const uint64_t* src = keys + (uint64_t)blockIdx.x * q_stride;
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) sk[i] = i < n_keys ? src[i] : KEY_NONE;
    __syncthreads();
    block_bitonic_sort(sk, np2);
    if (threadIdx.x == 0) tau[blockIdx.x] = sk[k - 1];
}

// exact |row|^2 of int8 rows (int32)
__global__ void __launch_bounds__(256) row_norms_i8_kernel(const uint8_t* base, uint32_t stride, uint32_t units, uint64_t n, int* norms) {
    const int g = threadIdx.x & 3;
    const uint64_t n_iter = (n + 63) / 64;
    for (uint64_t itn = blockIdx.x; itn < n_iter; itn += gridDim.x) {
        const uint64_t row = itn * 64 + (threadIdx.x >> 2);
        const uint4* r = (const uint4*)(base + (row < n ? row : 0) * (uint64_t)stride);
        const float v = query_const(r, row < n ? units : 0, g, 1);
        if (row < n && g == 0) norms[row] = __float_as_int(v);
    }
}

}  // namespace vg

// =====================================================================================================
// f32 L1 scan with swizzled TMA boxes.  L1 must add |a_i - b_i| strictly left to right
// (src/distance/scalar.rs:31-35), i.e. one thread per row, which makes plain row-major shared memory
// 8-way bank-conflicted.  A 2-D tensor-map load with SWIZZLE_128B stores unit u of row r at u ^ (r & 7), so
// the 8 lanes of a quarter-warp (8 consecutive rows) hit 8 different 16-byte units: conflict-free, one TMA
// instruction per 32-row x 128-byte box (4 KB) instead of 32 padded per-row copies.
// Warp w owns tiles w, w+C, ... of 32 rows and a private ring of D boxes which it refills itself.
// =====================================================================================================
namespace vg {

struct L1Params {
    const uint8_t* skip;
    const uint8_t* queries;   // nq_total rows of row_stride bytes (device)
    uint64_t* out_keys;       // [nq_total][gridDim.x][k]
    uint64_t n_rows;
    uint32_t nq_total, k, row_stride, q_stride;  // q_stride: row_stride rounded up to 128 (shared-memory query pitch)
    uint32_t n_chunks;        // ceil(row_stride / 128)
    uint32_t ring;            // D
    uint32_t n_warps;         // C
    uint32_t list_stride;
};

template <int QB>
__global__ void __launch_bounds__(512, 1) scan_l1_tma_kernel(const __grid_constant__ CUtensorMap mapX, const L1Params p) {
    using T = F32L1<QB>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t C = p.n_warps, D = p.ring;
    // layout: [C*D boxes of 4 KB][queries QB*q_stride][hdr C*QB][lists QB*list_stride][barriers C*D]
    uint8_t* s_box = smem;
    uint8_t* s_query = s_box + (size_t)C * D * 4096;
    ListHdr* s_hdr = (ListHdr*)(s_query + (size_t)QB * p.q_stride);
    uint64_t* s_list = (uint64_t*)(s_hdr + C * QB);
    uint64_t* s_bar = s_list + (size_t)QB * p.list_stride;
    const uint32_t q0 = blockIdx.y * QB;
    const uint32_t nq_here = min((uint32_t)QB, p.nq_total - q0);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < C * D; ++s) mbar_init(smem_u32(s_bar + s), 1);
        mbar_fence_init();
    }
    {
        const uint32_t qunits = p.q_stride / 16, runits = p.row_stride / 16;
        uint4* sq = (uint4*)s_query;
        for (uint32_t i = threadIdx.x; i < QB * qunits; i += blockDim.x) {
            const uint32_t qi = i / qunits, u = i - qi * qunits;
            sq[i] = (qi < nq_here && u < runits) ? ((const uint4*)(p.queries + (size_t)(q0 + qi) * p.row_stride))[u] : make_uint4(0, 0, 0, 0);
        }
        for (uint32_t i = threadIdx.x; i < C * QB; i += blockDim.x) {
            s_hdr[i].tau = KEY_NONE;
            s_hdr[i].maxpos = 0;
            s_hdr[i].cnt = 0;
        }
    }
    __syncthreads();
    const uint64_t n_tiles = (p.n_rows + 31) / 32;
    const uint64_t first_tile = blockIdx.x + (uint64_t)warp * gridDim.x, tile_step = (uint64_t)C * gridDim.x;
    const uint32_t my_tiles = first_tile < n_tiles ? (uint32_t)((n_tiles - first_tile + tile_step - 1) / tile_step) : 0u;
    const uint32_t my_iters = my_tiles * p.n_chunks;
    const uint32_t my_bar = smem_u32(s_bar + warp * D);
    const uint32_t my_box = smem_u32(s_box + (size_t)warp * D * 4096);
    auto issue = [&](uint32_t lit) {
        if (lit >= my_iters) return;
        fence_proxy_async();
        if (lane == 0) {
            const uint32_t jl = lit / p.n_chunks, c = lit - jl * p.n_chunks, slot = lit % D;
            mbar_expect_tx(my_bar + 8 * slot, 4096);
            tma_load_2d(my_box + slot * 4096, &mapX, (int)(c * 128), (int)((first_tile + (uint64_t)jl * tile_step) * 32), my_bar + 8 * slot);
        }
        __syncwarp();
    };
    for (uint32_t lit = 0; lit < D; ++lit) issue(lit);
    uint64_t* my_list = s_list + (size_t)warp * p.k;
    ListHdr* my_hdr = s_hdr + warp * QB;
    const uint32_t q_base = smem_u32(s_query);
    const uint32_t sw = (uint32_t)(lane & 7);
    for (uint32_t jl = 0; jl < my_tiles; ++jl) {
        const uint64_t row = (first_tile + (uint64_t)jl * tile_step) * 32 + lane;
        typename T::Acc acc;
        T::init(acc);
        for (uint32_t c = 0; c < p.n_chunks; ++c) {
            const uint32_t lit = jl * p.n_chunks + c, slot = lit % D;
            mbar_wait(my_bar + 8 * slot, (lit / D) & 1);
            const uint32_t xb = my_box + slot * 4096 + lane * 128, qb = q_base + c * 128;
            uint4 xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = lds128(xb + (((uint32_t)u ^ sw) << 4));  // de-swizzle: conflict-free
#pragma unroll
            for (int u = 0; u < 8; ++u) {  // strict left-to-right order over the row
                uint4 qv[QB];
#pragma unroll
                for (int i = 0; i < QB; ++i) qv[i] = lds128(qb + i * p.q_stride + u * 16);
                T::step(acc, xv[u], qv);
            }
            __syncwarp();
            issue(lit + D);
        }
        bool live = row < p.n_rows;
        if (live && p.skip != nullptr) live = p.skip[row] == 0;
#pragma unroll
        for (int i = 0; i < QB; ++i) {
            const float d = T::finish(acc, i, nullptr);
            list_offer(my_list + (size_t)i * p.list_stride, my_hdr + i, p.k, make_key(d, (uint32_t)row), live, lane);
        }
    }
    __syncthreads();
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

}  // namespace vg
