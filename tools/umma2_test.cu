// tcgen05 cta_group::2 bring-up: a CTA PAIR computes C[256 x 256] = A[256 x K] * B[256 x K]^T (kind::tf32, M=256).
// CTA r holds rows r*128.. of A and rows r*128.. of B (= columns r*128.. of C) in ITS shared memory; both CTAs issue their
// own TMA loads, which complete on the LEADER's mbarrier (cp.async.bulk.tensor...cta_group::2, peer bit of the barrier address
// cleared); the leader issues tcgen05.mma.cta_group::2 and commits with a multicast arrive on both CTAs' barriers; each CTA
// reads its 128 accumulator rows (all 256 columns) from its own TMEM.  Validates the encodings before tc_scan_kernel uses them.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma2_test tools/umma2_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; spins < (1u << 26); ++spins)
        if (mbar_test(bar, parity)) return;
    __trap();
}
static constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);  // F32 acc, TF32, K-major, N=256, M=256

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_test(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out, int nk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;            // 128 rows x 128 B: this CTA's rows of A
    uint8_t* sB = smem + 16384;    // 128 rows x 128 B: this CTA's rows of B
    uint64_t* bars = (uint64_t*)(smem + 32768);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t full_bar = smem_u32(bars), mma_bar = smem_u32(bars + 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(full_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    cluster_sync_all();  // both CTAs' barriers exist before any remote completion can land
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        for (int kc = 0; kc < nk; ++kc) {
            if (rank == 0) mbar_expect_tx(full_bar, 65536);  // A and B halves of BOTH CTAs land on the leader's barrier
            tma_load_2d_pair(smem_u32(sA), &mapA, kc * 32, (int)rank * 128, full_bar);
            tma_load_2d_pair(smem_u32(sB), &mapB, kc * 32, (int)rank * 128, full_bar);
            if (rank == 0) {
                mbar_wait(full_bar, kc & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int k = 0; k < 4; ++k)
                    umma2_tf32(tmem_base, make_desc(smem_u32(sA) + k * 32), make_desc(smem_u32(sB) + k * 32), IDESC2, (kc | k) != 0);
                umma2_commit_mc(mma_bar, 0b11);  // both CTAs may overwrite their stage / read their accumulator
            }
            mbar_wait(mma_bar, kc & 1);
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)(rank * 128 + warp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // the peer may still be reading operands / TMEM of the pair
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
    const int nk = argc > 1 ? atoi(argv[1]) : 4, K = 32 * nk, ldk = K + 8;
    std::vector<float> A((size_t)256 * ldk), B((size_t)256 * ldk);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, 256 * 256 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, 256 * 256 * 4);
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    auto make_map = [&](CUtensorMap* m, float* base, int rows, int box_rows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)ldk * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    CUtensorMap mA, mB;
    make_map(&mA, dA, 256, 128);
    make_map(&mB, dB, 256, 128);
    cudaFuncSetAttribute(umma2_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    umma2_test<<<2, 128, 32768 + 1024 + 64>>>(mA, mB, dC, nk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> C(256 * 256);
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double max_t = 0, ref_mag = 0;
    int bad_i = -1, bad_j = -1;
    for (int i = 0; i < 256; ++i)
        for (int j = 0; j < 256; ++j) {
            double st = 0;
            for (int k = 0; k < K; ++k) st += (double)trunc_tf32(A[(size_t)i * ldk + k]) * trunc_tf32(B[(size_t)j * ldk + k]);
            const double err = fabs(C[i * 256 + j] - st);
            if (err > max_t) { max_t = err; bad_i = i; bad_j = j; }
            ref_mag = fmax(ref_mag, fabs(st));
        }
    printf("K=%d  max|C-ref| vs trunc-tf32 inputs %.3e at (%d,%d)  (max |ref| %.3f)\n", K, max_t, bad_i, bad_j, ref_mag);
    printf("C[0][0..1] = %f %f  C[0][128] = %f  C[128][0] = %f  C[255][255] = %f\n", C[0], C[1], C[128], C[128 * 256], C[255 * 256 + 255]);
    printf(max_t < 1e-4 ? "UMMA2_TEST PASS\n" : "UMMA2_TEST FAIL\n");
    return 0;
}
