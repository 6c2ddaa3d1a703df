// SWIZZLE_64B variant (16-float k-chunks). Minimal tcgen05 (kind::tf32) + TMA (2-D tensor map, SWIZZLE_128B) + TMEM bring-up test.
// One CTA computes C[128 x 256] = A[128 x K] * B[256 x K]^T with K = 32*nk and compares with the CPU.
// Validates the descriptor encodings used by the batched-query kernel before it is built on top.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_test tools/umma_test.cu
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B, rows of 128 bytes: 8-row groups are 1024 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)0 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                 // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                          // layout type: SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // F32 acc, TF32 x TF32, K-major, N=256, M=128

__global__ void __launch_bounds__(128, 1) umma_test(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out, int nk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 8192;
    uint64_t* bars = (uint64_t*)(smem + 24576);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tma_bar = smem_u32(bars), mma_bar = smem_u32(bars + 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(tma_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        for (int kc = 0; kc < nk; ++kc) {
            mbar_expect_tx(tma_bar, 24576);
            tma_load_2d(smem_u32(sA), &mapA, kc * 16, 0, tma_bar);
            tma_load_2d(smem_u32(sB), &mapB, kc * 16, 0, tma_bar);
            mbar_wait(tma_bar, kc & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < 2; ++k)
                umma_tf32(tmem_base, make_desc(smem_u32(sA) + k * 32), make_desc(smem_u32(sB) + k * 32), IDESC, (kc | k) != 0);
            umma_commit(mma_bar);
            mbar_wait(mma_bar, kc & 1);
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w..32w+31; thread = row, 32 columns per load
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
        for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rn_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0xFFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
    const int nk = argc > 1 ? atoi(argv[1]) : 4, K = 16 * nk, ldk = K + 8;  // padded pitch: exercises the global stride
    std::vector<float> A((size_t)128 * ldk), B((size_t)256 * ldk);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, 128 * 256 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    auto make_map = [&](CUtensorMap* m, float* base, int rows, int box_rows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)ldk * 4};
        cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    CUtensorMap mA, mB;
    make_map(&mA, dA, 128, 128);
    make_map(&mB, dB, 256, 256);
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    umma_test<<<1, 128, 24576 + 64>>>(mA, mB, dC, nk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> C(128 * 256);
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double max_t = 0, max_r = 0, max_f = 0, ref_mag = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 256; ++j) {
            double st = 0, sr = 0, sf = 0;
            for (int k = 0; k < K; ++k) {
                float a = A[(size_t)i * ldk + k], b = B[(size_t)j * ldk + k];
                st += (double)trunc_tf32(a) * trunc_tf32(b);
                sr += (double)rn_tf32(a) * rn_tf32(b);
                sf += (double)a * b;
            }
            double c = C[i * 256 + j];
            max_t = fmax(max_t, fabs(c - st)); max_r = fmax(max_r, fabs(c - sr)); max_f = fmax(max_f, fabs(c - sf));
            ref_mag = fmax(ref_mag, fabs(sf));
        }
    printf("K=%d  max|C-ref|: vs trunc-tf32 inputs %.3e, vs rn-tf32 inputs %.3e, vs full fp32 %.3e  (max |ref| %.3f)\n", K, max_t, max_r, max_f, ref_mag);
    printf("C[0][0..3] = %f %f %f %f\n", C[0], C[1], C[2], C[3]);
    printf(max_t < 1e-4 || max_r < 1e-4 ? "UMMA_TEST PASS\n" : "UMMA_TEST FAIL\n");
    return 0;
}
