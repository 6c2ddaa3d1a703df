// Measured tensor-pipe peaks for the two tcgen05 kinds the batched kernels use (kind::tf32 and kind::i8, M=128 x N=256,
// cta_group::1, both operands in shared memory, SWIZZLE_128B K-major): one CTA per SM, one thread issues back-to-back MMAs
// over resident operands (no loads, no epilogue), two accumulators alternating, one commit at the end.
// This is the denominator for the batched rooflines ("measure on the box").
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_peak tools/mma_peak.cu && tools/mma_peak
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, unsigned long long timeout_cycles, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;             // 128 rows x 128 B
    uint8_t* sB = smem + 16384;     // 256 rows x 128 B
    uint64_t* bars = (uint64_t*)(smem + 49152);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5;
    const uint32_t mma_bar = smem_u32(bars);
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        ((uint32_t*)smem)[i] = KIND == 0 ? (0x3F000000u | (h & 0x007FE000u)) : (h & 0x7F7F7F7Fu);  // tf32 values in [0.5,1) / small int8
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = KIND == 0 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24))
                                     : ((2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24));
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
                const int k = j & 3;  // the four 32-byte K-steps of a 128-byte swizzle row
                umma<KIND>(tmem_base + ((j >> 2) & 1) * 256, make_desc(smem_u32(sA) + k * 32), make_desc(smem_u32(sB) + k * 32), idesc, (it | j) >= 8);
            }
        }
        umma_commit(mma_bar);  // ONE commit: it completes when every MMA issued before it has
        const long long t0 = clock64();
        while (!mbar_try(mma_bar, 0)) {
            if ((unsigned long long)(clock64() - t0) > timeout_cycles) {
                *err = 1;
                break;
            }
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

extern int g_iters;
template <int KIND>
static void run(const char* name, double ops_per_mma, const char* unit) {
    int* d_err;
    cudaMalloc(&d_err, 4);
    cudaMemset(d_err, 0, 4);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(mma_peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 4; ++rep) {
        const int iters = rep == 0 ? 64 : g_iters;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
        mma_peak_kernel<KIND><<<sms, 128, 49152 + 1024 + 64>>>(iters, 4000000000ull, d_err);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        int err = 0;
        cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
        const double ops = (double)sms * iters * 32 * ops_per_mma;
        if (rep) printf("%s: %d SMs x %d x 32 MMAs (M=128 N=256) in %.3f ms = %.1f %s%s\n", name, sms, iters, ms, ops / (ms * 1e-3) / 1e12, unit, err ? "  [TIMEOUT]" : "");
    }
}

int g_iters = 4096;
int main(int argc, char** argv) {
    if (argc > 1) g_iters = atoi(argv[1]);  // e.g. 200000 for a ~0.5 s launch: shows the rate under the power cap
    run<0>("kind::tf32 K=8 ", 2.0 * 128 * 256 * 8, "TFLOP/s");
    run<1>("kind::i8   K=32", 2.0 * 128 * 256 * 32, "TOP/s");
    return 0;
}
