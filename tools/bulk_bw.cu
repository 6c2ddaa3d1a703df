// Microbenchmark: HBM->shared streaming with cp.async.bulk rings, no compute.
// Finds the best (warps, ring depth, stage bytes) for the scan pipeline.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__global__ void stream_kernel(const uint8_t* src, uint64_t total_bytes, uint32_t stage_bytes, uint32_t D, uint32_t fence, uint32_t touch, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t C = blockDim.x >> 5;
    uint64_t* bars = (uint64_t*)(smem + (size_t)C * D * stage_bytes);
    if (threadIdx.x == 0) { for (uint32_t s = 0; s < C * D; ++s) mbar_init(smem_u32(bars + s), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint64_t n_tiles = total_bytes / stage_bytes;
    const uint64_t first = blockIdx.x + (uint64_t)warp * gridDim.x, step = (uint64_t)C * gridDim.x;
    const uint32_t mine = first < n_tiles ? (uint32_t)((n_tiles - first + step - 1) / step) : 0;
    uint8_t* my_stage = smem + (size_t)warp * D * stage_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * D);
    auto issue = [&](uint32_t it) {
        if (it >= mine) return;
        if (fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) {
            const uint32_t s = it % D;
            mbar_expect_tx(my_bar + 8 * s, stage_bytes);
            bulk_g2s(smem_u32(my_stage + (size_t)s * stage_bytes), src + (first + (uint64_t)it * step) * stage_bytes, stage_bytes, my_bar + 8 * s);
        }
        __syncwarp();
    };
    for (uint32_t it = 0; it < D; ++it) issue(it);
    uint32_t acc = 0;
    for (uint32_t it = 0; it < mine; ++it) {
        const uint32_t s = it % D;
        mbar_wait(my_bar + 8 * s, (it / D) & 1);
        if (touch) {  // read every 16-byte unit once (what a scan must do at least)
            const uint4* p = (const uint4*)(my_stage + (size_t)s * stage_bytes);
            for (uint32_t u = lane; u < stage_bytes / 16; u += 32) { uint4 v = p[u]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        }
        __syncwarp();
        issue(it + D);
    }
    if (acc == 0x12345678) sink[0] = acc;
}
int main(int argc, char** argv) {
    const uint64_t total = 8ull << 30;
    uint8_t* d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
    uint32_t* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int warps[] = {2, 4, 8, 16};
    const int kbs[] = {4, 8, 12, 16, 24, 32, 48};
    for (int touch = 0; touch < 2; ++touch)
    for (int fence = 0; fence < 2; ++fence)
    for (int w : warps) for (int kb : kbs) for (int D = 2; D <= 8; D *= 2) {
        size_t smem = (size_t)w * D * kb * 1024 + w * D * 8 + 128;
        if (smem > 227 * 1024) continue;
        if (fence == 0 && touch == 1) continue;
        float best = 1e9;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a);
            stream_kernel<<<148, 32 * w, smem>>>(d, total, kb * 1024, D, fence, touch, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (r > 0 && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("touch=%d fence=%d warps=%2d stage=%2dKB D=%d inflight=%3dKB : %.3f ms %.0f GB/s\n", touch, fence, w, kb, D, w * D * kb, best, total / best / 1e6);
        fflush(stdout);
    }
    return 0;
}
