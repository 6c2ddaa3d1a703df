// xchg.cuh — the exchange step of a sharded exact KNN, fused with the final merge over NVLink peer memory.
//
// The reference is one process on one CPU (src/vtab.rs:2286-2305 calls brute_force_search inline); sharding a slab by
// contiguous rowid range over the GPUs of a box is this library's own addition, so there is no reference line to
// restate here — only the contract: the merged result must equal a scan of one slab holding all rows.
//
// Every rank (one GPU) owns a GATHER BUFFER in its own HBM, mapped into every peer (same process:
// cudaDeviceEnablePeerAccess; other processes: CUDA IPC).  For exchange number `epoch`:
//
//   xpush_kernel        after the local top-k is final, each rank writes its packed result — k x (i64 rowid, f32
//                       distance) + a count per query = 12 bytes per entry — STRAIGHT INTO EVERY PEER'S buffer with
//                       plain stores over NVLink, fences at system scope, and the last CTA publishes `epoch` in every
//                       peer's flag word (release).  No NCCL call, no pack/unpack kernels, no staging copy.
//   xwait_merge_kernel  one CTA per query: acquires the `world` flag words of ITS OWN buffer (local L2 polling, no
//                       traffic on the links), then merges the `world` lists.  Rank r holds rowids below rank r+1's, so
//                       (order_bits(d), source slot) orders exactly like (d, rowid): one u64 key per entry, one register
//                       sort by a single warp when world * k <= 256 (8 GPUs x k <= 32), a shared-memory bitonic sort above.
//
// Two buffer halves alternate by epoch parity.  A rank pushes epoch e+2 only after its own merge of e+1 has run, which
// needs every peer's push of e+1, which — in stream order on the peer — follows the peer's merge of e: so half e % 2 is
// never overwritten while a peer still reads it, provided every rank issues the exchanges in the same order on one stream
// (collective semantics, as with NCCL).  A peer that never arrives trips a clock64() timeout: the kernel raises an error
// word instead of spinning forever.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "xpush.cuh"

namespace vg {

__global__ void __launch_bounds__(128) xpush_kernel(const XPushParams p) {
    for (uint32_t q = blockIdx.x; q < p.nq; q += gridDim.x) xpush_query(p, q);
    xpush_publish(p, gridDim.x);
}

struct XWaitMergeParams {
    const XPeerTable* tab;
    XLayout lay;
    uint32_t nq, k, rank, epoch;
    uint32_t src_mask;         // ranks whose lists take part (all ranks of the world in practice)
    uint32_t np2;              // power of two >= world * k (shared-memory path)
    int64_t pad_rowid;
    int64_t* out_rowids;       // [nq][k]
    float* out_dists;
    uint32_t* out_counts;      // [nq] or nullptr
    uint32_t* err;             // set to 1 when a peer's flag did not arrive in time
    unsigned long long timeout_cycles;
};

// One CTA per query.  Warp 0 / lanes < world wait for the peers' flags; the merge follows.
__global__ void __launch_bounds__(256) xwait_merge_kernel(const XWaitMergeParams p) {
    extern __shared__ __align__(16) uint8_t xsm[];
    uint64_t* keys = (uint64_t*)xsm;
    __shared__ uint32_t s_ok;
    const uint32_t half = p.epoch & 1u, W = p.lay.world, q = blockIdx.x;
    uint8_t* mine = p.tab->base[p.rank];
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < W && ((p.src_mask >> threadIdx.x) & 1u)) {
        const uint32_t* f = x_flag(mine, p.lay, half, threadIdx.x);
        const long long t0 = clock64();
        // epochs only grow; a flag may already be ahead by a multiple of 2 only if this rank lagged a whole exchange,
        // which the protocol excludes — equality is the expected value, >= (wrap-safe) is accepted
        while ((int32_t)(ld_acquire_sys(f) - p.epoch) < 0) {
            if ((unsigned long long)(clock64() - t0) > p.timeout_cycles) {
                s_ok = 0;
                *(volatile uint32_t*)p.err = 1u;  // plain store: the word may live in mapped host memory
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!s_ok) {  // a peer is missing: no result (the host reports the error); leave padding
        for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
            p.out_rowids[(size_t)q * p.k + j] = p.pad_rowid;
            p.out_dists[(size_t)q * p.k + j] = __int_as_float(0x7F800000);
        }
        if (threadIdx.x == 0 && p.out_counts) p.out_counts[q] = 0;
        return;
    }
    const uint32_t n = W * p.k;
    auto load_key = [&](uint32_t j) -> uint64_t {
        if (j >= n) return KEY_NONE;
        const uint32_t src = j / p.k, i = j - src * p.k;
        if (!((p.src_mask >> src) & 1u)) return KEY_NONE;
        if (i >= x_counts(mine, p.lay, half, src)[q]) return KEY_NONE;
        const float d = x_dists(mine, p.lay, half, src)[(size_t)q * p.k + i];
        return ((uint64_t)order_bits(d) << 32) | j;
    };
    auto emit = [&](uint32_t slot, uint64_t key) -> uint32_t {
        const size_t o = (size_t)q * p.k + slot;
        if (key == KEY_NONE) {
            p.out_rowids[o] = p.pad_rowid;
            p.out_dists[o] = __int_as_float(0x7F800000);
            return 0;
        }
        const uint32_t j = (uint32_t)key, src = j / p.k, i = j - src * p.k;
        p.out_rowids[o] = x_rowids(mine, p.lay, half, src)[(size_t)q * p.k + i];
        p.out_dists[o] = x_dists(mine, p.lay, half, src)[(size_t)q * p.k + i];  // the stored bits (NaN payloads survive)
        return 1;
    };
    if (n <= 256) {
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        uint64_t v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = load_key((uint32_t)lane * 8 + r);
        warp_sort256(v, lane);
        uint32_t cnt = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t j = (uint32_t)lane * 8 + r;
            if (j < p.k) cnt += emit(j, v[r]);
        }
        if (p.out_counts) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
            if (lane == 0) p.out_counts[q] = cnt;
        }
        return;
    }
    for (uint32_t j = threadIdx.x; j < p.np2; j += blockDim.x) keys[j] = load_key(j);
    __syncthreads();
    block_bitonic_sort(keys, p.np2);
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) cnt += emit(j, keys[j]);
    if (cnt) atomicAdd(&total, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && p.out_counts) p.out_counts[q] = total;
}

}  // namespace vg
