// xpush.cuh — push side of the sharded-KNN exchange (see xchg.cuh for the protocol): layout of a rank's gather buffer and
// the device functions that write a local top-k into the peers' buffers and publish the epoch flag.  Kept separate so that
// scan_kernel (kernels.cuh) can push from its fused tail.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vg {

static constexpr uint32_t XCHG_MAX_WORLD = 16;

struct XLayout {               // identical on every rank (checked at attach time)
    uint64_t cap_entries;      // nq * k of one exchange must fit
    uint32_t cap_q;            // queries per exchange
    uint32_t world;
    uint64_t region_bytes;     // [rowids cap_entries x 8][dists cap_entries x 4][counts cap_q x 4], 16-byte aligned
    uint64_t half_bytes;       // world regions
    uint64_t flags_off;        // 2 halves x world x u32 (written by peers), after the two halves
    uint64_t total_bytes;
};

__host__ __device__ inline XLayout xlayout(uint32_t world, uint32_t cap_q, uint64_t cap_entries) {
    XLayout l;
    l.world = world;
    l.cap_q = cap_q;
    l.cap_entries = cap_entries;
    l.region_bytes = (cap_entries * 12 + (uint64_t)cap_q * 4 + 15) & ~(uint64_t)15;
    l.half_bytes = l.region_bytes * world;
    l.flags_off = 2 * l.half_bytes;
    l.total_bytes = l.flags_off + 2 * (uint64_t)XCHG_MAX_WORLD * 4 + 256;
    return l;
}

struct XPeerTable {            // device-resident, filled once at attach
    uint8_t* base[XCHG_MAX_WORLD];   // base[p] = rank p's gather buffer as mapped on this device (own rank: local memory)
};

__device__ __forceinline__ int64_t* x_rowids(uint8_t* base, const XLayout& l, uint32_t half, uint32_t src) {
    return (int64_t*)(base + half * l.half_bytes + src * l.region_bytes);
}
__device__ __forceinline__ float* x_dists(uint8_t* base, const XLayout& l, uint32_t half, uint32_t src) {
    return (float*)(base + half * l.half_bytes + src * l.region_bytes + l.cap_entries * 8);
}
__device__ __forceinline__ uint32_t* x_counts(uint8_t* base, const XLayout& l, uint32_t half, uint32_t src) {
    return (uint32_t*)(base + half * l.half_bytes + src * l.region_bytes + l.cap_entries * 12);
}
__device__ __forceinline__ uint32_t* x_flag(uint8_t* base, const XLayout& l, uint32_t half, uint32_t src) {
    return (uint32_t*)(base + l.flags_off) + half * XCHG_MAX_WORLD + src;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct XPushParams {
    const XPeerTable* tab;
    XLayout lay;
    const int64_t* src_rowids;   // [nq][k] local top-k (padding: pad rowid / +inf)
    const float* src_dists;
    const uint32_t* src_counts;  // [nq] valid entries per query, or nullptr: derived (leading entries that are not padding)
    uint32_t nq, k, rank, epoch;
    uint32_t peer_mask;          // bit p set: push to rank p
    uint32_t* done;              // local counter for the last-CTA-publishes pattern (reset by the last CTA)
};

// Per query q handled by this CTA: copy k entries + the count into region[half][rank] of every peer in the mask.
__device__ __forceinline__ void xpush_query(const XPushParams& p, uint32_t q) {
    const uint32_t half = p.epoch & 1u;
    const int64_t* sr = p.src_rowids + (size_t)q * p.k;
    const float* sd = p.src_dists + (size_t)q * p.k;
    uint32_t cnt = p.k;
    if (p.src_counts) {
        cnt = p.src_counts[q];
    } else if (threadIdx.x == 0) {  // device-array API: padding is (INT64_MAX, +inf), always at the end
        cnt = 0;
        while (cnt < p.k && !(sr[cnt] == INT64_MAX && __float_as_uint(sd[cnt]) == 0x7F800000u)) ++cnt;
    }
    for (uint32_t pr = 0; pr < p.lay.world; ++pr) {
        if (!((p.peer_mask >> pr) & 1u)) continue;
        uint8_t* b = p.tab->base[pr];
        int64_t* dr = x_rowids(b, p.lay, half, p.rank) + (size_t)q * p.k;
        float* dd = x_dists(b, p.lay, half, p.rank) + (size_t)q * p.k;
        for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
            dr[j] = sr[j];
            dd[j] = sd[j];
        }
        if (threadIdx.x == 0) x_counts(b, p.lay, half, p.rank)[q] = cnt;
    }
}

// Publish: every CTA fences its stores at system scope and counts itself; the last one writes `epoch` into the flag word
// [half][rank] of every peer.
__device__ __forceinline__ void xpush_publish(const XPushParams& p, uint32_t n_ctas) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(p.done, 1u);
        if (prev + 1 == n_ctas) {
            *p.done = 0;  // ready for the next exchange (stream order separates the launches)
            __threadfence_system();
            const uint32_t half = p.epoch & 1u;
            for (uint32_t pr = 0; pr < p.lay.world; ++pr)
                if ((p.peer_mask >> pr) & 1u) st_release_sys(x_flag(p.tab->base[pr], p.lay, half, p.rank), p.epoch);
        }
    }
}

}  // namespace vg
