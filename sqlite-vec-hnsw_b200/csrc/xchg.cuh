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
                atomicExch(p.err, 1u);
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
