// hnsw_dev.cuh — K6: the whole layered HNSW search on the device, one warp per query.
//
// Replaces search_hnsw / search_layer (src/hnsw/search.rs:267-543) and the search half of insert_hnsw
// (src/hnsw/insert.rs:396-430) when the adjacency lists are resident in HBM: no host round trip per
// expansion step.  The traversal is the reference's, statement for statement:
//   - the entry point is scored first (search.rs:385-398);
//   - pop the closest unexpanded candidate; the layer ends when it is farther than the worst result (:406-410);
//   - its neighbours are filtered through a visited set BEFORE scoring (:424-434), scored (K5 arithmetic, canonical
//     order, bit-identical to pair_kernel) and admitted IN ADJACENCY ORDER with `len < ef || d < worst` (:516),
//     results trimmed to ef by evicting the largest (d, node) (:528-531);
//   - the closest result seeds the next layer (:318-323); layers above the collecting ones run with ef = 1.
//
// Data structure: ONE sorted array per warp in shared memory holds both heaps of the reference.  Key =
// order_bits(d) << 32 | node << 1 | expanded.  The first min(len, ef) entries are the result set; the unexpanded
// entries are the candidate heap.  An admitted entry that falls behind position ef with a distance strictly larger
// than the worst result can never be expanded (the reference would stop on popping it), so it is dropped; entries
// behind ef that TIE with the worst result's distance stay (the reference does expand those).  With that rule the
// reference's stop condition is exactly "no unexpanded entry left".
// The visited set is an open-addressing table in global memory (L2-resident), one per warp.  At most ef - 1 entries can sit behind position ef (they are copies of the worst result's distance pushed out one
// admission at a time), so an array of 2 ef slots never overflows.
// The visited set is bounded: when it is 3/4 full (or the step limit is hit) status[q] = 1 and the host re-runs that
// query through the lockstep driver: never a silent approximation.
#pragma once
#include "kernels.cuh"

namespace vg {

struct HGraphDev {
    const uint32_t* nbr0;        // [node][max_m0]
    const uint16_t* deg0;        // [node]
    const uint32_t* upper_base;  // [node] first upper-level slot of the node
    const uint32_t* nbrU;        // [slot][M]
    const uint16_t* degU;        // [slot]
    uint32_t max_m0, M;
};

struct HSearchParams {
    HGraphDev g;
    const uint8_t* a_base;       // query rows
    uint32_t a_stride;
    const uint32_t* a_index;     // row of query q in a_base (NULL: q)
    const uint8_t* b_base;       // slab rows
    uint32_t b_stride, units, qc_kind;
    uint32_t nq, entry;
    int32_t entry_level;
    const int8_t* node_level;    // inserts: level of the new node (layers <= it collect ef_wide results); NULL: queries
    uint32_t ef_wide, cap, take; // cap = array capacity (> ef_wide), take = entries written per collected layer
    uint32_t* visited;           // [warp][vis_size], vis_size a power of two
    uint32_t vis_size;
    const uint32_t* out_off;     // first output slot of query q (NULL: q); layer lv goes to slot out_off[q] + lv
    uint64_t* out_keys;          // [slot][take]
    uint32_t* out_cnt;           // [slot]
    uint32_t* status;            // [q]
    unsigned int* next_q;        // work counter
    unsigned long long* scored;  // distances computed
    unsigned long long* hist;    // [5] expansions by number of unvisited neighbours scored: 1-4, 5-16, 17-32, 33-64, 65+
                                 // (the reference's BATCH_SIZE_* counters, src/hnsw/search.rs:443-455)
    uint32_t prefetch;           // CTA kernel: 1 = the helper warps fetch the likely next expansion while warp 0 updates the array
    uint32_t spec_rows;          // CTA kernel: 1 = ... all the way into a second row buffer in shared memory (0: only into L2)
    uint32_t cta_vis;            // CTA kernel: visited slots in shared memory (power of two)
    uint32_t batch_admit;        // CTA kernel: 1 = admit the neighbours of an expansion in one merge (0: one sorted insert each)
    unsigned long long* prof;    // developer hook (VECGPU_HNSW_TIMING): [20] counters of the instrumented CTA kernel — cycles of its 4 phases,
                                 // merged / one-by-one admission batches, staged / all expansions, admission split, helper timeline; or NULL
    uint32_t max_steps;
    uint32_t q_smem;             // 1: each warp stages its query in shared memory (units * 16 bytes per warp)
};

static constexpr uint32_t HV_EMPTY = 0xFFFFFFFFu;
static constexpr uint32_t HV_TOMB = 0xFFFFFFFEu;  // a slot given back (never matches a node id, never reused)

__device__ __forceinline__ bool hvis_insert(uint32_t* t, uint32_t mask, uint32_t key) {  // true if newly inserted
    uint32_t h = (key * 2654435761u) & mask;
    while (true) {
        const uint32_t old = atomicCAS(&t[h], HV_EMPTY, key);
        if (old == HV_EMPTY) return true;
        if (old == key) return false;
        h = (h + 1) & mask;
    }
}

// insert x into the ascending array L[0..len) (order = key >> 1; nodes are unique so there are no equal keys).
// Warp-uniform arguments; returns the position.
__device__ __forceinline__ uint32_t hlist_insert(uint64_t* L, uint32_t len, uint64_t x, int lane) {
    const uint64_t xk = x | 1ull;
    // The position is searched from the TAIL: an admitted neighbour almost always lands just ahead of the worst result, so
    // one or two 32-entry chunks are looked at (and shifted) instead of the whole array from the front — the sorted insert
    // was most of a single query's latency (~0.7 us each at ef = 200, thousands per query).
    uint32_t pos = len;
    for (int c = len ? (int)((len - 1) >> 5) : -1; c >= 0; --c) {
        const uint32_t i = ((uint32_t)c << 5) + lane;
        const bool gt = i < len && (L[i] | 1ull) > xk;
        const uint32_t m = __ballot_sync(0xffffffffu, gt);
        const uint32_t n_gt = __popc(m), valid = min(32u, len - ((uint32_t)c << 5));
        pos -= n_gt;
        if (n_gt < valid) break;  // this chunk holds an entry below x: everything further down is below x too
    }
    if (len > 0) {  // shift [pos, len) up by one, highest chunk first
        for (int c = (int)((len - 1) >> 5); c >= (int)(pos >> 5); --c) {
            const uint32_t i = ((uint32_t)c << 5) + lane;
            const bool mv = i < len && i >= pos;
            const uint64_t v = mv ? L[i] : 0ull;
            __syncwarp();
            if (mv) L[i + 1] = v;
        }
    }
    __syncwarp();
    if (lane == 0) L[pos] = x;
    __syncwarp();
    return pos;
}

#ifndef VECGPU_HW_GATHER
#define VECGPU_HW_GATHER 4
#endif
static constexpr int HW_GATHER = VECGPU_HW_GATHER;  // row pieces in flight per lane in the one-warp walk

template <class T>
__global__ void __launch_bounds__(256) hnsw_search_kernel(const HSearchParams p) {
    constexpr int LPR = T::LPR;
    constexpr int GPW = 32 / LPR;  // rows scored per pass of the warp
    extern __shared__ __align__(16) uint8_t h_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane % LPR;
    const uint32_t pend_cap = (p.g.max_m0 + 31u) & ~31u;
    const size_t per_warp = (size_t)p.cap * 8 + (size_t)pend_cap * 4 + (p.q_smem ? (size_t)p.units * 16 : 0);
    uint64_t* L = (uint64_t*)(h_smem + (size_t)warp * per_warp);
    uint32_t* pend = (uint32_t*)(L + p.cap);
    uint4* sq = (uint4*)(pend + pend_cap);  // the query, staged once: the row stream evicts it from L1 between passes
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* vt = p.visited + (size_t)gw * p.vis_size;

    while (true) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(p.next_q, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= p.nq) break;
        const uint32_t ai = p.a_index ? p.a_index[q] : q;
        const uint4* a = (const uint4*)(p.a_base + (uint64_t)ai * p.a_stride);
        if (p.q_smem) {
            __syncwarp();
            for (uint32_t u = lane; u < p.units; u += 32) sq[u] = __ldg(a + u);
            __syncwarp();
            a = sq;
        }
        float qc = 0.f;
        if (T::HAS_QC) qc = query_const(a, p.units, lane & 3, p.qc_kind);
        const int nlev = p.node_level ? (int)p.node_level[q] : -1;
        const uint32_t slot0 = p.out_off ? p.out_off[q] : q;
        uint32_t entry = p.entry;
        uint32_t status = 0;
        unsigned long long nscored = 0;
        uint32_t hb[5] = {0, 0, 0, 0, 0};

        for (int level = p.entry_level; level >= 0 && !status; --level) {
            const bool wide = nlev < 0 ? level == 0 : level <= nlev;
            const uint32_t ef = wide ? p.ef_wide : 1u;
            const uint32_t vsz = ef == 1u ? min(p.vis_size, 4096u) : p.vis_size;
            const uint32_t vmask = vsz - 1, vlimit = vsz - (vsz >> 2);
            for (uint32_t i = (uint32_t)lane * 4; i < vsz; i += 128) *(uint4*)(vt + i) = make_uint4(HV_EMPTY, HV_EMPTY, HV_EMPTY, HV_EMPTY);
            __syncwarp();
            uint32_t len = 0, lo = 0, vcount = 1, npend = 1, steps = 0;
            uint32_t worst_hi = 0xFFFFFFFFu;
            if (lane == 0) {
                pend[0] = entry;
                hvis_insert(vt, vmask, entry);
            }
            __syncwarp();

            while (true) {
                // ---- score the pending nodes, GPW at a time, and admit them in order
                for (uint32_t base = 0; base < npend && !status; base += GPW) {
                    const uint32_t r = base + (uint32_t)(lane / LPR);
                    const bool valid = r < npend;
                    const uint32_t node = valid ? pend[r] : 0u;
                    typename T::Acc acc;
                    T::init(acc);
                    if (valid) {
                        const uint4* b = (const uint4*)(p.b_base + (uint64_t)node * p.b_stride);
                        // HW_GATHER 16-byte pieces of the row are requested before the first is used: the walk is bound by the
                        // latency of these gathers, and left alone the compiler keeps two in flight per lane.  Measured on
                        // 20 000 queries over 1 M x 384 (ef = 200): 550 k q/s as compiled before, 727 k with 4 (64 registers,
                        // 4 CTAs per SM), 651 k with 8 (80 registers, 3 CTAs), 557 k with 12 (122 registers, 2 CTAs)
#pragma unroll 1
                        for (uint32_t u0 = g; u0 < p.units; u0 += LPR * HW_GATHER) {
                            uint4 xv[HW_GATHER];
#pragma unroll
                            for (int j = 0; j < HW_GATHER; ++j) {
                                const uint32_t u = u0 + (uint32_t)j * LPR;
                                if (u < p.units) xv[j] = __ldg(b + u);
                            }
#pragma unroll
                            for (int j = 0; j < HW_GATHER; ++j) {
                                const uint32_t u = u0 + (uint32_t)j * LPR;
                                if (u < p.units) {
                                    const uint4 qv[1] = {a[u]};
                                    T::step(acc, xv[j], qv);
                                }
                            }
                        }
                    }
                    const float d = T::finish(acc, 0, &qc);
                    const uint32_t cnt = min((uint32_t)GPW, npend - base);
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const float dj = __shfl_sync(0xffffffffu, d, (int)j * LPR);
                        const uint32_t nj = __shfl_sync(0xffffffffu, node, (int)j * LPR);
                        if (dj != dj) continue;  // NaN never enters a heap
                        const uint32_t oj = order_bits(dj);
                        if (len < ef || oj < worst_hi) {  // search.rs:516 (strict <)
                            const uint64_t key = ((uint64_t)oj << 32) | ((uint64_t)nj << 1);
                            const uint32_t pos = hlist_insert(L, len, key, lane);
                            ++len;
                            if (pos < lo) lo = pos;
                            if (len > ef) {  // keep only the entries behind ef that tie with the worst result
                                const uint32_t wd = (uint32_t)(L[ef - 1] >> 32);
                                uint32_t keep = ef;
                                for (uint32_t c = ef; c < len; c += 32) {
                                    const uint32_t i = c + lane;
                                    const bool tie = i < len && (uint32_t)(L[i] >> 32) == wd;
                                    const uint32_t m = __ballot_sync(0xffffffffu, tie);
                                    keep += __popc(m);
                                    if (m != 0xffffffffu) break;
                                }
                                len = keep;
                            }
                            worst_hi = (uint32_t)(L[min(len, ef) - 1] >> 32);
                            if (len >= p.cap) {
                                status = 1;
                                break;
                            }
                        }
                    }
                }
                nscored += npend;
                if (status) break;

                // ---- pop the closest unexpanded entry until one has unvisited neighbours
                npend = 0;
                bool layer_done = false;
                while (npend == 0) {
                    uint32_t ci = 0xFFFFFFFFu;
                    for (uint32_t c = lo & ~31u; c < len; c += 32) {
                        const uint32_t i = c + lane;
                        const bool un = i < len && i >= lo && !(L[i] & 1ull);
                        const uint32_t m = __ballot_sync(0xffffffffu, un);
                        if (m) {
                            ci = c + (uint32_t)__ffs(m) - 1u;
                            break;
                        }
                    }
                    if (ci == 0xFFFFFFFFu) {  // == the reference's stop: every remaining candidate is farther than the worst result
                        layer_done = true;
                        break;
                    }
                    const uint64_t ck = L[ci];
                    __syncwarp();
                    if (lane == 0) L[ci] = ck | 1ull;
                    __syncwarp();
                    lo = ci + 1;
                    const uint32_t cn = (uint32_t)(ck & 0xFFFFFFFFull) >> 1;
                    const uint32_t* nb;
                    uint32_t deg;
                    if (level == 0) {
                        nb = p.g.nbr0 + (size_t)cn * p.g.max_m0;
                        deg = p.g.deg0[cn];
                    } else {
                        const size_t slot = (size_t)p.g.upper_base[cn] + (size_t)(level - 1);
                        nb = p.g.nbrU + slot * p.g.M;
                        deg = p.g.degU[slot];
                    }
                    if (vcount + deg > vlimit || ++steps > p.max_steps) {
                        status = 1;
                        break;
                    }
                    for (uint32_t i0 = 0; i0 < deg; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        uint32_t v = 0;
                        bool isnew = false;
                        if (i < deg) {
                            v = nb[i];
                            isnew = hvis_insert(vt, vmask, v);
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, isnew);
                        if (isnew) pend[npend + __popc(m & ((1u << lane) - 1u))] = v;
                        npend += __popc(m);
                    }
                    vcount += npend;
                    __syncwarp();
                }
                if (layer_done || status) break;
                hb[npend <= 4 ? 0 : npend <= 16 ? 1 : npend <= 32 ? 2 : npend <= 64 ? 3 : 4]++;
            }
            if (status) break;
            if (len > 0) entry = (uint32_t)(L[0] & 0xFFFFFFFFull) >> 1;  // closest result seeds the next layer
            if (wide) {
                const uint32_t slot = slot0 + (uint32_t)level;
                const uint32_t cnt = min(min(len, ef), p.take);
                for (uint32_t i = lane; i < cnt; i += 32) p.out_keys[(size_t)slot * p.take + i] = L[i];
                if (lane == 0) p.out_cnt[slot] = cnt;
            }
            __syncwarp();
        }
        if (lane == 0) {
            p.status[q] = status;
            atomicAdd(p.scored, nscored);
            if (p.hist)
                for (int b = 0; b < 5; ++b)
                    if (hb[b]) atomicAdd(p.hist + b, (unsigned long long)hb[b]);
        }
    }
}

// ---- K6c: the same walk with ONE CTA per query — the latency form (few queries: SQL issues one per MATCH, src/vtab.rs:2249) ----
// A lone warp spends most of an expansion waiting: its <= 32 fresh neighbours are scored 8 at a time, every admitted one is
// a sorted insert, the visited set is a table of atomics in global memory — ~24 us per expansion.  A single query is a chain
// of ~ef dependent expansions, so what counts is the length of that chain in (mostly dependent) instructions of one warp:
// about 6-9 cycles each.  Here 8 warps share one query and the chain is cut three ways:
//   - SCORING (warps 0..3): every 4-lane (or 1-lane) group takes one pending neighbour; the rows are in shared memory (one
//     bulk copy per row, all in flight at once) and each lane requests 8 pieces of row and query before using the first.
//   - ADMISSION (warps 0..3): the <= 32 scored neighbours that can still beat the worst result are merged into the sorted
//     result/candidate array in ONE pass — every key's place is its rank (new keys below it: warps 0..2, a third each;
//     entries of the array below it: warp 3, binary search), the old entries take the remaining places in order (32 places
//     per warp at a time), written into a second array that then becomes the current one.  Without distance ties at the cut
//     after ef entries this equals the reference's one-by-one admission (search.rs:516) whatever the order; with a tie AT
//     THE CUT the order matters, and the batch is replayed one by one (hlist_admit_one_by_one), exactly as the one-warp kernel.
//   - LOOK-AHEAD (warps 4..7, led by HC_LEAD_WARP): while the others score, the lead finds the closest entry not expanded yet
//     and asks for the adjacency lists of every node that can be popped next into L2; when the scores are known it names the
//     candidate the pop WILL return (the closer of that entry and the best admissible new neighbour — provably the next pop
//     unless a tie at the cut intervenes) and expands it already, while warps 0..3 are busy with the array: neighbours read,
//     the unvisited ones entered into the visited set and listed, their rows copied into the second row buffer (all four
//     helper warps issue; a lone warp would serialise 23 copies).  Warp 0's pop compares: same candidate -> list and rows ARE
//     the next expansion (no adjacency read, no inserts, no fetch on the critical path: ~97 % of the expansions of a
//     1 M-row walk); different candidate -> the helper's visited entries are taken back (tombstones) and the step is done
//     the plain way.  Every path yields the one-warp kernel's results bit for bit (tests compare them, ties included).
//   - the visited set is an open-addressing table in SHARED memory (16 K slots for ef <= 256 — which leaves room for the
//     second row buffer — else 32 K), cleared per layer by the whole CTA.
// Queries are still handed out by the atomic counter (grid = min(nq, SMs)).  A query that overflows the table or the
// array is flagged and answered by the other paths, exactly as before.  Measured (profiles/r2_hnsw_latency4.txt): 0.63 ms
// per single query at ef = 200 on 1 M x 384 (one-warp kernel: 3.4-5.5 ms), 0.23 ms at ef = 50.
static constexpr uint32_t HC_VIS = 32768;       // shared-memory visited slots (power of two)
static constexpr uint32_t HC_THREADS = 256;
static constexpr uint32_t HC_ROWS = 32;         // neighbour rows staged in shared memory per scoring pass
static constexpr uint32_t HC_MERGE_WARPS = 4;   // warps 0..3 update the result array, warps 4..7 fetch ahead
static constexpr uint32_t HC_LEAD_WARP = 5;     // ... led by this one (not on warp 0's scheduler: warp w issues on sub-partition w % 4)


__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem, uint32_t bytes) {  // bytes: multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool hvis_insert_smem(uint32_t* t, uint32_t mask, uint32_t key) {
    uint32_t h = (key * 2654435761u) & mask;
    while (true) {
        const uint32_t old = atomicCAS(&t[h], HV_EMPTY, key);
        if (old == HV_EMPTY) return true;
        if (old == key) return false;
        h = (h + 1) & mask;
    }
}

// One-by-one admission of the neighbours flagged in `todo` (adjacency order, each re-checked against the then-current worst
// result: search.rs:516) — the general form, needed only when a batch's cut falls inside a distance tie.  Out of line: the
// walk's inner loop is executed by ONE warp, so its instruction footprint is its speed (I-cache: 6 KB L0, 32 KB L1.5).
__device__ __noinline__ uint32_t hlist_admit_one_by_one(uint64_t* L, uint32_t len, uint32_t lo, uint32_t ef, uint32_t cap,
                                                         const uint32_t* pend_chunk, uint32_t ol, uint32_t todo, uint32_t* len_lo_out) {
    const int lane = threadIdx.x & 31;
    uint32_t status = 0;
    uint32_t worst_hi = len ? (uint32_t)(L[min(len, ef) - 1] >> 32) : 0xFFFFFFFFu;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t oj = __shfl_sync(0xffffffffu, ol, src);
        const uint32_t nj = pend_chunk[src];
        if (len < ef || oj < worst_hi) {  // search.rs:516 (strict <)
            const uint64_t key = ((uint64_t)oj << 32) | ((uint64_t)nj << 1);
            const uint32_t pos = hlist_insert(L, len, key, lane);
            ++len;
            if (pos < lo) lo = pos;
            if (len > ef) {  // keep only the entries behind ef that tie with the worst result
                const uint32_t wd = (uint32_t)(L[ef - 1] >> 32);
                uint32_t keep = ef;
                for (uint32_t c = ef; c < len; c += 32) {
                    const uint32_t i = c + lane;
                    const bool tie = i < len && (uint32_t)(L[i] >> 32) == wd;
                    const uint32_t m = __ballot_sync(0xffffffffu, tie);
                    keep += __popc(m);
                    if (m != 0xffffffffu) break;
                }
                len = keep;
            }
            worst_hi = (uint32_t)(L[min(len, ef) - 1] >> 32);
            if (len >= cap) {
                status = 1;
                break;
            }
        }
    }
    if (lane == 0) {  // (shared memory: keeps the caller's len / lo in registers)
        len_lo_out[0] = len;
        len_lo_out[1] = lo;
    }
    __syncwarp();
    return status;
}

template <class T, bool PROF>
__global__ void __launch_bounds__(HC_THREADS) hnsw_search_cta_kernel(const HSearchParams p) {
    constexpr int LPR = T::LPR;
    constexpr int GROUPS = HC_THREADS / LPR;  // rows scored per pass of the CTA
    extern __shared__ __align__(16) uint8_t h_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = threadIdx.x % LPR, grp = threadIdx.x / LPR;
    const uint32_t pend_cap = (p.g.max_m0 + 31u) & ~31u;
    // layout: [visited cta_vis x 4][L cap x 8][L2 cap x 8][pend ids][pend dist][predicted ids][query]
    //         [staged rows HC_ROWS x units x 16][second row buffer, if spec_rows]
    uint32_t* vt = (uint32_t*)h_smem;
    uint64_t* La = (uint64_t*)(vt + p.cta_vis);
    uint64_t* Lb = La + p.cap;  // the batch admission merges from one array into the other
    uint32_t* pend = (uint32_t*)(La + 2 * (size_t)p.cap);
    float* pend_d = (float*)(pend + pend_cap);
    uint32_t* ppend = (uint32_t*)(pend_d + pend_cap);
    uint4* sq = (uint4*)(ppend + pend_cap);
    uint4* srow = sq + p.units;                          // rows being scored
    uint4* srow_next = srow + (size_t)HC_ROWS * p.units;  // rows of the predicted next expansion (spec_rows)
    __shared__ uint32_t s_q, s_npend, s_status, s_done, s_entry, s_pred, s_pnp, s_hb[5];
    __shared__ uint32_t s_len, s_worst, s_cur, s_mstatus, s_rank[4][32], s_lenlo[2], s_lo, s_vcount, s_pred_cn, s_accepted;  // the admission's state, shared by warps 0..3
    __shared__ __align__(8) uint64_t s_bar[2];  // mbarriers of the row copies: [0] rows fetched ahead by the helper warp, [1] rows fetched on demand
    __shared__ unsigned long long s_pf[20];
    __shared__ long long s_t0;
    const uint32_t bar_ahead = smem_u32(&s_bar[0]), bar_now = smem_u32(&s_bar[1]);
    uint32_t par_ahead = 0, par_now = 0;  // phase parities, tracked alike by every thread
    const uint32_t row_bytes = p.units * 16u;
    if (threadIdx.x == 0) {
        mbar_init(bar_ahead, 1);
        mbar_init(bar_now, 1);
        mbar_fence_init();
    }
    __syncthreads();

    while (true) {
        if (threadIdx.x == 0) s_q = atomicAdd(p.next_q, 1u);
        if (threadIdx.x < 5) s_hb[threadIdx.x] = 0;
        if (PROF && threadIdx.x < 20) s_pf[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= p.nq) break;
        const uint32_t ai = p.a_index ? p.a_index[q] : q;
        const uint4* ag = (const uint4*)(p.a_base + (uint64_t)ai * p.a_stride);
#pragma unroll 1
        for (uint32_t u = threadIdx.x; u < p.units; u += HC_THREADS) sq[u] = __ldg(ag + u);
        __syncthreads();
        const uint4* a = sq;
        float qc = 0.f;
        if (T::HAS_QC) qc = query_const(a, p.units, lane & 3, p.qc_kind);  // every 4-lane group computes the same value
        const int nlev = p.node_level ? (int)p.node_level[q] : -1;
        const uint32_t slot0 = p.out_off ? p.out_off[q] : q;
        if (threadIdx.x == 0) {
            s_entry = p.entry;
            s_status = 0;
        }
        unsigned long long nscored = 0;
        long long pt = PROF ? clock64() : 0;  // thread 0: row fetch, scoring, admission, pop + adjacency + visited
        long long pt2 = 0;
        auto sub = [&](int i) {  // finer split of the admission phase (slots 8..11: prediction, ranks, placement, finish)
            if (PROF && threadIdx.x == 0) {
                const long long t = clock64();
                if (i >= 0) s_pf[i] += (unsigned long long)(t - pt2);
                pt2 = t;
            }
        };
        auto lap = [&](int i) {
            if (PROF && threadIdx.x == 0) {
                const long long t = clock64();
                s_pf[i] += (unsigned long long)(t - pt);
                pt = t;
            }
        };
        __syncthreads();

        for (int level = p.entry_level; level >= 0; --level) {
            if (s_status) break;
            const bool wide = nlev < 0 ? level == 0 : level <= nlev;
            const uint32_t ef = wide ? p.ef_wide : 1u;
            const uint32_t vsz = ef == 1u ? 4096u : p.cta_vis;
            const uint32_t vmask = vsz - 1, vlimit = vsz - (vsz >> 2);
#pragma unroll 2
            for (uint32_t i = threadIdx.x * 4; i < vsz; i += HC_THREADS * 4) *(uint4*)(vt + i) = make_uint4(HV_EMPTY, HV_EMPTY, HV_EMPTY, HV_EMPTY);
            __syncthreads();
            // warp 0's registers carry the layer state (len / worst / which array are mirrored in shared memory for warps 1..3)
            uint32_t len = 0, lo = 0, vcount = 1, steps = 0;
            uint32_t worst_hi = 0xFFFFFFFFu;
            uint64_t* L = La;
            if (threadIdx.x == 0) {
                s_len = 0;
                s_lo = 0;
                s_vcount = 1;
                s_pred_cn = 0xFFFFFFFFu;
                s_accepted = 0;
                s_worst = 0xFFFFFFFFu;
                s_cur = 0;
                s_mstatus = 0;
                pend[0] = s_entry;
                hvis_insert_smem(vt, vmask, s_entry);
                s_npend = 1;
                s_done = 0;
                s_pnp = 0xFFFFFFFFu;
            }
            __syncthreads();

            while (true) {
                // ---- scoring phase: one group per pending node, all groups at once
                const uint32_t npend = s_npend;
                const uint32_t pnp = s_pnp;
                const bool ahead = p.spec_rows && pnp >= 1u && pnp <= HC_ROWS;  // the helper warps have copies in flight into srow_next
                const bool accepted = s_accepted != 0;  // warp 0 took the helper's list of fresh neighbours as this expansion
                const bool staged = ahead && accepted;  // ... so those copies are this expansion's rows
                if (accepted) {
                    uint32_t* t = pend;
                    pend = ppend;
                    ppend = t;
                }
                if (ahead) {  // wait for them either way: the buffer and the barrier are used again
                    mbar_wait(bar_ahead, par_ahead);
                    par_ahead ^= 1u;
                }
                if (PROF && threadIdx.x == 0 && ahead) s_pf[17] += (unsigned long long)(clock64() - s_t0);
                if (staged) {
                    uint4* t = srow;
                    srow = srow_next;
                    srow_next = t;
                }
                if (PROF && threadIdx.x == 0) {
                    s_pf[6] += staged ? 1 : 0;
                    s_pf[7] += 1;
                }
                // The helper warp has no rows to score (<= 32 rows occupy warps 0..3).  While the others score it finds the
                // closest entry not expanded yet: together with the closest admissible new neighbour (known after scoring)
                // that names the candidate that will most likely be expanded next — a hint, what it fetches is verified.
                uint64_t pred_key = ~0ull;
                if (warp == HC_LEAD_WARP && p.prefetch) {
                    const uint64_t* Lc = s_cur ? Lb : La;
                    const uint32_t clen = s_len, clo = s_lo;
#pragma unroll 1
                    for (uint32_t c = clo & ~31u; c < clen; c += 32) {
                        const uint32_t i = c + lane;
                        const uint32_t m = __ballot_sync(0xffffffffu, i < clen && i >= clo && !(Lc[i] & 1ull));
                        if (m) {
                            pred_key = Lc[c + (uint32_t)__ffs(m) - 1u];
                            break;
                        }
                    }
                    // ... and asks for the adjacency lists of everyone who can be next (that entry and the nodes being
                    // scored), so that the winner's list comes from L2 instead of HBM
                    if (level == 0) {
#pragma unroll 1
                        for (uint32_t j = lane; j <= npend; j += 32) {
                            const uint32_t node = j < npend ? pend[j] : (pred_key == ~0ull ? 0xFFFFFFFFu : (uint32_t)(pred_key & 0xFFFFFFFFull) >> 1);
                            if (node != 0xFFFFFFFFu) {
                                const uint8_t* row = (const uint8_t*)(p.g.nbr0 + (size_t)node * p.g.max_m0);
                                for (uint32_t o = 0; o < p.g.max_m0 * 4u; o += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.g.deg0 + node));
                            }
                        }
                    }
                }
                for (uint32_t base = 0; base < npend; base += HC_ROWS) {
                    // every warp fires the 16-byte pieces of its rows at once (cp.async: no registers, no waiting in
                    // between): the whole expansion costs ONE memory round trip, not units / 4 dependent ones per lane
                    const uint32_t nr = min(HC_ROWS, npend - base);
                    if (!staged) {  // one bulk copy per row (TMA engine), all in flight at once: ONE memory round trip
                        if (warp == 0) {
                            if (lane == 0) mbar_expect_tx(bar_now, nr * row_bytes);
                            __syncwarp();
                            if ((uint32_t)lane < nr)
                                bulk_g2s(smem_u32(srow + (size_t)lane * p.units), p.b_base + (uint64_t)pend[base + lane] * p.b_stride, row_bytes, bar_now);
                        }
                        mbar_wait(bar_now, par_now);
                        par_now ^= 1u;
                    }
                    lap(0);
#pragma unroll 1
                    for (uint32_t r0 = 0; r0 < nr; r0 += GROUPS) {  // uniform trip count: T::finish shuffles across the warp
                        const uint32_t r = r0 + (uint32_t)grp;
                        const bool valid = r < nr;
                        typename T::Acc acc;
                        T::init(acc);
                        if (valid) {
                            // eight 16-byte pieces of the row and of the query are requested before the first is used
                            // (left to itself the compiler loads a pair, waits ~30 cycles, uses it, loads the next pair)
                            const uint32_t bs = smem_u32(srow + (size_t)r * p.units), as = smem_u32(a);
#pragma unroll 1
                            for (uint32_t u0 = g; u0 < p.units; u0 += LPR * 8) {
                                uint4 bv[8], av[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t u = u0 + (uint32_t)j * LPR;
                                    if (u < p.units) {
                                        bv[j] = lds128(bs + u * 16u);
                                        av[j] = lds128(as + u * 16u);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t u = u0 + (uint32_t)j * LPR;
                                    if (u < p.units) {
                                        const uint4 qv[1] = {av[j]};
                                        T::step(acc, bv[j], qv);
                                    }
                                }
                            }
                        }
                        const float d = T::finish(acc, 0, &qc);
                        if (valid && g == 0) pend_d[base + r] = d;
                    }
                    __syncthreads();
                }
                nscored += (threadIdx.x == 0) ? npend : 0;
                lap(1);
                if (PROF && threadIdx.x == 0) s_t0 = clock64();

                // ---- control phase.  Warps 0..3 admit the scored neighbours (one merge per 32, the work split over the four
                // warps), then warp 0 pops until a candidate has fresh neighbours; warps 4..7 fetch ahead (below).
                uint32_t status = 0;  // warp 0: a table or the array is full (the query is answered by the other paths)
                if (warp < HC_MERGE_WARPS) {
                    sub(-1);
                    sub(8);
#pragma unroll 1
                    for (uint32_t j0 = 0; j0 < npend; j0 += 32) {
                        // The worst result only ever gets better, so a neighbour that fails `d < worst` NOW (with the set
                        // full) fails when its turn comes too: those are discarded 32 at a time; the survivors are admitted
                        // in one merge, or one by one, in adjacency order, re-checked against the then-current worst
                        // (search.rs:516), when the merge finds a tie at its cut.  Every key's place in the merge is its
                        // rank: (new keys below it: warps 0..2, a third of the keys each) + (entries of L below it: warp 3);
                        // the old entries take the remaining places in order (all four warps, 32 places at a time).
                        const uint32_t clen = s_len, cworst = s_worst;  // == warp 0's len / worst_hi
                        uint64_t* Lc = s_cur ? Lb : La;
                        uint64_t* Ln = s_cur ? La : Lb;
                        const uint32_t jj = j0 + lane;
                        const float dl = jj < npend ? pend_d[jj] : __int_as_float(0x7FC00000);
                        const uint32_t ol = order_bits(dl);
                        const bool maybe = jj < npend && dl == dl && (clen < ef || ol < cworst);
                        const uint32_t todo = __ballot_sync(0xffffffffu, maybe);
                        if (!todo) continue;  // the same in all four warps
                        if (p.batch_admit) {
                            const uint64_t key = ((uint64_t)ol << 32) | ((uint64_t)(maybe ? pend[jj] : 0u) << 1);
                            uint32_t r = 0;
                            if (warp < 3) {
#pragma unroll 1
                                for (uint32_t t = todo & (warp == 0 ? 0x7FFu : warp == 1 ? 0x3FF800u : 0xFFC00000u); t; t &= t - 1) {
                                    const uint64_t kb = __shfl_sync(0xffffffffu, key, __ffs(t) - 1);
                                    r += kb < key ? 1u : 0u;  // nodes are unique: no equal keys
                                }
                            } else if (maybe) {  // lower bound of the key in L
                                const uint64_t xk = key | 1ull;
                                uint32_t lb = 0, ub = clen;
#pragma unroll 1
                                while (lb < ub) {
                                    const uint32_t mid = (lb + ub) >> 1;
                                    if ((Lc[mid] | 1ull) < xk) lb = mid + 1;
                                    else ub = mid;
                                }
                                r = lb;
                            }
                            s_rank[warp][lane] = r;
                            asm volatile("bar.sync 3, %0;" ::"n"(32 * HC_MERGE_WARPS) : "memory");
                            sub(9);
                            const uint32_t pos = s_rank[0][lane] + s_rank[1][lane] + s_rank[2][lane] + s_rank[3][lane];
                            const uint32_t total = clen + __popc(todo), out_n = min(total, ef + 1u);  // entries [0, ef] are all that can matter
                            if (warp == 0 && maybe && pos < out_n) Ln[pos] = key;
#pragma unroll 1
                            for (uint32_t bo = 32u * warp; bo < out_n; bo += 32u * HC_MERGE_WARPS) {
                                const uint32_t rel = pos - bo;
                                const uint32_t mask_new = __reduce_or_sync(0xffffffffu, (maybe && rel < 32u) ? (1u << rel) : 0u);
                                const uint32_t below = __popc(__ballot_sync(0xffffffffu, maybe && pos < bo));
                                const uint32_t t = bo + lane;
                                if (!((mask_new >> lane) & 1u) && t < out_n) Ln[t] = Lc[t - below - __popc(mask_new & ((1u << lane) - 1u))];
                            }
                            asm volatile("bar.sync 3, %0;" ::"n"(32 * HC_MERGE_WARPS) : "memory");
                            sub(10);
                            if (warp == 0) {
                                if (total > ef && (uint32_t)(Ln[ef - 1] >> 32) == (uint32_t)(Ln[ef] >> 32)) {  // tie at the cut: order matters
                                    status = hlist_admit_one_by_one(L, len, lo, ef, p.cap, pend + j0, ol, todo, s_lenlo);
                                    len = s_lenlo[0];
                                    lo = s_lenlo[1];
                                    if (PROF && lane == 0) s_pf[5] += 1;
                                } else {
                                    len = min(total, ef);
                                    lo = min(lo, __reduce_min_sync(0xffffffffu, maybe ? pos : 0xFFFFFFFFu));
                                    L = Ln;
                                    if (lane == 0) s_cur ^= 1u;
                                    if (PROF && lane == 0) s_pf[4] += 1;
                                }
                            }
                        } else if (warp == 0) {
                            status = hlist_admit_one_by_one(L, len, lo, ef, p.cap, pend + j0, ol, todo, s_lenlo);
                            len = s_lenlo[0];
                            lo = s_lenlo[1];
                            if (PROF && lane == 0) s_pf[5] += 1;
                        }
                        if (warp == 0) {
                            worst_hi = (uint32_t)(L[min(len, ef) - 1] >> 32);
                            if (lane == 0) {
                                s_len = len;
                                s_worst = worst_hi;
                                s_mstatus = status;
                            }
                        }
                        sub(11);
                        if (j0 + 32 >= npend) break;  // (warps 1..3 next read the state after the CTA-wide barrier)
                        asm volatile("bar.sync 3, %0;" ::"n"(32 * HC_MERGE_WARPS) : "memory");
                        if (s_mstatus) break;
                    }
                }
                if (warp == 0) {
                    lap(2);
                    uint32_t np2 = 0;
                    bool layer_done = false, accept = false;
                    uint32_t spec_cn = 0xFFFFFFFFu, spec_np = 0;
                    if (p.prefetch) {  // the helper warp's speculative expansion (its list is complete behind this barrier)
                        asm volatile("bar.sync 1, 64;" ::: "memory");
                        spec_cn = *(volatile uint32_t*)&s_pred_cn;
                        spec_np = *(volatile uint32_t*)&s_pnp;
                    }
#pragma unroll 1
                    while (!status && np2 == 0) {
                        uint32_t ci = 0xFFFFFFFFu;
#pragma unroll 1
                        for (uint32_t c = lo & ~31u; c < len; c += 32) {
                            const uint32_t i = c + lane;
                            const bool un = i < len && i >= lo && !(L[i] & 1ull);
                            const uint32_t m = __ballot_sync(0xffffffffu, un);
                            if (m) {
                                ci = c + (uint32_t)__ffs(m) - 1u;
                                break;
                            }
                        }
                        if (ci == 0xFFFFFFFFu) {  // == the reference's stop (search.rs:406-410)
                            layer_done = true;
                            break;
                        }
                        const uint64_t ck = L[ci];
                        __syncwarp();
                        if (lane == 0) L[ci] = ck | 1ull;
                        __syncwarp();
                        lo = ci + 1;
                        const uint32_t cn = (uint32_t)(ck & 0xFFFFFFFFull) >> 1;
                        accept = false;
                        if (spec_cn != 0xFFFFFFFFu) {
                            // The helper warp has already expanded the candidate it predicted: neighbours read, the fresh
                            // ones entered into the visited set and listed (ppend), their rows on the way.  If that is
                            // the candidate popped here, its work IS this expansion; if not, its entries are taken back.
                            const bool hit = cn == spec_cn;
                            spec_cn = 0xFFFFFFFFu;
                            if (hit) {
                                if (++steps > p.max_steps) {
                                    status = 1;
                                    break;
                                }
                                np2 = spec_np;
                                vcount += np2;
                                accept = true;
                                continue;
                            }
#pragma unroll 1
                            for (uint32_t j = lane; j < spec_np; j += 32) {
                                const uint32_t v = ppend[j];
                                uint32_t h = (v * 2654435761u) & vmask;
                                while (vt[h] != v) h = (h + 1) & vmask;
                                vt[h] = HV_TOMB;
                            }
                            vcount += spec_np;  // the slots stay occupied
                            __syncwarp();
                        }
                        const uint32_t* nb;
                        const uint16_t* dg;
                        uint32_t maxd;
                        if (level == 0) {
                            nb = p.g.nbr0 + (size_t)cn * p.g.max_m0;
                            dg = p.g.deg0 + cn;
                            maxd = p.g.max_m0;
                        } else {
                            const size_t slot = (size_t)p.g.upper_base[cn] + (size_t)(level - 1);
                            nb = p.g.nbrU + slot * p.g.M;
                            dg = p.g.degU + slot;
                            maxd = p.g.M;
                        }
                        const uint32_t first = (uint32_t)lane < maxd ? __ldg(nb + lane) : 0u;  // with the degree: one round trip
                        const uint32_t deg = *dg;
                        if (vcount + deg > vlimit || ++steps > p.max_steps) {
                            status = 1;
                            break;
                        }
#pragma unroll 1
                        for (uint32_t i0 = 0; i0 < deg; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            uint32_t v = 0;
                            bool isnew = false;
                            if (i < deg) {
                                v = i0 == 0 ? first : __ldg(nb + i);
                                isnew = hvis_insert_smem(vt, vmask, v);
                            }
                            const uint32_t m = __ballot_sync(0xffffffffu, isnew);
                            if (isnew) pend[np2 + __popc(m & ((1u << lane) - 1u))] = v;
                            np2 += __popc(m);
                        }
                        vcount += np2;
                        __syncwarp();
                    }
                    lap(3);
                    if (lane == 0) {
                        s_npend = np2;
                        s_lo = lo;
                        s_vcount = vcount;
                        s_accepted = (accept && np2 > 0) ? 1u : 0u;
                        s_done = (layer_done || status) ? 1u : 0u;
                        if (status) s_status = 1;
                        if (!layer_done && !status) s_hb[np2 <= 4 ? 0 : np2 <= 16 ? 1 : np2 <= 32 ? 2 : np2 <= 64 ? 3 : 4] += 1;
                    }
                    if (layer_done && !status) {  // results of this layer
                        if (len > 0 && lane == 0) s_entry = (uint32_t)(L[0] & 0xFFFFFFFFull) >> 1;  // closest result seeds the next layer
                        if (wide) {
                            const uint32_t slot = slot0 + (uint32_t)level;
                            const uint32_t cnt = min(min(len, ef), p.take);
#pragma unroll 1
                            for (uint32_t i = lane; i < cnt; i += 32) p.out_keys[(size_t)slot * p.take + i] = L[i];
                            if (lane == 0) p.out_cnt[slot] = cnt;
                        }
                    }
                } else if (warp >= HC_MERGE_WARPS && p.prefetch) {
                    // ---- helper warps: fetch what the predicted candidate's expansion will need
                    auto hs = [&](int i) {
                        if (PROF && threadIdx.x == 32 * HC_LEAD_WARP) s_pf[i] += (unsigned long long)(clock64() - *(volatile long long*)&s_t0);
                    };
                    if (warp == HC_LEAD_WARP) {
                        {  // the closest admissible new neighbour
                            const uint32_t clen = s_len, cworst = s_worst;
#pragma unroll 1
                            for (uint32_t j0 = 0; j0 < npend; j0 += 32) {
                                const uint32_t jj = j0 + lane;
                                const float dl = jj < npend ? pend_d[jj] : __int_as_float(0x7FC00000);
                                const uint32_t ol = order_bits(dl);
                                if (jj < npend && dl == dl && (clen < ef || ol < cworst)) pred_key = min(pred_key, ((uint64_t)ol << 32) | ((uint64_t)pend[jj] << 1));
                            }
                        }
                        const uint32_t bh = (uint32_t)(pred_key >> 32), mh = __reduce_min_sync(0xffffffffu, bh);
                        const uint32_t ml = __reduce_min_sync(0xffffffffu, bh == mh ? (uint32_t)pred_key : 0xFFFFFFFFu);
                        const uint32_t cn = (mh & ml) == 0xFFFFFFFFu ? 0xFFFFFFFFu : ml >> 1;
                        hs(12);
                        uint32_t np = 0xFFFFFFFFu;
                        if (cn != 0xFFFFFFFFu) {
                            const uint32_t* nb;
                            const uint16_t* dg;
                            uint32_t maxd;
                            if (level == 0) {
                                nb = p.g.nbr0 + (size_t)cn * p.g.max_m0;
                                dg = p.g.deg0 + cn;
                                maxd = p.g.max_m0;
                            } else {
                                const size_t slot = (size_t)p.g.upper_base[cn] + (size_t)(level - 1);
                                nb = p.g.nbrU + slot * p.g.M;
                                dg = p.g.degU + slot;
                                maxd = p.g.M;
                            }
                            const uint32_t first = (uint32_t)lane < maxd ? __ldg(nb + lane) : 0u;  // with the degree: one round trip
                            const uint32_t deg = *dg;
                            if (s_vcount + deg <= vlimit) {  // (else warp 0 meets the full table itself)
                                // exactly warp 0's expansion step (below): the visited set is this warp's alone until warp 0
                                // has finished the array
                                np = 0;
#pragma unroll 1
                                for (uint32_t i0 = 0; i0 < deg; i0 += 32) {
                                    const uint32_t i = i0 + lane;
                                    uint32_t v = 0;
                                    bool fresh = false;
                                    if (i < deg) {
                                        v = i0 == 0 ? first : __ldg(nb + i);
                                        fresh = hvis_insert_smem(vt, vmask, v);
                                    }
                                    const uint32_t m = __ballot_sync(0xffffffffu, fresh);
                                    if (fresh) {
                                        ppend[np + __popc(m & ((1u << lane) - 1u))] = v;
                                        if (!p.spec_rows) l2_prefetch_bulk(p.b_base + (uint64_t)v * p.b_stride, row_bytes);
                                    }
                                    np += __popc(m);
                                }
                            }
                        }
                        if (lane == 0) {
                            s_pnp = np;
                            s_pred_cn = np == 0xFFFFFFFFu ? 0xFFFFFFFFu : cn;
                            if (p.spec_rows && np >= 1u && np <= HC_ROWS) mbar_expect_tx(bar_ahead, np * row_bytes);
                        }
                        __syncwarp();
                        asm volatile("bar.arrive 1, 64;" ::: "memory");  // warp 0 may take the list from here on
                        hs(13);
                    }
                    if (p.spec_rows) {  // all four helper warps issue the copies (one per row; a lone warp would serialise them)
                        asm volatile("bar.sync 2, %0;" ::"n"(HC_THREADS - 32 * HC_MERGE_WARPS) : "memory");
                        const uint32_t np = *(volatile uint32_t*)&s_pnp;
                        const uint32_t r = (uint32_t)lane * (HC_THREADS / 32 - HC_MERGE_WARPS) + ((uint32_t)warp - HC_MERGE_WARPS);
                        if (np <= HC_ROWS && r < np)
                            bulk_g2s(smem_u32(srow_next + (size_t)r * p.units), p.b_base + (uint64_t)ppend[r] * p.b_stride, row_bytes, bar_ahead);
                        __syncwarp();
                        hs(14);
                    }
                }
                if (PROF && threadIdx.x == 0) s_pf[15] += (unsigned long long)(clock64() - s_t0);
                __syncthreads();
                if (PROF && threadIdx.x == 0) s_pf[16] += (unsigned long long)(clock64() - s_t0);
                if (s_done) break;
            }
            {  // copies fetched ahead for an expansion that never came: let them land before the buffer is used again
                const uint32_t pnp = s_pnp;
                if (p.spec_rows && pnp >= 1u && pnp <= HC_ROWS) {
                    mbar_wait(bar_ahead, par_ahead);
                    par_ahead ^= 1u;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            p.status[q] = s_status;
            atomicAdd(p.scored, nscored);
            if (p.hist)
#pragma unroll 1
                for (int b = 0; b < 5; ++b)
                    if (s_hb[b]) atomicAdd(p.hist + b, (unsigned long long)s_hb[b]);
            if (PROF && p.prof)
#pragma unroll 1
                for (int b = 0; b < 20; ++b) atomicAdd(p.prof + b, s_pf[b]);
        }
        __syncthreads();
    }
}

// ---- Vec0Tab::update / delete on the resident graph (src/vtab.rs:1340-1397, 1860-1895): every edge to `node` is removed
// (one thread per adjacency list, the rest of the list keeps its order) and its own lists are emptied.
__global__ void __launch_bounds__(256) hnsw_unlink_kernel(uint32_t* nbr, float* dist, uint16_t* deg, uint64_t n_lists, uint32_t width,
                                                          uint32_t node, uint64_t own_first, uint64_t own_count) {
    for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < n_lists; l += (uint64_t)gridDim.x * blockDim.x) {
        if (l >= own_first && l < own_first + own_count) {  // the node's own lists
            deg[l] = 0;
            continue;
        }
        uint32_t* nb = nbr + l * width;
        float* ds = dist + l * width;
        const uint32_t d = deg[l];
        uint32_t w = 0;
        for (uint32_t i = 0; i < d; ++i) {
            const uint32_t v = nb[i];
            if (v != node) {
                if (w != i) {
                    nb[w] = v;
                    ds[w] = ds[i];
                }
                ++w;
            }
        }
        if (w != d) deg[l] = (uint16_t)w;
    }
}

// ---- a row inserted OUT of rowid order moved every later row one position up: node ids are row positions, so every
// neighbour id >= first_moved goes up by one and the per-node arrays shift by one row from there.  The level-0 arrays are
// rewritten into a spare allocation of the same size (the two swap roles call after call: no allocation per insert), the
// renumbering fused into the copy of the ids; coalesced, one pass: 2 x (read + write) of the level-0 lists per insert.
// Ids behind a list's degree are renumbered too — they are never read.
template <typename T, bool RENUMBER>
__global__ void __launch_bounds__(256) hnsw_shift_rows_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t n_items,
                                                              uint32_t width, uint32_t first_moved) {
    const uint64_t hole = (uint64_t)first_moved * width;  // first item of the new (empty) row
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_items + width; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i >= hole && i < hole + width) {
            out[i] = 0;
            continue;
        }
        T v = in[i < hole ? i : i - width];
        if (RENUMBER && v >= (T)first_moved) v += 1;
        out[i] = v;
    }
}

// the upper lists are few (one node in M has any): renumbered in place, one thread per list
__global__ void __launch_bounds__(256) hnsw_renumber_kernel(uint32_t* nbr, const uint16_t* deg, uint64_t n_lists, uint32_t width,
                                                            uint32_t first_moved) {
    for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < n_lists; l += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t* nb = nbr + l * width;
        const uint32_t d = deg[l];
        for (uint32_t i = 0; i < d; ++i)
            if (nb[i] >= first_moved) nb[i] += 1;
    }
}

// ---- linking on the device (insert.rs:408-498 for a whole batch) ---------------------------------------------------
// The search kernel left, per (insert, layer), the sorted closest results in out_keys.  Forward edges: the new node's own
// list is exactly its first min(cnt, maxc) results, in order.  Reverse edges: every (neighbour, layer) list receives the
// new nodes in the order the sequential loop would add them (insert order, layers top-down, results closest first);
// the operations are keyed (list id, sequence number), radix-sorted, and one thread per list replays its run with the
// reference's rule: append while there is room, else replace the worst (distance, node) if the newcomer is better
// (simple_prune keeps the closest, src/hnsw/insert.rs:144-222).  Same lists, same order, same stored distances as the
// host loop (tests compare the exported graphs byte for byte).
struct HLinkParams {
    uint32_t* nbr0;
    float* dist0;
    uint16_t* deg0;
    const uint32_t* upper_base;
    uint32_t* nbrU;
    float* distU;
    uint16_t* degU;
    uint32_t max_m0, M;
    uint32_t n_rows;             // level-0 list id = node, upper list id = n_rows + slot
    const uint32_t* a_index;     // [nq] the new nodes
    const uint32_t* out_off;     // [nq] first result slot of each insert
    const int8_t* node_level;    // [nq]
    int32_t entry_level;         // entry level the batch searched from
    uint32_t nq, take;
    const uint64_t* out_keys;    // [slot][take]
    const uint32_t* out_cnt;     // [slot]
    uint64_t* op_key;            // [slots * take] (list id << 32) | sequence, ~0 = no operation
    uint64_t* op_val;            // [slots * take] (distance bits << 32) | new node
    uint32_t n_ops;
};

__device__ __forceinline__ void hlink_list(const HLinkParams& p, uint32_t lid, uint32_t** nb, float** dist, uint16_t** deg, uint32_t* maxc) {
    if (lid < p.n_rows) {
        *nb = p.nbr0 + (size_t)lid * p.max_m0;
        *dist = p.dist0 + (size_t)lid * p.max_m0;
        *deg = p.deg0 + lid;
        *maxc = p.max_m0;
    } else {
        const size_t slot = lid - p.n_rows;
        *nb = p.nbrU + slot * p.M;
        *dist = p.distU + slot * p.M;
        *deg = p.degU + slot;
        *maxc = p.M;
    }
}

// one warp per insert: write the new node's own lists and emit the reverse-edge operations
__global__ void __launch_bounds__(256) hnsw_link_forward_kernel(const HLinkParams p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = w; i < p.nq; i += nw) {
        const uint32_t node = p.a_index[i];
        const int nl = min((int)p.node_level[i], p.entry_level) + 1;
        for (int lv = 0; lv < nl; ++lv) {
            const uint32_t slot = p.out_off[i] + (uint32_t)lv;
            const uint32_t lid = lv == 0 ? node : p.n_rows + p.upper_base[node] + (uint32_t)(lv - 1);
            uint32_t* nb;
            float* dist;
            uint16_t* deg;
            uint32_t maxc;
            hlink_list(p, lid, &nb, &dist, &deg, &maxc);
            const uint32_t tk = min(p.out_cnt[slot], maxc);
            for (uint32_t j = lane; j < p.take; j += 32) {
                const size_t o = (size_t)slot * p.take + j;
                if (j < tk) {
                    const uint64_t key = p.out_keys[o];
                    const uint32_t wn = (uint32_t)(key & 0xFFFFFFFFull) >> 1;
                    const float d = order_bits_inv((uint32_t)(key >> 32));
                    nb[j] = wn;
                    dist[j] = d;
                    const uint32_t tl = lv == 0 ? wn : p.n_rows + p.upper_base[wn] + (uint32_t)(lv - 1);
                    const uint32_t seq = ((i * 16u + (uint32_t)(15 - lv)) << 8) | j;
                    p.op_key[o] = ((uint64_t)tl << 32) | seq;
                    p.op_val[o] = ((uint64_t)__float_as_uint(d) << 32) | node;
                } else {
                    p.op_key[o] = ~0ull;
                }
            }
            if (lane == 0) *deg = (uint16_t)tk;
        }
    }
}

// One warp per list: the warp that finds the head of a list's run in its 32-operation window replays the whole run in
// order with the list held in registers (entry e*32+lane in k[e] as (order_bits(distance) << 32) | node, so the worst
// entry is a warp-wide maximum).  Hub nodes receive thousands of reverse edges per batch: a run costs ~30 instructions
// per operation instead of a 32-entry scan of global memory.
__global__ void __launch_bounds__(256) hnsw_link_reverse_kernel(const HLinkParams p, const uint64_t* skey, const uint64_t* sval) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = w * 32; base < p.n_ops; base += nw * 32) {
        const uint32_t t = base + lane;
        bool head = false;
        if (t < p.n_ops) {
            const uint64_t key = skey[t];
            if (key != ~0ull) head = t == 0 || (uint32_t)(skey[t - 1] >> 32) != (uint32_t)(key >> 32);
        }
        uint32_t hm = __ballot_sync(0xffffffffu, head);
        while (hm) {
            const uint32_t t0 = base + (uint32_t)__ffs(hm) - 1u;
            hm &= hm - 1;
            const uint32_t lid = (uint32_t)(skey[t0] >> 32);
            uint32_t* nb;
            float* dist;
            uint16_t* deg;
            uint32_t maxc;
            hlink_list(p, lid, &nb, &dist, &deg, &maxc);
            uint32_t d = *deg;
            uint64_t k[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t idx = (uint32_t)e * 32 + lane;
                k[e] = idx < d ? (((uint64_t)order_bits(dist[idx]) << 32) | nb[idx]) : 0ull;
            }
            for (uint32_t u = t0; u < p.n_ops; ++u) {
                const uint64_t ku = skey[u];
                if (ku == ~0ull || (uint32_t)(ku >> 32) != lid) break;
                const uint64_t v = sval[u];
                const uint64_t nk = ((uint64_t)order_bits(__uint_as_float((uint32_t)(v >> 32))) << 32) | (uint32_t)v;
                if (d < maxc) {  // room: append (insert.rs: no pruning needed)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if ((uint32_t)e == (d >> 5) && lane == (d & 31)) k[e] = nk;
                    ++d;
                } else {         // full: the newcomer replaces the worst (distance, node) if it is better
                    uint64_t mx = k[0];
#pragma unroll
                    for (int e = 1; e < 8; ++e) mx = k[e] > mx ? k[e] : mx;
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) {
                        const uint64_t o = shfl_xor_u64(mx, m);
                        mx = o > mx ? o : mx;
                    }
                    if (nk < mx) {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (k[e] == mx) k[e] = nk;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t idx = (uint32_t)e * 32 + lane;
                if (idx < d) {
                    nb[idx] = (uint32_t)k[e];
                    dist[idx] = order_bits_inv((uint32_t)(k[e] >> 32));
                }
            }
            if (lane == 0) *deg = (uint16_t)d;
            __syncwarp();
        }
    }
}

// scatter staged adjacency lists into the device graph: item = [list id][degree][width x u32]
__global__ void hnsw_scatter_kernel(const uint32_t* staged, uint32_t n_items, uint32_t width, uint32_t* nbr, uint16_t* deg) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t it = w; it < n_items; it += nw) {
        const uint32_t* src = staged + (size_t)it * (width + 2);
        const uint32_t id = src[0];
        if (lane == 0) deg[id] = (uint16_t)src[1];
        for (uint32_t i = lane; i < width; i += 32) nbr[(size_t)id * width + i] = src[2 + i];
    }
}

}  // namespace vg
