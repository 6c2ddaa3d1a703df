// vecgpu.cu — C ABI (include/vecgpu.h) + slab management + launch planning.
// Host side of libvecgpu.so.  See kernels.cuh for the device code.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/vecgpu.h"
#include <omp.h>
#include <unistd.h>

#include "kernels.cuh"
#include "xchg.cuh"
#include "tc_batch.cuh"
#include "hnsw_dev.cuh"

using namespace vg;

// ---------------------------------------------------------------------------
// errors: Result<T, Error> of src/error.rs:5-36 becomes (code, thread-local text)
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_tc_queries{0}, g_tc_fallbacks{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(VECGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                         \
    } while (0)

#define LAUNCHED()                                                                                               \
    do {                                                                                                         \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                                      \
        cudaError_t e_ = cudaGetLastError();                                                                     \
        if (e_ != cudaSuccess)                                                                                   \
            return fail(VECGPU_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                               \
    } while (0)

// Nothing may unwind across the C boundary (the header promises it; the caller is Rust, cf. catch_unwind at src/lib.rs:149):
// every extern "C" entry point that can allocate runs inside VG_TRY / VG_CATCH and turns an exception into status 4.
static int vg_caught() noexcept {
    try {
        throw;
    } catch (const std::bad_alloc&) {
        return fail(VECGPU_ERR_CUDA, "out of host memory");
    } catch (const std::exception& e) {
        return fail(VECGPU_ERR_CUDA, "internal error: %s", e.what());
    } catch (...) {
        return fail(VECGPU_ERR_CUDA, "internal error (unknown exception)");
    }
}
#define VG_TRY try {
#define VG_CATCH \
    }            \
    catch (...) { return vg_caught(); }
#define VG_CATCH_VOID \
    }                 \
    catch (...) { vg_caught(); }

static constexpr uint32_t K_FUSED_MAX = 1024;     // above this knn uses emit + radix sort
static constexpr size_t DIRECT_OUT_MAX = 4096;    // result entries (nq*k) up to which kernels write straight into mapped pinned memory
static constexpr uint32_t DIMS_MAX = 65536;       // keeps int8 partial sums inside int32
static constexpr size_t SMEM_MAX = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100
static constexpr size_t SCAN_SMEM_MAX = SMEM_MAX - 1024;  // scan_kernel also has a few static __shared__ words (fused tail)

static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return (uint32_t)strtoul(v, nullptr, 10);
}

// ---------------------------------------------------------------------------
// slab
// ---------------------------------------------------------------------------
struct vecgpu_slab {
    int elem = 0, device = 0, num_sms = 148;
    uint32_t dims = 0, row_bytes = 0, row_stride = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // rows, ascending rowid order
    uint8_t* d_vec = nullptr;
    uint64_t cap = 0, rows = 0;
    uint64_t layout_gen = 0;  // bumped whenever row positions move (out-of-order insert, compaction, reload): position-based indexes go stale
    // rowids: dense (first_rowid + position) until a gap appears, then explicit arrays
    bool dense = true;
    int64_t first_rowid = 1;
    std::vector<int64_t> h_rowids;
    int64_t* d_rowids = nullptr;
    uint64_t cap_rowids = 0, rowids_synced = 0;  // d_rowids[0, rowids_synced) == h_rowids[0, rowids_synced)
    // skip flags (tombstones / wrong-length blobs), allocated on first use
    std::vector<uint8_t> h_skip;
    uint8_t* d_skip = nullptr;
    uint64_t cap_skip = 0, n_skip = 0, skip_synced = 0;  // d_skip[0, skip_synced) == h_skip[0, skip_synced)
    // canonical |row|^2 cache for the tensor-core batched path (invalidated by any vector write)
    float* d_norms = nullptr;
    uint32_t* d_x2max = nullptr;
    uint32_t* d_unsafe = nullptr;   // [1 + TC_MAX_UNSAFE]: count + positions of rows with unusable norms
    uint32_t n_unsafe = 0;
    uint64_t cap_norms = 0;
    bool norms_valid = false;
    // workspaces
    void* d_ws[27] = {nullptr};
    size_t ws_cap[27] = {0};
    void* h_pin[2] = {nullptr};
    size_t pin_cap[2] = {0};
    // bulk-load staging: two pinned buffers so the host-side copy of chunk i+1 overlaps the DMA of chunk i
    void* h_stage[2] = {nullptr};
    size_t stage_cap[2] = {0};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // sharded queries: a push the next fused scan tail should carry (set by shard_knn_locked around knn_core)
    const XPushParams* fuse_push = nullptr;
    bool fuse_push_done = false;
    // host-API queries: staged in pinned memory, uploaded only if the chosen path needs them in device memory (a single
    // query for the streaming scan rides in the kernel parameters instead)
    const uint8_t* h_q_staged = nullptr;
    size_t h_q_bytes = 0;
    bool q_uploaded = true;
    // Device-API callers may pipeline independent single-query scans on TWO streams (the tail of one scan — stragglers,
    // list merges — then overlaps the start of the next): the streaming-scan path has a second set of scratch buffers.
    // slot_stream[i] = the stream that currently owns scratch set i.
    cudaStream_t slot_stream[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    cudaEvent_t slot_ev = nullptr;
};

enum { WS_QUERY = 0, WS_PART = 1, WS_OUT_ROWID = 2, WS_OUT_DIST = 3, WS_OUT_CNT = 4, WS_TMP = 5, WS_TMP2 = 6, WS_TMP3 = 7,
       WS_TC_CANDV = 8, WS_TC_CANDR = 9, WS_TC_CNT = 10, WS_TC_TAU = 11, WS_TC_PAIRQ = 12, WS_TC_PAIRPOS = 13, WS_TC_DIST = 14,
       WS_TC_KEYS = 15, WS_TC_QNORM = 16, WS_TC_FLAGS = 17, WS_TC_LOCK = 18, WS_TC_BUF = 19, WS_X_ROWID = 20, WS_X_DIST = 21, WS_X_CNT = 22,
       WS_TICKET = 23, WS_PART1 = 24, WS_TICKET1 = 25, WS_QUERY1 = 26, WS_COUNT = 27 };

static int ws_reserve(vecgpu_slab* s, int i, size_t bytes) {
    if (bytes <= s->ws_cap[i]) return 0;
    if (s->d_ws[i]) CU(cudaFree(s->d_ws[i]));
    s->d_ws[i] = nullptr;
    s->ws_cap[i] = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CU(cudaMalloc(&s->d_ws[i], want));
    s->ws_cap[i] = want;
    return 0;
}
static int pin_reserve(vecgpu_slab* s, int i, size_t bytes) {
    if (bytes <= s->pin_cap[i]) return 0;
    if (s->h_pin[i]) CU(cudaFreeHost(s->h_pin[i]));
    s->h_pin[i] = nullptr;
    s->pin_cap[i] = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CU(cudaMallocHost(&s->h_pin[i], want));
    s->pin_cap[i] = want;
    return 0;
}

static int grow_dev(void** p, uint64_t* cap_items, uint64_t used_items, uint64_t want_items, size_t item_bytes,
                    cudaStream_t st) {
    if (want_items <= *cap_items) return 0;
    void* np = nullptr;
    CU(cudaMalloc(&np, (size_t)want_items * item_bytes));
    if (*p && used_items) CU(cudaMemcpyAsync(np, *p, (size_t)used_items * item_bytes, cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    if (*p) CU(cudaFree(*p));
    *p = np;
    *cap_items = want_items;
    return 0;
}

static int slab_reserve_rows(vecgpu_slab* s, uint64_t want) {
    if (want <= s->cap) return 0;
    if (want >= 0xFFFFFFFFull) return fail(VECGPU_ERR_INVALID_PARAM, "a slab holds at most 2^32-2 rows");
    uint64_t ncap = std::max<uint64_t>(want, s->cap + s->cap / 2);
    if (s->cap == 0) ncap = want;
    int rc = grow_dev((void**)&s->d_vec, &s->cap, s->rows, ncap, s->row_stride, s->stream);
    if (rc) return rc;
    return 0;
}

// Upload the part of the host rowid mirror the device does not have yet (non-dense mode).  Appends cost O(appended rows):
// only [rowids_synced, size) travels; the arrays grow geometrically with the slab.  Whoever rewrites entries below
// rowids_synced (out-of-order insert, compaction, reload) lowers rowids_synced first.
static int slab_sync_rowids(vecgpu_slab* s) {
    if (s->dense) return 0;
    const uint64_t n = s->h_rowids.size();
    if (n > s->cap_rowids) {
        const uint64_t ncap = std::max<uint64_t>(std::max<uint64_t>(n, s->cap), s->cap_rowids + s->cap_rowids / 2);
        int64_t* np = nullptr;
        CU(cudaMalloc((void**)&np, ncap * sizeof(int64_t)));
        if (s->d_rowids && s->rowids_synced) {
            cudaError_t e = cudaMemcpy(np, s->d_rowids, s->rowids_synced * sizeof(int64_t), cudaMemcpyDeviceToDevice);
            if (e != cudaSuccess) {
                cudaFree(np);
                return fail(VECGPU_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
            }
        }
        if (s->d_rowids) cudaFree(s->d_rowids);
        s->d_rowids = np;
        s->cap_rowids = ncap;
    }
    if (s->rowids_synced > n) s->rowids_synced = n;
    if (s->rowids_synced < n) {
        CU(cudaMemcpy(s->d_rowids + s->rowids_synced, s->h_rowids.data() + s->rowids_synced,
                      (n - s->rowids_synced) * sizeof(int64_t), cudaMemcpyHostToDevice));
        s->rowids_synced = n;
    }
    return 0;
}

// Same for the skip flags.  The device copy is only meaningful below skip_synced: every path that clears or recreates
// h_skip (load, fill_synthetic, compaction) resets skip_synced to 0, so stale tombstones can never be read by a later
// scan (they used to survive a reload: the next single-byte update found cap_skip >= rows and uploaded one byte only).
static int slab_sync_skip(vecgpu_slab* s) {
    if (s->h_skip.empty()) {
        s->skip_synced = 0;
        return 0;
    }
    const uint64_t n = s->h_skip.size();
    if (n > s->cap_skip) {
        const uint64_t ncap = std::max<uint64_t>(std::max<uint64_t>(n, s->cap), s->cap_skip + s->cap_skip / 2);
        uint8_t* np = nullptr;
        CU(cudaMalloc((void**)&np, ncap));
        if (s->d_skip && s->skip_synced) {
            cudaError_t e = cudaMemcpy(np, s->d_skip, s->skip_synced, cudaMemcpyDeviceToDevice);
            if (e != cudaSuccess) {
                cudaFree(np);
                return fail(VECGPU_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
            }
        }
        if (s->d_skip) cudaFree(s->d_skip);
        s->d_skip = np;
        s->cap_skip = ncap;
    }
    if (s->skip_synced > n) s->skip_synced = n;
    if (s->skip_synced < n) {
        CU(cudaMemcpy(s->d_skip + s->skip_synced, s->h_skip.data() + s->skip_synced, n - s->skip_synced, cudaMemcpyHostToDevice));
        s->skip_synced = n;
    }
    return 0;
}

// forget every row: host mirrors AND what the device copies are known to hold
static void slab_reset_rows(vecgpu_slab* s) {
    s->rows = 0;
    s->dense = true;
    s->first_rowid = 1;
    s->h_rowids.clear();
    s->rowids_synced = 0;
    s->h_skip.clear();
    s->n_skip = 0;
    s->skip_synced = 0;
    s->norms_valid = false;
    ++s->layout_gen;  // every row position is replaced: position-based indexes (HNSW) go stale
}

static void slab_materialize_rowids(vecgpu_slab* s) {
    if (!s->dense) return;
    s->h_rowids.resize(s->rows);
    for (uint64_t i = 0; i < s->rows; ++i) s->h_rowids[i] = s->first_rowid + (int64_t)i;
    s->rowids_synced = 0;
    s->dense = false;
}

static int64_t slab_last_rowid(const vecgpu_slab* s) {
    if (s->rows == 0) return INT64_MIN;
    return s->dense ? s->first_rowid + (int64_t)s->rows - 1 : s->h_rowids.back();
}

// position of rowid, or -1; *ins = insertion point when absent
static int64_t slab_find(const vecgpu_slab* s, int64_t rowid, uint64_t* ins) {
    if (s->dense) {
        if (s->rows && rowid >= s->first_rowid && (uint64_t)(rowid - s->first_rowid) < s->rows)
            return rowid - s->first_rowid;
        if (ins) *ins = (s->rows == 0 || rowid > slab_last_rowid(s)) ? s->rows : 0;
        return -1;
    }
    auto it = std::lower_bound(s->h_rowids.begin(), s->h_rowids.end(), rowid);
    if (ins) *ins = (uint64_t)(it - s->h_rowids.begin());
    if (it != s->h_rowids.end() && *it == rowid) return (int64_t)(it - s->h_rowids.begin());
    return -1;
}

static int slab_set_skip(vecgpu_slab* s, uint64_t pos, uint8_t v) {
    if (s->h_skip.size() < s->rows) s->h_skip.resize(s->rows, 0);
    if (s->h_skip[pos] == v) return slab_sync_skip(s);
    s->h_skip[pos] = v;
    if (v) ++s->n_skip; else --s->n_skip;
    if (pos >= s->skip_synced) return slab_sync_skip(s);  // the tail upload carries it
    int rc = slab_sync_skip(s);
    if (rc) return rc;
    CU(cudaMemcpy(s->d_skip + pos, &s->h_skip[pos], 1, cudaMemcpyHostToDevice));
    return 0;
}

static int slab_update_norms(vecgpu_slab* s, uint64_t pos, uint64_t n);

static int query_upload_if_needed(vecgpu_slab* s, cudaStream_t st);

static constexpr size_t STAGE_BYTES = 16u << 20;  // per staging buffer

// copy n host rows (row_bytes each) into slab rows [pos, pos+n): through two pinned staging buffers, chunk by chunk, with
// asynchronous copies on the slab's stream (the host-side memcpy of the next chunk runs while the previous one is on the
// bus); returns after the last byte has landed.  The cached row norms of the tensor-core paths are updated for the written
// rows only (they used to be thrown away by any write, so a single upsert made the next batched query re-read the slab).
static int slab_write_rows(vecgpu_slab* s, uint64_t pos, const void* vectors, uint64_t n) {
    if (n == 0) return 0;
    const uint64_t chunk_rows = std::max<uint64_t>(1, STAGE_BYTES / s->row_bytes);
    const size_t want = (size_t)std::min<uint64_t>(n, chunk_rows) * s->row_bytes;
    const int nbuf = n > chunk_rows ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
        if (s->stage_cap[b] < want) {
            if (s->h_stage[b]) CU(cudaFreeHost(s->h_stage[b]));
            s->h_stage[b] = nullptr;
            s->stage_cap[b] = 0;
            const size_t cap = n > chunk_rows ? std::max(want, STAGE_BYTES) : std::max(want, (size_t)65536);
            CU(cudaMallocHost(&s->h_stage[b], cap));
            s->stage_cap[b] = cap;
        }
        if (!s->stage_ev[b]) CU(cudaEventCreateWithFlags(&s->stage_ev[b], cudaEventDisableTiming));
    }
    const uint8_t* src = (const uint8_t*)vectors;
    uint8_t* dst = s->d_vec + pos * s->row_stride;
    uint64_t i = 0;
    for (uint64_t off = 0; off < n; off += chunk_rows, ++i) {
        const int b = (int)(i & 1);
        const uint64_t m = std::min(chunk_rows, n - off);
        if (i >= 2) CU(cudaEventSynchronize(s->stage_ev[b]));  // the DMA out of this buffer two chunks ago is done
        const size_t bytes = (size_t)m * s->row_bytes;
        if (bytes >= (4u << 20)) {
            const int parts = 8;
#pragma omp parallel for schedule(static) num_threads(parts)
            for (int t = 0; t < parts; ++t) {
                const size_t lo = bytes * t / parts, hi = bytes * (t + 1) / parts;
                memcpy((uint8_t*)s->h_stage[b] + lo, src + off * s->row_bytes + lo, hi - lo);
            }
        } else {
            memcpy(s->h_stage[b], src + off * s->row_bytes, bytes);
        }
        if (s->row_bytes == s->row_stride) {
            CU(cudaMemcpyAsync(dst + off * s->row_stride, s->h_stage[b], bytes, cudaMemcpyHostToDevice, s->stream));
        } else {
            CU(cudaMemsetAsync(dst + off * s->row_stride, 0, (size_t)m * s->row_stride, s->stream));
            CU(cudaMemcpy2DAsync(dst + off * s->row_stride, s->row_stride, s->h_stage[b], s->row_bytes, s->row_bytes, (size_t)m,
                                 cudaMemcpyHostToDevice, s->stream));
        }
        CU(cudaEventRecord(s->stage_ev[b], s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    return slab_update_norms(s, pos, n);
}

extern "C" {

const char* vecgpu_last_error(void) { return g_err; }
const char* vecgpu_version(void) { return "vecgpu 0.1.0 sm_100a"; }
uint64_t vecgpu_launch_count(void) { return g_launches.load(); }
void vecgpu_tc_stats(uint64_t* queries, uint64_t* fallbacks) {
    if (queries) *queries = g_tc_queries.load();
    if (fallbacks) *fallbacks = g_tc_fallbacks.load();
}

int vecgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

uint32_t vecgpu_row_bytes(int elem, uint32_t dims) {
    switch (elem) {
        case VECGPU_F32: return dims * 4u;
        case VECGPU_I8: return dims;
        case VECGPU_BIT: return (dims + 7u) / 8u;
        default: return 0;
    }
}

int vecgpu_metric_supported(int elem, int metric) {
    if (elem == VECGPU_F32 || elem == VECGPU_I8)
        return metric == VECGPU_L2 || metric == VECGPU_L1 || metric == VECGPU_COSINE;
    if (elem == VECGPU_BIT) return metric == VECGPU_HAMMING;
    return 0;
}

}  // extern "C"

// Developer hook (not part of the reference surface): a device buffer of (num_sms * 4 + 1) u64 that the next single-query
// scans fill with globaltimer stamps per CTA; NULL switches it off.  tools/scan_timeline.py reads it.
static unsigned long long* g_scan_dbg = nullptr;
extern "C" void vecgpu_debug_scan_timeline(void* d_buf) { g_scan_dbg = (unsigned long long*)d_buf; }

static int check_pair(int elem, int metric) {
    if (elem < 0 || elem > 2) return fail(VECGPU_ERR_UNSUPPORTED, "invalid vector type %d", elem);
    if (metric < 0 || metric > 3) return fail(VECGPU_ERR_INVALID_PARAM, "invalid distance metric %d", metric);
    if (!vecgpu_metric_supported(elem, metric)) {
        static const char* mn[] = {"L2", "L1", "Cosine", "Hamming"};
        static const char* en[] = {"Float32", "Int8", "Bit"};
        // wording follows src/distance/mod.rs:78-82
        return fail(VECGPU_ERR_UNSUPPORTED, "Distance metric %s not supported for vector type %s", mn[metric], en[elem]);
    }
    return 0;
}

static int use_device(int device) {
    int n = vecgpu_device_count();
    if (n <= 0) return fail(VECGPU_ERR_CUDA, "no CUDA device is available (libvecgpu has no CPU fallback)");
    if (device < 0 || device >= n) return fail(VECGPU_ERR_INVALID_PARAM, "device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    return 0;
}

extern "C" int vecgpu_slab_create(int elem, uint32_t dims, uint64_t capacity_hint, int device, vecgpu_slab** out) {
    VG_TRY
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "out is NULL");
    *out = nullptr;
    if (elem < 0 || elem > 2) return fail(VECGPU_ERR_UNSUPPORTED, "invalid vector type %d", elem);
    if (dims == 0 || dims > DIMS_MAX) return fail(VECGPU_ERR_INVALID_PARAM, "dims must be in 1..%u", DIMS_MAX);
    int rc = use_device(device);
    if (rc) return rc;
    vecgpu_slab* s = new (std::nothrow) vecgpu_slab();
    if (!s) return fail(VECGPU_ERR_CUDA, "out of host memory");
    s->elem = elem;
    s->device = device;
    s->dims = dims;
    s->row_bytes = vecgpu_row_bytes(elem, dims);
    s->row_stride = (s->row_bytes + 15u) & ~15u;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete s;
        return fail(VECGPU_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (capacity_hint) {
        rc = slab_reserve_rows(s, capacity_hint);
        if (rc) {
            cudaStreamDestroy(s->stream);
            delete s;
            return rc;
        }
    }
    *out = s;
    return 0;
    VG_CATCH
}

extern "C" void vecgpu_slab_destroy(vecgpu_slab* s) {
    VG_TRY
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_vec);
    cudaFree(s->d_rowids);
    cudaFree(s->d_skip);
    for (int i = 0; i < WS_COUNT; ++i) cudaFree(s->d_ws[i]);
    cudaFree(s->d_norms);
    cudaFree(s->d_x2max);
    cudaFree(s->d_unsafe);
    for (int i = 0; i < 2; ++i) {
        if (s->h_pin[i]) cudaFreeHost(s->h_pin[i]);
        if (s->h_stage[i]) cudaFreeHost(s->h_stage[i]);
        if (s->stage_ev[i]) cudaEventDestroy(s->stage_ev[i]);
    }
    if (s->slot_ev) cudaEventDestroy(s->slot_ev);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
    VG_CATCH_VOID
}

static int slab_append_locked(vecgpu_slab* s, const int64_t* rowids, const void* vectors, uint64_t n) {
    if (n == 0) return 0;
    if (!vectors) return fail(VECGPU_ERR_INVALID_PARAM, "vectors is NULL");
    const int64_t last = slab_last_rowid(s);
    if (rowids) {
        if (s->rows && rowids[0] <= last)
            return fail(VECGPU_ERR_INVALID_PARAM, "appended rowids must be greater than the slab's last rowid");
        for (uint64_t i = 1; i < n; ++i)
            if (rowids[i] <= rowids[i - 1]) return fail(VECGPU_ERR_INVALID_PARAM, "rowids must be strictly ascending");
    } else if (s->rows && last == INT64_MAX) {
        return fail(VECGPU_ERR_INVALID_PARAM, "rowid overflow");
    }
    int rc = slab_reserve_rows(s, s->rows + n);
    if (rc) return rc;
    // the device write comes first: a failure leaves rows / rowids / skip flags exactly as they were
    rc = slab_write_rows(s, s->rows, vectors, n);
    if (rc) return rc;
    // still dense?
    bool stays_dense = s->dense;
    if (s->dense && rowids) {
        const int64_t expect0 = s->rows ? last + 1 : rowids[0];
        for (uint64_t i = 0; i < n && stays_dense; ++i) stays_dense = rowids[i] == expect0 + (int64_t)i;
    }
    const uint64_t old_rows = s->rows;
    const bool was_dense = s->dense;
    const int64_t old_first = s->first_rowid;
    const size_t old_skip = s->h_skip.size();
    if (s->dense && !stays_dense) slab_materialize_rowids(s);
    if (s->dense) {
        if (s->rows == 0) s->first_rowid = rowids ? rowids[0] : 1;
    } else {
        const int64_t start = s->rows ? last + 1 : 1;
        for (uint64_t i = 0; i < n; ++i) s->h_rowids.push_back(rowids ? rowids[i] : start + (int64_t)i);
    }
    s->rows += n;
    if (!s->h_skip.empty()) s->h_skip.resize(s->rows, 0);
    rc = slab_sync_skip(s);
    if (!rc) rc = slab_sync_rowids(s);
    if (rc) {  // roll the host state back to what the device arrays still describe
        s->rows = old_rows;
        s->h_skip.resize(old_skip);
        s->skip_synced = std::min<uint64_t>(s->skip_synced, old_skip);
        if (was_dense) {
            s->dense = true;
            s->first_rowid = old_first;
            s->h_rowids.clear();
            s->rowids_synced = 0;
        } else {
            s->h_rowids.resize(old_rows);
            s->rowids_synced = std::min<uint64_t>(s->rowids_synced, old_rows);
        }
        return rc;
    }
    return 0;
}

extern "C" int vecgpu_slab_append(vecgpu_slab* s, const int64_t* rowids, const void* vectors, uint64_t n) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    return slab_append_locked(s, rowids, vectors, n);
    VG_CATCH
}

extern "C" int vecgpu_slab_load(vecgpu_slab* s, const int64_t* rowids, const void* vectors, uint64_t n) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    slab_reset_rows(s);  // the previous contents (and every row position) are replaced
    return slab_append_locked(s, rowids, vectors, n);
    VG_CATCH
}

extern "C" int vecgpu_slab_upsert(vecgpu_slab* s, int64_t rowid, const void* vec, uint32_t nbytes) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    const bool good = vec != nullptr && nbytes == s->row_bytes;
    uint64_t ins = 0;
    int64_t pos = slab_find(s, rowid, &ins);
    if (pos >= 0) {
        if (good) {
            rc = slab_write_rows(s, (uint64_t)pos, vec, 1);
            if (rc) return rc;
            if (!s->h_skip.empty()) return slab_set_skip(s, (uint64_t)pos, 0);
            return 0;
        }
        return slab_set_skip(s, (uint64_t)pos, 1);
    }
    std::vector<uint8_t> zero;
    const void* src = vec;
    if (!good) {
        zero.assign(s->row_bytes, 0);
        src = zero.data();
    }
    if (s->rows == 0 || rowid > slab_last_rowid(s)) {
        rc = slab_append_locked(s, &rowid, src, 1);
        if (rc) return rc;
        if (!good) return slab_set_skip(s, s->rows - 1, 1);
        return 0;
    }
    // out-of-order insert (rare: explicit rowid below MAX): the tail moves up by one row ON THE DEVICE — vectors, rowids and
    // skip flags alike — and only the new element is uploaded (it used to re-upload both whole mirrors: 90 MB per insert at
    // 10 M rows).  The host mirrors shift with one memmove each.
    slab_materialize_rowids(s);
    if ((rc = slab_sync_rowids(s))) return rc;
    if ((rc = slab_sync_skip(s))) return rc;
    slab_find(s, rowid, &ins);
    rc = slab_reserve_rows(s, s->rows + 1);
    if (rc) return rc;
    const uint64_t tail = s->rows - ins;
    const bool has_skip = !s->h_skip.empty();
    // make room for one more element in the device mirrors (both are fully synced here; growing preserves the contents)
    s->h_rowids.push_back(0);
    if (has_skip) s->h_skip.push_back(0);
    rc = slab_sync_rowids(s);
    if (!rc && has_skip) rc = slab_sync_skip(s);
    s->h_rowids.pop_back();
    if (has_skip) s->h_skip.pop_back();
    s->rowids_synced = std::min<uint64_t>(s->rowids_synced, s->rows);
    s->skip_synced = std::min<uint64_t>(s->skip_synced, has_skip ? s->rows : 0);
    if (rc) return rc;
    if (tail) {
        // [ins, rows) moves up by one item, from the end backwards, in chunks through a bounded scratch buffer (the whole tail
        // at once would need a second copy of up to the whole slab next to it)
        const size_t chunk = (size_t)std::max(1u, env_u32("VECGPU_SHIFT_CHUNK_MB", 256)) << 20;
        rc = ws_reserve(s, WS_TMP, std::min(chunk, std::max((size_t)tail * s->row_stride, (size_t)tail * 8)));
        if (rc) return rc;
        auto shift_up = [&](uint8_t* base, size_t item) -> int {
            const size_t bytes = (size_t)tail * item;
            uint8_t* first = base + ins * item;
            for (size_t done = 0; done < bytes;) {
                const size_t m = std::min(chunk, bytes - done), off = bytes - done - m;
                CU(cudaMemcpyAsync(s->d_ws[WS_TMP], first + off, m, cudaMemcpyDeviceToDevice, s->stream));
                CU(cudaMemcpyAsync(first + off + item, s->d_ws[WS_TMP], m, cudaMemcpyDeviceToDevice, s->stream));
                done += m;
            }
            return 0;
        };
        // (allocations first: nothing has moved yet if one fails)
        if (s->norms_valid && s->elem != VECGPU_BIT && s->rows + 1 > s->cap_norms) {
            const uint64_t cap = std::max<uint64_t>(s->rows + 1, s->cap);
            float* nn = nullptr;
            CU(cudaMalloc((void**)&nn, cap * sizeof(float)));
            CU(cudaMemcpyAsync(nn, s->d_norms, s->cap_norms * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            cudaFree(s->d_norms);
            s->d_norms = nn;
            s->cap_norms = cap;
        }
        if ((rc = shift_up(s->d_vec, s->row_stride))) return rc;
        if ((rc = shift_up((uint8_t*)s->d_rowids, 8))) return rc;
        if (has_skip && (rc = shift_up(s->d_skip, 1))) return rc;
        // the cached |row|^2 are position-indexed too: they move with the rows (and the few positions on the always-re-ranked
        // list are renumbered) instead of being discarded — the next batched query would recompute all of them
        if (s->norms_valid && s->elem != VECGPU_BIT) {
            if ((rc = shift_up((uint8_t*)s->d_norms, 4))) return rc;
            if (s->elem == VECGPU_F32 && s->d_unsafe) {
                renumber_positions_kernel<<<1, 256, 0, s->stream>>>(s->d_unsafe, TC_MAX_UNSAFE, (uint32_t)ins);
                LAUNCHED();
            }
        } else {
            s->norms_valid = false;
        }
        CU(cudaStreamSynchronize(s->stream));
    }
    ++s->layout_gen;
    s->h_rowids.insert(s->h_rowids.begin() + (ptrdiff_t)ins, rowid);
    if (has_skip) s->h_skip.insert(s->h_skip.begin() + (ptrdiff_t)ins, 0);
    s->rows += 1;
    const uint8_t zero_flag = 0;
    CU(cudaMemcpy(s->d_rowids + ins, &rowid, 8, cudaMemcpyHostToDevice));
    if (has_skip) CU(cudaMemcpy(s->d_skip + ins, &zero_flag, 1, cudaMemcpyHostToDevice));
    s->rowids_synced = s->rows;
    if (has_skip) s->skip_synced = s->rows;
    rc = slab_write_rows(s, ins, src, 1);
    if (rc) return rc;
    if (!good) return slab_set_skip(s, ins, 1);
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_slab_delete(vecgpu_slab* s, int64_t rowid) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    int64_t pos = slab_find(s, rowid, nullptr);
    if (pos < 0) return 0;
    return slab_set_skip(s, (uint64_t)pos, 1);
    VG_CATCH
}

// Drop the skipped rows (tombstones of vecgpu_slab_delete, wrong-length blobs) physically: the kept rows are gathered, in
// order, into a fresh allocation that replaces the old one (needs kept_rows * row_stride bytes next to the slab for the
// duration of the call).  Scans never read skipped rows' flags unless a candidate passes, but they do stream their bytes;
// after many deletes a compaction gives that bandwidth back.
extern "C" int vecgpu_slab_compact(vecgpu_slab* s, uint64_t* removed) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    if (removed) *removed = 0;
    if (s->n_skip == 0) return 0;
    int rc = use_device(s->device);
    if (rc) return rc;
    slab_materialize_rowids(s);
    std::vector<uint32_t> keep;
    keep.reserve(s->rows - s->n_skip);
    for (uint64_t p = 0; p < s->rows; ++p)
        if (!s->h_skip[p]) keep.push_back((uint32_t)p);
    const uint64_t kept = keep.size();
    struct DevBuf {  // the new allocation is released if anything fails before it replaces the old one
        uint8_t* p = nullptr;
        ~DevBuf() { if (p) cudaFree(p); }
    } nb;
    const uint64_t ncap = std::max<uint64_t>(kept, 1);
    CU(cudaMalloc((void**)&nb.p, (size_t)ncap * s->row_stride));
    const uint64_t chunk = 1u << 22;  // positions uploaded per gather launch
    for (uint64_t off = 0; off < kept; off += chunk) {
        const uint64_t m = std::min(chunk, kept - off);
        if ((rc = ws_reserve(s, WS_TMP, (size_t)m * 4))) return rc;
        CU(cudaMemcpyAsync(s->d_ws[WS_TMP], keep.data() + off, (size_t)m * 4, cudaMemcpyHostToDevice, s->stream));
        gather_rows_kernel<<<(uint32_t)std::min<uint64_t>((m + 7) / 8, (uint64_t)s->num_sms * 16), 256, 0, s->stream>>>(
            s->d_vec, (const uint32_t*)s->d_ws[WS_TMP], m, s->row_stride / 16, nb.p + off * s->row_stride);
        LAUNCHED();
        CU(cudaStreamSynchronize(s->stream));
    }
    CU(cudaFree(s->d_vec));
    s->d_vec = nb.p;
    nb.p = nullptr;
    s->cap = ncap;
    for (uint64_t i = 0; i < kept; ++i) s->h_rowids[i] = s->h_rowids[keep[i]];
    s->h_rowids.resize(kept);
    if (removed) *removed = s->rows - kept;
    s->rows = kept;
    s->h_skip.assign(kept, 0);
    s->n_skip = 0;
    s->norms_valid = false;
    ++s->layout_gen;
    s->rowids_synced = 0;
    s->skip_synced = 0;
    if ((rc = slab_sync_rowids(s))) return rc;
    return slab_sync_skip(s);
    VG_CATCH
}

extern "C" int vecgpu_slab_count(vecgpu_slab* s, uint64_t* rows, uint64_t* live) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    if (rows) *rows = s->rows;
    if (live) *live = s->rows - s->n_skip;
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_slab_get(vecgpu_slab* s, int64_t rowid, void* out_vec, int* found) {
    VG_TRY
    if (!s || !found) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    *found = 0;
    int64_t pos = slab_find(s, rowid, nullptr);
    if (pos < 0 || (!s->h_skip.empty() && s->h_skip[(size_t)pos])) return 0;
    if (out_vec) CU(cudaMemcpy(out_vec, s->d_vec + (uint64_t)pos * s->row_stride, s->row_bytes, cudaMemcpyDeviceToHost));
    *found = 1;
    return 0;
    VG_CATCH
}

// ---------------------------------------------------------------------------
// scan planning
// ---------------------------------------------------------------------------
struct ScanCfg {
    uint32_t C, QB, R, CB, n_chunks, srs, S /* ring depth per consumer warp */, contig;
    size_t smem;
};

static uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// per-query key slots for the C per-warp lists (query-major, padded to a power of two for the in-CTA bitonic merge)
static uint32_t list_stride_for(uint32_t C, uint32_t k) { return std::max(2u, next_pow2(C * k)); }

static size_t scan_fixed_smem(uint32_t C, uint32_t QB, uint32_t row_stride, uint32_t k, bool emit) {
    return (size_t)QB * row_stride + 64 + (size_t)C * QB * sizeof(ListHdr) +
           (emit ? 0 : (size_t)QB * list_stride_for(C, k) * 8) + 2 * 32 * 8 + 64 * 4 + 128;  // ... barriers (S <= 64), tile slots, slack
}

static int plan_scan(int lpr, bool strict, uint32_t row_stride, uint32_t k, uint32_t nq, bool emit, ScanCfg& c) {
    const uint32_t RPW = 32 / lpr;
    c.C = std::min(16u, std::max(1u, env_u32("VECGPU_SCAN_WARPS", 16)));
    c.QB = emit ? 1 : (nq >= 8 ? 8 : nq >= 4 ? 4 : nq >= 2 ? 2 : 1);
    const uint32_t qb_cap = env_u32("VECGPU_SCAN_QB", 8);
    while (c.QB > 1 && c.QB > qb_cap) c.QB >>= 1;
    // keep lists + queries under ~1/3 of shared memory
    while (c.QB > 1 && scan_fixed_smem(c.C, c.QB, row_stride, k, emit) > SCAN_SMEM_MAX / 3) c.QB >>= 1;
    while (c.C > 1 && scan_fixed_smem(c.C, c.QB, row_stride, k, emit) > SCAN_SMEM_MAX / 2) c.C >>= 1;
    const size_t fixed = scan_fixed_smem(c.C, c.QB, row_stride, k, emit);
    if (fixed + 2 * 16 * RPW > SCAN_SMEM_MAX) return fail(VECGPU_ERR_INVALID_PARAM, "row too wide for the scan kernel");
    const size_t avail = SCAN_SMEM_MAX - fixed;
    const uint32_t stage_target = std::max(1u, env_u32("VECGPU_SCAN_STAGE_KB", 8)) * 1024;
    const bool force_rows = env_u32("VECGPU_SCAN_PERROW", 0) != 0;

    if (!strict && !force_rows && (size_t)RPW * row_stride * 3 <= avail) {
        // contig mode: a stage is RS whole rows, one bulk copy
        c.contig = 1;
        c.CB = row_stride;
        c.n_chunks = 1;
        c.srs = row_stride;
        uint32_t m = std::max(1u, stage_target / (RPW * row_stride));
        while (m > 1 && (size_t)m * RPW * row_stride * 4 > avail) --m;
        c.R = RPW * m;
    } else {
        // per-row mode: padded stage, one bulk copy per row chunk
        c.contig = 0;
        uint32_t cb_limit = (uint32_t)(avail / 3 / RPW);
        cb_limit = cb_limit > 128 ? ((cb_limit - 64) / 64) * 64 : 64;
        uint32_t cb_target = std::min(std::max(64u, (env_u32("VECGPU_SCAN_CB", strict ? 2048 : 4096) / 64) * 64), cb_limit);
        if (row_stride <= cb_target) {
            c.n_chunks = 1;
            c.CB = row_stride;
        } else {
            c.n_chunks = (row_stride + cb_target - 1) / cb_target;
            c.CB = (((row_stride + c.n_chunks - 1) / c.n_chunks) + 63) / 64 * 64;
            c.n_chunks = (row_stride + c.CB - 1) / c.CB;
        }
        uint32_t pad;
        if (lpr == 4) pad = (64 + 128 - (c.CB % 128)) % 128;   // row pitch == 64 (mod 128): 4-lane groups never collide
        else pad = ((c.CB / 16) % 2 == 0) ? 16 : 0;            // odd unit pitch: thread-per-row never collides
        c.srs = c.CB + pad;
        uint32_t m = 1;
        if (c.n_chunks == 1) {
            m = std::max(1u, stage_target / (RPW * c.srs));
            while (m > 1 && (size_t)m * RPW * c.srs * 4 > avail) --m;
        }
        c.R = RPW * m;
    }
    const size_t stage = (size_t)c.R * c.srs;
    uint32_t s_max = (uint32_t)std::min<size_t>(64, avail / stage);
    const uint32_t s_cap = env_u32("VECGPU_SCAN_STAGES", 64);
    if (s_max > s_cap && s_cap >= 1) s_max = s_cap;
    if (s_max < 2) return fail(VECGPU_ERR_INVALID_PARAM, "scan plan does not fit shared memory (row_stride=%u k=%u)", row_stride, k);
    // private ring of D >= 2 stages per consumer warp (so a warp's next stage loads while it computes)
    while (c.C > 1 && s_max / c.C < 2) --c.C;
    c.S = std::max(1u, std::min(s_max / c.C, env_u32("VECGPU_SCAN_RING", 4)));   // D
    c.smem = (size_t)c.S * c.C * stage + fixed;
    return 0;
}

template <class T, int QB, bool EMIT>
static int launch_scan_inst(const ScanParams& p, const ScanCfg& c, dim3 grid, cudaStream_t st) {
    static int configured_for_device = -1;  // per instantiation; attribute is per device
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (configured_for_device != dev) {
        CU(cudaFuncSetAttribute(scan_kernel<T, QB, EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM_MAX));
        configured_for_device = dev;
    }
    scan_kernel<T, QB, EMIT><<<grid, 32 * c.C, c.smem, st>>>(p);
    LAUNCHED();
    return 0;
}

template <template <int> class TT>
static int launch_scan_qb(const ScanParams& p, const ScanCfg& c, bool emit, dim3 grid, cudaStream_t st) {
    if (emit) return launch_scan_inst<TT<1>, 1, true>(p, c, grid, st);
    switch (c.QB) {
        case 1: return launch_scan_inst<TT<1>, 1, false>(p, c, grid, st);
        case 2: return launch_scan_inst<TT<2>, 2, false>(p, c, grid, st);
        case 4: return launch_scan_inst<TT<4>, 4, false>(p, c, grid, st);
        default: return launch_scan_inst<TT<8>, 8, false>(p, c, grid, st);
    }
}

template <int QB> using I8L2T = I8Dot<QB, false>;
template <int QB> using I8CosT = I8Dot<QB, true>;

static bool metric_strict(int elem, int metric) { return elem == VECGPU_F32 && metric == VECGPU_L1; }
static int metric_lpr(int elem, int metric) { return (metric_strict(elem, metric) || elem == VECGPU_BIT) ? 1 : 4; }
static uint32_t metric_qc_kind(int elem) { return elem == VECGPU_I8 ? 1u : 0u; }

static int launch_scan(int elem, int metric, const ScanParams& p, const ScanCfg& c, bool emit, dim3 grid, cudaStream_t st) {
    if (elem == VECGPU_F32) {
        if (metric == VECGPU_L2) return launch_scan_qb<F32L2>(p, c, emit, grid, st);
        if (metric == VECGPU_L1) return launch_scan_qb<F32L1>(p, c, emit, grid, st);
        return launch_scan_qb<F32Cos>(p, c, emit, grid, st);
    }
    if (elem == VECGPU_I8) {
        if (metric == VECGPU_L2) return launch_scan_qb<I8L2T>(p, c, emit, grid, st);
        if (metric == VECGPU_L1) return launch_scan_qb<I8L1>(p, c, emit, grid, st);
        return launch_scan_qb<I8CosT>(p, c, emit, grid, st);
    }
    return launch_scan_qb<BitHamming>(p, c, emit, grid, st);
}

// final selection over the per-CTA partial lists: [nq][n_cand] keys -> k smallest per query, decoded
static int launch_merge(vecgpu_slab* s, const MergeParams& mp, uint32_t nq, cudaStream_t st) {
    if (mp.k <= 32 && mp.n_cand <= 2048 && env_u32("VECGPU_MERGE_SMALL", 1)) {
        merge_small_kernel<<<nq, 256, 0, st>>>(mp);  // two register sorts per query
        LAUNCHED();
        return 0;
    }
    if (mp.n_cand <= 16384) {
        // fits one CTA's shared memory: one bitonic sort per query
        const uint32_t np2 = std::max(2u, next_pow2((uint32_t)mp.n_cand));
        static int cfg_dev = -1;
        int dev = 0;
        CU(cudaGetDevice(&dev));
        if (cfg_dev != dev) {
            CU(cudaFuncSetAttribute(merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            cfg_dev = dev;
        }
        const uint32_t threads = std::min(1024u, std::max(32u, np2 / 2));
        merge_sort_kernel<<<nq, threads, (size_t)np2 * 8, st>>>(mp, np2);
        LAUNCHED();
        return 0;
    }
    // large candidate sets (k in the hundreds): radix-sort each query's candidates, decode the first k
    int rc = ws_reserve(s, WS_TMP, (size_t)mp.n_cand * 8);
    if (rc) return rc;
    size_t sort_bytes = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, mp.keys, (uint64_t*)s->d_ws[WS_TMP], (int)mp.n_cand, 0, 64, st));
    rc = ws_reserve(s, WS_TMP3, sort_bytes);
    if (rc) return rc;
    for (uint32_t q = 0; q < nq; ++q) {
        CU(cub::DeviceRadixSort::SortKeys(s->d_ws[WS_TMP3], sort_bytes, mp.keys + (size_t)q * mp.n_cand,
                                          (uint64_t*)s->d_ws[WS_TMP], (int)mp.n_cand, 0, 64, st));
        g_launches.fetch_add(4, std::memory_order_relaxed);
        decode_segments_kernel<<<1, 256, 0, st>>>((const uint64_t*)s->d_ws[WS_TMP], mp.n_cand, mp.k, mp.rowids, mp.first_rowid,
                                                  mp.pad_rowid, mp.out_rowids + (size_t)q * mp.k,
                                                  mp.out_dists + (size_t)q * mp.k, mp.out_counts ? mp.out_counts + q : nullptr);
        LAUNCHED();
    }
    return 0;
}

// Which scratch set a call on stream `st` uses (see vecgpu_slab::slot_stream).  A third stream takes over set 0 after waiting
// for its previous owner.
static int slab_pick_slot(vecgpu_slab* s, cudaStream_t st, int* slot) {
    for (int i = 0; i < 2; ++i)
        if (s->slot_used[i] && s->slot_stream[i] == st) {
            *slot = i;
            return 0;
        }
    for (int i = 0; i < 2; ++i)
        if (!s->slot_used[i]) {
            s->slot_used[i] = true;
            s->slot_stream[i] = st;
            *slot = i;
            return 0;
        }
    if (!s->slot_ev) CU(cudaEventCreateWithFlags(&s->slot_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(s->slot_ev, s->slot_stream[0]));
    CU(cudaStreamWaitEvent(st, s->slot_ev, 0));
    s->slot_stream[0] = st;
    *slot = 0;
    return 0;
}
// Paths that only have ONE set of scratch buffers run exclusively: a call from the second stream first waits for the first
// stream's work (and the first stream for this call afterwards: slab_exclusive_end).
static int slab_exclusive_begin(vecgpu_slab* s, cudaStream_t st, int slot) {
    if (slot == 0 && !s->slot_used[1]) return 0;
    const int other = 1 - slot;
    if (!s->slot_used[other] || s->slot_stream[other] == st) return 0;
    if (!s->slot_ev) CU(cudaEventCreateWithFlags(&s->slot_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(s->slot_ev, s->slot_stream[other]));
    CU(cudaStreamWaitEvent(st, s->slot_ev, 0));
    return 0;
}
static int slab_exclusive_end(vecgpu_slab* s, cudaStream_t st, int slot) {
    const int other = 1 - slot;
    if (!s->slot_used[other] || s->slot_stream[other] == st) return 0;
    CU(cudaEventRecord(s->slot_ev, st));
    CU(cudaStreamWaitEvent(s->slot_stream[other], s->slot_ev, 0));
    return 0;
}

// f32 L1 single/batched scan through swizzled TMA boxes (scan_l1_tma_kernel); returns 1 if not applicable
static int make_tile_map(CUtensorMap* m, CUtensorMapDataType dtype, uint32_t inner_elems_total, uint32_t box_inner, const void* base,
                         uint64_t rows, uint32_t stride_bytes, uint32_t box_rows, CUtensorMapSwizzle swz);
template <int QB>
static int launch_l1_inst(const CUtensorMap& map, const L1Params& p, dim3 grid, size_t smem, cudaStream_t st) {
    static int cfg_dev = -1;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (cfg_dev != dev) {
        CU(cudaFuncSetAttribute(scan_l1_tma_kernel<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        cfg_dev = dev;
    }
    scan_l1_tma_kernel<QB><<<grid, 32 * p.n_warps, smem, st>>>(map, p);
    LAUNCHED();
    return 0;
}
static int knn_l1_tma(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq, uint32_t k, uint64_t** out_keys, uint32_t* out_gx, cudaStream_t st) {
    if (env_u32("VECGPU_L1_TMA", 1) == 0 || s->rows >= 0x7FFFFFFFull || s->row_stride < 128 || k > 256) return 1;
    uint32_t QB = nq >= 8 ? 8 : nq >= 4 ? 4 : nq >= 2 ? 2 : 1;
    const uint32_t q_stride = (s->row_stride + 127u) & ~127u;
    uint32_t C = 12, D = 4;
    auto fixed = [&](uint32_t c, uint32_t qb) {
        return (size_t)qb * q_stride + (size_t)c * qb * sizeof(ListHdr) + (size_t)qb * list_stride_for(c, k) * 8 + (size_t)c * D * 8 + 1024 + 256;
    };
    while (QB > 1 && fixed(C, QB) > SMEM_MAX / 3) QB >>= 1;
    while (C > 2 && (size_t)C * D * 4096 + fixed(C, QB) > SMEM_MAX) --C;
    if ((size_t)C * D * 4096 + fixed(C, QB) > SMEM_MAX) return 1;
    const size_t smem = (size_t)C * D * 4096 + fixed(C, QB);
    CUtensorMap map;
    int rc = make_tile_map(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, s->row_stride, 128, s->d_vec, s->rows, s->row_stride, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const uint64_t n_tiles = (s->rows + 31) / 32;
    const uint32_t gx = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, n_tiles / 4), (uint64_t)s->num_sms);
    const uint32_t gy = (nq + QB - 1) / QB;
    if ((rc = ws_reserve(s, WS_PART, (size_t)nq * gx * k * 8))) return rc;
    L1Params p{};
    p.skip = s->n_skip ? s->d_skip : nullptr;
    p.queries = d_q;
    p.out_keys = (uint64_t*)s->d_ws[WS_PART];
    p.n_rows = s->rows;
    p.nq_total = nq;
    p.k = k;
    p.row_stride = s->row_stride;
    p.q_stride = q_stride;
    p.n_chunks = q_stride / 128;
    p.ring = D;
    p.n_warps = C;
    p.list_stride = list_stride_for(C, k);
    switch (QB) {
        case 1: rc = launch_l1_inst<1>(map, p, dim3(gx, gy), smem, st); break;
        case 2: rc = launch_l1_inst<2>(map, p, dim3(gx, gy), smem, st); break;
        case 4: rc = launch_l1_inst<4>(map, p, dim3(gx, gy), smem, st); break;
        default: rc = launch_l1_inst<8>(map, p, dim3(gx, gy), smem, st); break;
    }
    if (rc) return rc;
    *out_keys = p.out_keys;
    *out_gx = gx;
    return 0;
}

// queries already on the device, padded to row_stride.  Results to device arrays.
static int knn_exact(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq, uint32_t k, int metric, int64_t* d_out_rowids,
                    float* d_out_dists, uint32_t* d_out_counts, int64_t pad_rowid, cudaStream_t st, int slot = 0) {
    int rc;
    if (k == 0 || nq == 0) return 0;
    const uint8_t* d_skip = s->n_skip ? s->d_skip : nullptr;
    const int64_t* d_rowids = s->dense ? nullptr : s->d_rowids;
    if (s->rows == 0) {
        // nothing to scan: all slots are padding
        MergeParams mp{};
        rc = ws_reserve(s, WS_PART, 8);
        if (rc) return rc;
        CU(cudaMemsetAsync(s->d_ws[WS_PART], 0xFF, 8, st));
        mp.keys = (const uint64_t*)s->d_ws[WS_PART];
        mp.n_cand = 0;
        mp.k = std::min(k, K_FUSED_MAX);
        mp.kp2 = next_pow2(mp.k);
        mp.rowids = nullptr;
        mp.first_rowid = 0;
        mp.pad_rowid = pad_rowid;
        // k may exceed the fused limit: fill by chunks of query-major rows is overkill; do it with memset-like kernel
        if (k <= K_FUSED_MAX) {
            mp.out_rowids = d_out_rowids;
            mp.out_dists = d_out_dists;
            mp.out_counts = d_out_counts;
            return launch_merge(s, mp, nq, st);
        }
        for (uint32_t q = 0; q < nq; ++q) {
            decode_sorted_kernel<<<64, 256, 0, st>>>((const uint64_t*)s->d_ws[WS_PART], 0, k, nullptr, 0, pad_rowid,
                                                     d_out_rowids + (size_t)q * k, d_out_dists + (size_t)q * k, nullptr);
            LAUNCHED();
        }
        if (d_out_counts) CU(cudaMemsetAsync(d_out_counts, 0, nq * sizeof(uint32_t), st));
        return 0;
    }
    const int lpr = metric_lpr(s->elem, metric);
    ScanParams p{};
    p.vectors = s->d_vec;
    p.skip = d_skip;
    p.queries = d_q;
    p.n_rows = s->rows;
    p.nq_total = nq;
    p.row_stride = s->row_stride;
    p.qc_kind = metric_qc_kind(s->elem);

    if (metric_strict(s->elem, metric) && k <= K_FUSED_MAX) {
        // strict-order f32 L1: swizzled TMA boxes (conflict-free thread-per-row) when applicable
        uint64_t* keys = nullptr;
        uint32_t gx = 0;
        rc = knn_l1_tma(s, d_q, nq, k, &keys, &gx, st);
        if (rc == 0) {
            MergeParams mp{};
            mp.keys = keys;
            mp.n_cand = (uint64_t)gx * k;
            mp.k = k;
            mp.kp2 = next_pow2(k);
            mp.rowids = d_rowids;
            mp.first_rowid = s->first_rowid;
            mp.out_rowids = d_out_rowids;
            mp.out_dists = d_out_dists;
            mp.out_counts = d_out_counts;
            mp.pad_rowid = pad_rowid;
            return launch_merge(s, mp, nq, st);
        }
        if (rc != 1) return rc;
    }
    if (s->elem == VECGPU_BIT && nq >= 16 && k <= 32 && s->row_stride <= 128 && s->rows >= 4096 && env_u32("VECGPU_HAM_BATCH", 1)) {
        // many Hamming queries over short rows: lane = query (ham_batch_kernel), 32 queries per pass over the data
        HamBatchParams hp{};
        hp.vectors = s->d_vec;
        hp.skip = d_skip;
        hp.queries = d_q;
        hp.n_rows = s->rows;
        hp.nq = nq;
        hp.k = k;
        hp.row_stride = s->row_stride;
        hp.n_warps = std::min(16u, 256u / k);       // the CTA merge sorts C * k <= 256 keys per query in registers
        hp.rows_per_tile = std::max(1u, 2048u / s->row_stride);
        hp.n_stages = 3;
        const uint64_t n_tiles = (s->rows + hp.rows_per_tile - 1) / hp.rows_per_tile;
        const uint32_t gx = (uint32_t)std::min<uint64_t>((n_tiles + hp.n_warps - 1) / hp.n_warps, (uint64_t)s->num_sms);
        const uint32_t gy = (nq + 31) / 32;
        rc = ws_reserve(s, WS_PART, (size_t)nq * gx * k * 8);
        if (rc) return rc;
        hp.out_keys = (uint64_t*)s->d_ws[WS_PART];
        const size_t smem = (size_t)hp.n_warps * hp.n_stages * hp.rows_per_tile * s->row_stride + (size_t)hp.n_warps * k * 32 * 8 +
                            (size_t)hp.n_warps * hp.n_stages * 8;
        const dim3 grid(gx, gy), block(hp.n_warps * 32);
        switch (s->row_stride / 4) {
#define VECGPU_HAM_CASE(W)                                                                                            \
    case W:                                                                                                           \
        CU(cudaFuncSetAttribute(ham_batch_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));    \
        ham_batch_kernel<W><<<grid, block, smem, st>>>(hp);                                                          \
        break;
            VECGPU_HAM_CASE(4) VECGPU_HAM_CASE(8) VECGPU_HAM_CASE(12) VECGPU_HAM_CASE(16) VECGPU_HAM_CASE(20) VECGPU_HAM_CASE(24)
            VECGPU_HAM_CASE(28) VECGPU_HAM_CASE(32)
#undef VECGPU_HAM_CASE
            default: return fail(VECGPU_ERR_INVALID_PARAM, "unexpected row stride for the Hamming batch kernel");
        }
        LAUNCHED();
        MergeParams mp{};
        mp.keys = hp.out_keys;
        mp.n_cand = (uint64_t)gx * k;
        mp.k = k;
        mp.kp2 = next_pow2(k);
        mp.rowids = d_rowids;
        mp.first_rowid = s->first_rowid;
        mp.out_rowids = d_out_rowids;
        mp.out_dists = d_out_dists;
        mp.out_counts = d_out_counts;
        mp.pad_rowid = pad_rowid;
        return launch_merge(s, mp, nq, st);
    }
    if (k <= K_FUSED_MAX) {
        ScanCfg c;
        rc = plan_scan(lpr, metric_strict(s->elem, metric), s->row_stride, k, nq, false, c);
        if (rc) return rc;
        const uint64_t n_tiles = (s->rows + c.R - 1) / c.R;
        uint32_t gx = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)s->num_sms);
        if (const uint32_t cap_gx = env_u32("VECGPU_SCAN_GX", 0)) {
            gx = std::max(1u, std::min(gx, cap_gx));
        } else if (nq <= c.QB && k >= 28 && (uint64_t)gx * k <= 16384) {
            // A small table with a long result list, one query pass: the fused final merge sorts next_pow2(gx * k) keys in ONE
            // CTA, once per query of the pass, which costs more than the scan itself (10 k x f32[384], k = 100: 236 us with 148
            // CTAs, 81 us with 16; four queries: 1101 vs 299 us — tools/small_table_gx.py).  Fewer CTAs stream more rows each
            // (~42 GB/s per CTA, ~0.43 more per extra query of the pass) but leave fewer partial lists; the fit below (us)
            // picks the count.  Tables of any size where streaming dominates keep one CTA per SM.
            const double bytes = (double)s->rows * s->row_stride * (1.0 + 0.43 * (nq - 1));
            auto est = [&](uint32_t g) { return bytes / ((double)g * 42e3) + 0.0115 * nq * (double)std::max(2u, next_pow2(g * k)); };
            uint32_t best = gx;
            double t_best = est(gx);
            for (uint32_t keys = nq == 1 ? 1024 : 2048; keys <= 8192; keys <<= 1) {
                const uint32_t g = std::min(gx, std::max(1u, keys / k));
                if (est(g) < t_best) {
                    t_best = est(g);
                    best = g;
                }
            }
            gx = best;
        }
        const uint32_t gy = (nq + c.QB - 1) / c.QB;
        const int ws_part = slot ? WS_PART1 : WS_PART, ws_ticket = slot ? WS_TICKET1 : WS_TICKET;  // per-stream scratch set
        rc = ws_reserve(s, ws_part, (size_t)nq * gx * k * 8);
        if (rc) return rc;
        p.out_keys = (uint64_t*)s->d_ws[ws_part];
        p.k = k;
        p.chunk_bytes = c.CB;
        p.n_chunks = c.n_chunks;
        p.smem_row_stride = c.srs;
        p.rows_per_stage = c.R;
        p.n_stages = c.S;
        p.contig = c.contig;
        p.n_consumers = c.C;
        p.list_stride = list_stride_for(c.C, k);
        MergeParams mp{};
        mp.keys = p.out_keys;
        mp.n_cand = (uint64_t)gx * k;
        mp.k = k;
        mp.kp2 = next_pow2(k);
        mp.rowids = d_rowids;
        mp.first_rowid = s->first_rowid;
        mp.out_rowids = d_out_rowids;
        mp.out_dists = d_out_dists;
        mp.out_counts = d_out_counts;
        mp.pad_rowid = pad_rowid;
        // One query pass (nq <= QB): the last CTA to finish merges the gx partial lists itself (and carries a sharded query's
        // push), so a query is ONE launch instead of scan + merge (+ push).
        const uint32_t np2 = std::max(2u, next_pow2((uint32_t)mp.n_cand));
        const bool fuse = gy == 1 && env_u32("VECGPU_FUSE_MERGE", 1) && mp.n_cand <= 16384 &&
                          (size_t)np2 * 8 <= (size_t)c.S * c.C * c.R * c.srs;
        // Dynamic tile hand-out (experimental, OFF by default).  The static round robin leaves the CTAs finishing 4-8 % apart
        // (tools/scan_timeline.py: first CTA done after 534 us, last after 557 us on a 3.84 GB shard), but every dynamic form
        // tried was slower overall: with 4 warps x 2 stages of 24 KB per SM each warp's refill is latency-critical, the
        // bookkeeping alone (ticket -> tile through shared memory) costs 9 % and handing tiles out in arrival order another
        // 6 % (profiles/r2_scan_timeline5.txt).  VECGPU_SCAN_DYNAMIC: 0 = static (default), 1 = 90 % static then dynamic,
        // 2 = all dynamic (16-tile blocks per CTA), 3 = static order through the dynamic code path.
        const uint32_t dyn_mode = env_u32("VECGPU_SCAN_DYNAMIC", 0);
        const uint64_t per_round = (uint64_t)gx * c.C;
        const bool dynamic = c.n_chunks == 1 && dyn_mode != 0 && n_tiles >= 16 * per_round && n_tiles < 0xFFFFFF00ull;
        if (dynamic) {
            const uint64_t rounds = n_tiles / per_round;
            p.tail.static_rounds = dyn_mode == 2 ? 0u : dyn_mode == 3 ? (uint32_t)rounds
                                   : (uint32_t)(rounds * env_u32("VECGPU_SCAN_STATIC_PCT", 90) / 100);
        }
        if (fuse || dynamic) {
            // tickets + tile counters, one pair per query pass; zeroed when (re)allocated, re-armed by the last CTA of every pass
            const size_t need = std::max<size_t>(256, (size_t)gy * 8);
            if (need > s->ws_cap[ws_ticket]) {
                if ((rc = ws_reserve(s, ws_ticket, need))) return rc;
                CU(cudaMemsetAsync(s->d_ws[ws_ticket], 0, s->ws_cap[ws_ticket], st));
            }
            p.tail.counter = (uint32_t*)s->d_ws[ws_ticket];
            p.tail.dyn = dynamic ? (uint32_t*)s->d_ws[ws_ticket] + gy : nullptr;
            p.tail.do_merge = fuse ? 1u : 0u;
        }
        if (fuse) {
            p.tail.np2 = np2;
            p.tail.mp = mp;
            if (s->fuse_push) {
                p.tail.push = *s->fuse_push;
                s->fuse_push_done = true;
            }
        }
        if (!s->q_uploaded && s->h_q_staged && nq == 1 && s->row_stride <= SCAN_INLINE_Q_MAX) {
            p.inline_q_bytes = s->row_stride;
            memcpy(p.inline_q, s->h_q_staged, s->row_stride);
        } else if ((rc = query_upload_if_needed(s, st))) {
            return rc;
        }
        p.dbg = g_scan_dbg;  // developer timeline (vecgpu_debug_scan_timeline), normally NULL
        rc = launch_scan(s->elem, metric, p, c, false, dim3(gx, gy), st);
        if (rc) return rc;
        if (fuse) return 0;
        return launch_merge(s, mp, nq, st);
    }

    // ---- large k: emit one key per row, radix sort, decode the first k ----
    ScanCfg c;
    rc = plan_scan(lpr, metric_strict(s->elem, metric), s->row_stride, 1, 1, true, c);
    if (rc) return rc;
    const uint64_t n_tiles = (s->rows + c.R - 1) / c.R;
    const uint32_t gx = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)s->num_sms);
    rc = ws_reserve(s, WS_TMP, (size_t)s->rows * 8);
    if (rc) return rc;
    rc = ws_reserve(s, WS_TMP2, (size_t)s->rows * 8);
    if (rc) return rc;
    size_t sort_bytes = 0;
    cub::DoubleBuffer<uint64_t> db((uint64_t*)s->d_ws[WS_TMP], (uint64_t*)s->d_ws[WS_TMP2]);
    CU(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, db, (int)s->rows, 0, 64, st));
    rc = ws_reserve(s, WS_TMP3, sort_bytes);
    if (rc) return rc;
    if (d_out_counts) CU(cudaMemsetAsync(d_out_counts, 0, nq * sizeof(uint32_t), st));
    p.k = 1;
    p.chunk_bytes = c.CB;
    p.n_chunks = c.n_chunks;
    p.smem_row_stride = c.srs;
    p.rows_per_stage = c.R;
    p.n_stages = c.S;
    p.contig = c.contig;
    p.n_consumers = c.C;
    p.list_stride = 2;
    p.nq_total = 1;
    for (uint32_t q = 0; q < nq; ++q) {
        cub::DoubleBuffer<uint64_t> dbq((uint64_t*)s->d_ws[WS_TMP], (uint64_t*)s->d_ws[WS_TMP2]);
        p.queries = d_q + (size_t)q * s->row_stride;
        p.out_keys = dbq.Current();
        rc = launch_scan(s->elem, metric, p, c, true, dim3(gx, 1), st);
        if (rc) return rc;
        CU(cub::DeviceRadixSort::SortKeys(s->d_ws[WS_TMP3], sort_bytes, dbq, (int)s->rows, 0, 64, st));
        g_launches.fetch_add(8, std::memory_order_relaxed);
        decode_sorted_kernel<<<256, 256, 0, st>>>(dbq.Current(), s->rows, k, d_rowids, s->first_rowid, pad_rowid,
                                                  d_out_rowids + (size_t)q * k, d_out_dists + (size_t)q * k,
                                                  d_out_counts ? d_out_counts + q : nullptr);
        LAUNCHED();
    }
    return 0;
}

// ---------------------------------------------------------------------------
// K2: tensor-core batched path (tc_batch.cuh)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tile_map(CUtensorMap* m, CUtensorMapDataType dtype, uint32_t inner_elems_total, uint32_t box_inner, const void* base,
                         uint64_t rows, uint32_t stride_bytes, uint32_t box_rows, CUtensorMapSwizzle swz) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn) return fail(VECGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {inner_elems_total, rows};
    cuuint64_t gstr[1] = {stride_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};  // inner box = one swizzle span (128 or 64 bytes)
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, dtype, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VECGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}
static int make_f32_map(CUtensorMap* m, const void* base, uint32_t dims, uint64_t rows, uint32_t stride_bytes, uint32_t box_rows) {
    return make_tile_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dims, TC_KC, base, rows, stride_bytes, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int slab_ensure_norms(vecgpu_slab* s, cudaStream_t st) {
    if (s->norms_valid) return 0;
    if (s->rows > s->cap_norms) {
        if (s->d_norms) CU(cudaFree(s->d_norms));
        s->d_norms = nullptr;
        s->cap_norms = std::max<uint64_t>(s->rows, s->cap);
        CU(cudaMalloc((void**)&s->d_norms, s->cap_norms * sizeof(float)));
    }
    if (s->elem == VECGPU_I8) {  // exact int32 |x|^2 (same 4-byte slots)
        row_norms_i8_kernel<<<(uint32_t)s->num_sms * 8, 256, 0, st>>>(s->d_vec, s->row_stride, s->row_stride / 16, s->rows, (int*)s->d_norms);
        LAUNCHED();
        s->n_unsafe = 0;
        s->norms_valid = true;
        return 0;
    }
    if (!s->d_x2max) CU(cudaMalloc((void**)&s->d_x2max, 4));
    if (!s->d_unsafe) CU(cudaMalloc((void**)&s->d_unsafe, (1 + TC_MAX_UNSAFE) * 4));
    CU(cudaMemsetAsync(s->d_x2max, 0, 4, st));
    CU(cudaMemsetAsync(s->d_unsafe, 0, (1 + TC_MAX_UNSAFE) * 4, st));
    row_norms_kernel<<<(uint32_t)s->num_sms * 8, 256, 0, st>>>(s->d_vec, s->row_stride, s->row_stride / 16, s->rows, s->d_norms,
                                                               s->d_x2max, s->d_unsafe, 0u);
    LAUNCHED();
    CU(cudaMemcpyAsync(&s->n_unsafe, s->d_unsafe, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->norms_valid = true;
    return 0;
}

// Rows [pos, pos+n) were just (re)written: refresh their cached norms in place instead of invalidating the cache (an upsert
// used to make the next batched query recompute all of them — a full pass over the slab).  The maximum only ever grows and a
// row that was unusable stays on the always-re-ranked list: both remain valid (conservative) bounds.
static int slab_update_norms(vecgpu_slab* s, uint64_t pos, uint64_t n) {
    if (!s->norms_valid || n == 0) return 0;
    if (s->elem == VECGPU_BIT) {
        s->norms_valid = false;
        return 0;
    }
    if (pos + n > s->cap_norms) {  // rows appended past the cache's capacity: it grows with the slab, the cached values move along
        const uint64_t cap = std::max<uint64_t>(pos + n, s->cap);
        float* nn = nullptr;
        CU(cudaMalloc((void**)&nn, cap * sizeof(float)));
        CU(cudaMemcpyAsync(nn, s->d_norms, s->cap_norms * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->d_norms);
        s->d_norms = nn;
        s->cap_norms = cap;
    }
    const uint8_t* base = s->d_vec + pos * s->row_stride;
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n + 63) / 64, (uint64_t)s->num_sms * 8));
    if (s->elem == VECGPU_I8) {
        row_norms_i8_kernel<<<blocks, 256, 0, s->stream>>>(base, s->row_stride, s->row_stride / 16, n, (int*)s->d_norms + pos);
        LAUNCHED();
        CU(cudaStreamSynchronize(s->stream));
        return 0;
    }
    row_norms_kernel<<<blocks, 256, 0, s->stream>>>(base, s->row_stride, s->row_stride / 16, n, s->d_norms + pos, s->d_x2max, s->d_unsafe,
                                                    (uint32_t)pos);
    LAUNCHED();
    CU(cudaMemcpyAsync(&s->n_unsafe, s->d_unsafe, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

static bool tc_eligible(const vecgpu_slab* s, uint32_t nq, uint32_t k, int metric) {
    if (env_u32("VECGPU_TC", 1) == 0) return false;
    // the tensor-core launch has ~0.3 ms of fixed cost (one compaction per query lane at the end of every CTA): batches of
    // fewer than 64 queries with less than about 2 M (query, row) pairs are faster through the exact multi-query scan on the
    // CUDA cores (measured, tools/tc_threshold.py: 32 x 10 k rows 155 vs 365 us, 32 x 100 k rows 830 vs 535 us, 100 x 10 k 440 vs 367 us)
    return s->elem == VECGPU_F32 && (metric == VECGPU_L2 || metric == VECGPU_COSINE) && nq >= env_u32("VECGPU_TC_MIN_NQ", 16) &&
           s->rows >= 8192 && s->rows < 0x7FFFFFFFull && k <= 96 && s->dims >= 16 &&
           (nq >= 64 || (uint64_t)nq * s->rows >= env_u32("VECGPU_TC_MIN_WORK", 2000000));
}

static int knn_exact(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq, uint32_t k, int metric, int64_t* d_out_rowids,
                     float* d_out_dists, uint32_t* d_out_counts, int64_t pad_rowid, cudaStream_t st, int slot);
static int launch_pairs(int elem, int metric, const PairParams& p, int num_sms, cudaStream_t st);

static int knn_tc(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq_all, uint32_t k, int metric, int64_t* d_out_rowids,
                  float* d_out_dists, uint32_t* d_out_counts, int64_t pad_rowid, cudaStream_t st) {
    int rc = slab_ensure_norms(s, st);
    if (rc) return rc;
    if (s->n_unsafe > TC_MAX_UNSAFE)  // too many rows with unusable norms: the whole batch goes through the exact scan
        return knn_exact(s, d_q, nq_all, k, metric, d_out_rowids, d_out_dists, d_out_counts, pad_rowid, st, 0);
    static int cfg_dev = -1;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const uint32_t kp = ((k + 32 + 7) / 8) * 8;
    const size_t smem = TC_STAGES * TC_STAGE_BYTES + 2 * 4 * TC_N * 4 + 32 * 8 + 4 * 2048 + 1024;  // stages, per-warp coefficients, barriers, rare-path staging, alignment slack
    if (cfg_dev != dev) {
        CU(cudaFuncSetAttribute(tc_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        CU(cudaFuncSetAttribute(tc_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        CU(cudaFuncSetAttribute(tc_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CU(cudaFuncSetAttribute(merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        cfg_dev = dev;
    }
    const uint64_t n_xt = (s->rows + TC_N - 1) / TC_N;
    const uint32_t max_qt = std::max(1u, std::min(16u, (uint32_t)s->num_sms));  // <= 2048 queries per launch
    CUtensorMap mapX;
    if ((rc = make_f32_map(&mapX, s->d_vec, s->dims, s->rows, s->row_stride, TC_N))) return rc;
    const uint64_t live = s->rows - s->n_skip;

    for (uint32_t qoff = 0; qoff < nq_all; qoff += max_qt * TC_M) {
        const uint32_t nq = std::min(nq_all - qoff, max_qt * TC_M);
        const uint8_t* dq = d_q + (size_t)qoff * s->row_stride;
        const uint32_t QT = (nq + TC_M - 1) / TC_M;
        uint32_t G = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)s->num_sms / QT, n_xt));
        // thread-block clusters: the CTAs of consecutive query tiles share every row tile through TMA multicast (each loads
        // 1/cs of it for all), so the slab crosses L2 -> SM once per cluster instead of once per query tile
        uint32_t cs = 1;
        {
            // measured on 1024 x 10 M x 768: 31.8 / 30.8 / 31.2 / 30.9 ms at cs = 8 / 4 / 2 / 1 (with the soft lockstep at cs = 1):
            // the single-pass kernel is bound by the depth of its 4 x 48 KB operand pipeline, not by L2 traffic, so clusters
            // are off by default (VECGPU_TC_CLUSTER = 2, 4, 8 enables them)
            const uint32_t want = std::min(8u, env_u32("VECGPU_TC_CLUSTER", 1));
            while (cs * 2 <= want && QT % (cs * 2) == 0) cs *= 2;
        }
        // CTA PAIRS (tcgen05 cta_group::2, VECGPU_TC_PAIR=1): two consecutive query tiles of a row group share every row tile —
        // M = 256 per instruction, each SM stages its own 128 queries and half of the rows (8 + 8 KB of shared-memory traffic
        // per K-step instead of 12 + 12).  Implemented, exact and tested, but measured NO faster on 1024 x 10 M x 768
        // (25.7 vs 25.6 ms, profiles/r2_tc_pair_perf.txt): the batch runs under the 1 kW power cap — even a pure MMA loop over
        // resident operands drops from 1117 to ~1050 TFLOP/s within half a second as the SM clock falls from 1965 to
        // 1665 MHz (profiles/r2_mma_peak_long.txt) — so halving the operand traffic does not buy throughput.  Off by default.
        bool pair = env_u32("VECGPU_TC_TERMS", 1) != 3 && QT % 2 == 0 && env_u32("VECGPU_TC_PAIR", 0) != 0 && cs == 1;
        if (pair) cs = 2;
        if (cs > 1) {
            cudaLaunchConfig_t occ{};
            occ.gridDim = dim3(QT * G);
            occ.blockDim = dim3(TC_THREADS);
            occ.dynamicSmemBytes = smem;
            cudaLaunchAttribute oa{};
            oa.id = cudaLaunchAttributeClusterDimension;
            oa.val.clusterDim.x = cs;
            oa.val.clusterDim.y = oa.val.clusterDim.z = 1;
            occ.attrs = &oa;
            occ.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, pair ? tc_scan_kernel<true> : tc_scan_kernel<false>, &occ) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                cs = 1;
                pair = false;
            } else {
                // all clusters must be co-resident (one wave): fewer row groups when the GPCs cannot hold QT*G/cs clusters
                G = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(G, (uint64_t)max_clusters * cs / QT));
            }
        }
        while (G > 1 && next_pow2(G * kp + TC_MAX_UNSAFE) > 8192) --G;
        const uint32_t cap = std::max(2u, next_pow2(G * kp + TC_MAX_UNSAFE));
        const uint32_t grid = QT * G;
        const size_t lists = (size_t)grid * TC_M;
        if ((rc = ws_reserve(s, WS_TC_BUF, lists * TC_BUF_CAP * 8))) return rc;
        if ((rc = ws_reserve(s, WS_TC_CANDV, lists * kp * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_CANDR, lists * kp * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_CNT, lists * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_TAU, lists * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_PAIRQ, (size_t)nq * cap * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_PAIRPOS, (size_t)nq * cap * 8))) return rc;
        if ((rc = ws_reserve(s, WS_TC_DIST, (size_t)nq * cap * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_KEYS, (size_t)nq * cap * 8))) return rc;
        if ((rc = ws_reserve(s, WS_TC_QNORM, (size_t)nq * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_FLAGS, (size_t)nq))) return rc;

        // |q|^2 in the canonical order
        row_norms_kernel<<<std::max(1u, std::min((nq + 63) / 64, 1024u)), 256, 0, st>>>(dq, s->row_stride, s->row_stride / 16, nq,
                                                                                        (float*)s->d_ws[WS_TC_QNORM], nullptr, nullptr, 0u);
        LAUNCHED();
        CUtensorMap mapQ;
        if ((rc = make_f32_map(&mapQ, dq, s->dims, nq, s->row_stride, TC_M))) return rc;
        TcParams tp{};
        tp.n_rows = s->rows;
        tp.nq = nq;
        tp.nk = (s->dims + TC_KC - 1) / TC_KC;
        tp.kp = kp;
        tp.cosine = metric == VECGPU_COSINE ? 1u : 0u;
        tp.terms = env_u32("VECGPU_TC_TERMS", 1) == 3 ? 3u : 1u;  // 1: one TF32 pass + wider certified bound (default); 3: 3xTF32
        tp.debug = env_u32("VECGPU_TCI_DEBUG", 0);
        tp.QT = QT;
        tp.G = G;
        tp.norms = s->d_norms;
        tp.skip = s->n_skip ? s->d_skip : nullptr;
        tp.buf_keys = (uint64_t*)s->d_ws[WS_TC_BUF];
        tp.cand_v = (float*)s->d_ws[WS_TC_CANDV];
        tp.cand_r = (uint32_t*)s->d_ws[WS_TC_CANDR];
        tp.cand_cnt = (uint32_t*)s->d_ws[WS_TC_CNT];
        tp.cand_tau = (float*)s->d_ws[WS_TC_TAU];
        tp.lockstep = nullptr;
        if (QT > 1 && QT <= 32 && env_u32("VECGPU_TC_LOCKSTEP", 1)) {  // the QT CTAs of a row group stream their tiles together
            if ((rc = ws_reserve(s, WS_TC_LOCK, (size_t)G * 32 * 4))) return rc;
            CU(cudaMemsetAsync(s->d_ws[WS_TC_LOCK], 0, (size_t)G * 32 * 4, st));
            tp.lockstep = (uint32_t*)s->d_ws[WS_TC_LOCK];
            tp.lock_slack = env_u32("VECGPU_TC_LOCKSLACK", 0);
        }
        tp.cs = cs;
        tp.pair = pair ? 1u : 0u;
        if (cs > 1) {
            CUtensorMap mapXs;  // row-tile slices of TC_N / cs rows
            if ((rc = make_f32_map(&mapXs, s->d_vec, s->dims, s->rows, s->row_stride, TC_N / cs))) return rc;
            if (cs == QT) tp.lockstep = nullptr;  // the shared stages already keep the whole row group together
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3(grid);
            lc.blockDim = dim3(TC_THREADS);
            lc.dynamicSmemBytes = smem;
            lc.stream = st;
            cudaLaunchAttribute la{};
            la.id = cudaLaunchAttributeClusterDimension;
            la.val.clusterDim.x = cs;
            la.val.clusterDim.y = la.val.clusterDim.z = 1;
            lc.attrs = &la;
            lc.numAttrs = 1;
            if (pair) CU(cudaLaunchKernelEx(&lc, tc_scan_kernel<true>, mapQ, mapXs, tp));
            else CU(cudaLaunchKernelEx(&lc, tc_scan_kernel<false>, mapQ, mapXs, tp));
        } else {
            tc_scan_kernel<false><<<grid, TC_THREADS, smem, st>>>(mapQ, mapX, tp);
        }
        LAUNCHED();

        TcCollectParams cp{};
        cp.t = tp;
        cp.k = (uint32_t)std::min<uint64_t>(k, live);
        cp.cap = cap;
        cp.dims = s->dims;
        cp.qnorm = (const float*)s->d_ws[WS_TC_QNORM];
        cp.x2max_bits = s->d_x2max;
        cp.n_live = live;
        cp.pair_q = (uint32_t*)s->d_ws[WS_TC_PAIRQ];
        cp.pair_pos = (int64_t*)s->d_ws[WS_TC_PAIRPOS];
        cp.fallback = (uint8_t*)s->d_ws[WS_TC_FLAGS];
        cp.unsafe = s->d_unsafe;
        if (cp.k == 0) cp.k = 1;
        tc_collect_kernel<<<nq, 512, (size_t)cap * 8, st>>>(cp);
        LAUNCHED();

        // exact re-rank of the certified candidates in the canonical order
        PairParams pp{};
        pp.a_base = dq;
        pp.b_base = s->d_vec;
        pp.a_stride = pp.b_stride = s->row_stride;
        pp.units = s->row_stride / 16;
        pp.a_index = cp.pair_q;
        pp.b_index = cp.pair_pos;
        pp.n_pairs = (uint64_t)nq * cap;
        pp.out = (float*)s->d_ws[WS_TC_DIST];
        pp.qc_kind = 0;
        if ((rc = launch_pairs(s->elem, metric, pp, s->num_sms, st))) return rc;
        tc_keys_kernel<<<(uint32_t)std::min<uint64_t>((pp.n_pairs + 255) / 256, 4096), 256, 0, st>>>(
            pp.out, cp.pair_pos, pp.n_pairs, (uint64_t*)s->d_ws[WS_TC_KEYS]);
        LAUNCHED();
        MergeParams mp{};
        mp.keys = (const uint64_t*)s->d_ws[WS_TC_KEYS];
        mp.n_cand = cap;
        mp.k = k;
        mp.kp2 = next_pow2(k);
        mp.rowids = s->dense ? nullptr : s->d_rowids;
        mp.first_rowid = s->first_rowid;
        mp.out_rowids = d_out_rowids + (size_t)qoff * k;
        mp.out_dists = d_out_dists + (size_t)qoff * k;
        mp.out_counts = d_out_counts ? d_out_counts + qoff : nullptr;
        mp.pad_rowid = pad_rowid;
        if ((rc = launch_merge(s, mp, nq, st))) return rc;

        // queries whose candidate bound could not be certified: exact scan (rare: massive ties, non-finite data)
        std::vector<uint8_t> flags(nq);
        CU(cudaMemcpyAsync(flags.data(), s->d_ws[WS_TC_FLAGS], nq, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        g_tc_queries.fetch_add(nq, std::memory_order_relaxed);
        for (uint32_t q = 0; q < nq; ++q)
            if (flags[q]) {
                g_tc_fallbacks.fetch_add(1, std::memory_order_relaxed);
                rc = knn_exact(s, dq + (size_t)q * s->row_stride, 1, k, metric, d_out_rowids + (size_t)(qoff + q) * k,
                               d_out_dists + (size_t)(qoff + q) * k, d_out_counts ? d_out_counts + qoff + q : nullptr, pad_rowid, st, 0);
                if (rc) return rc;
            }
    }
    return 0;
}

static bool tci8_eligible(const vecgpu_slab* s, uint32_t nq, uint32_t k, int metric) {
    if (env_u32("VECGPU_TC", 1) == 0) return false;
    return s->elem == VECGPU_I8 && metric == VECGPU_L2 && nq >= env_u32("VECGPU_TC_MIN_NQ", 16) && s->rows >= 8192 &&
           s->rows < 0x7FFFFFFFull && k <= 192 && s->dims >= 16 && s->dims <= 16384;
}

// int8 L2 batches: exact on the tensor cores (tci8_scan_kernel) + the ordinary final merge
static int knn_tci8(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq_all, uint32_t k, int metric, int64_t* d_out_rowids,
                    float* d_out_dists, uint32_t* d_out_counts, int64_t pad_rowid, cudaStream_t st) {
    (void)metric;
    int rc = slab_ensure_norms(s, st);
    if (rc) return rc;
    static int cfg_dev = -1;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const uint32_t cap = std::max(128u, next_pow2(k + 64));  // per-thread append buffer; k <= 192 -> cap <= 256
    const size_t smem = TCI_STAGES * TCI_STAGE_BYTES + TCI_EPI_WARPS * TCI_HALF * 4 + 32 * 8 +
                        TCI_EPI_WARPS * (size_t)2048 + 1024;
    if (cfg_dev != dev) {
        CU(cudaFuncSetAttribute(tci8_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        CU(cudaFuncSetAttribute(tci8_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        cfg_dev = dev;
    }
    const uint64_t n_xt = (s->rows + TC_N - 1) / TC_N;
    const uint32_t max_qt = 16;
    CUtensorMap mapX;
    if ((rc = make_tile_map(&mapX, CU_TENSOR_MAP_DATA_TYPE_UINT8, s->dims, 128, s->d_vec, s->rows, s->row_stride, TC_N, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    for (uint32_t qoff = 0; qoff < nq_all; qoff += max_qt * TC_M) {
        const uint32_t nq = std::min(nq_all - qoff, max_qt * TC_M);
        const uint8_t* dq = d_q + (size_t)qoff * s->row_stride;
        const uint32_t QT = (nq + TC_M - 1) / TC_M;
        uint32_t G = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)s->num_sms / QT, n_xt));
        while (G > 1 && (uint64_t)2 * G * k > 16384) --G;  // keep the final merge in one CTA's shared memory
        // Thresholds from a sample: a first launch ranks a small prefix of the slab, the k-th best key of every query then
        // bounds the main launch from its first tile on (instead of each thread rediscovering it on its 1/2G of the rows)
        uint64_t samp_tiles = std::max<uint64_t>(n_xt / 32, std::min<uint64_t>(n_xt / 8, 1024));  // ~3 % of the slab, at least 256 k rows
        if (const char* e = getenv("VECGPU_TCI_SAMPLE")) samp_tiles = std::min<uint64_t>(n_xt / 2, (uint64_t)atoll(e));
        const bool sample = samp_tiles >= 2ull * G && (uint64_t)4 * G * k <= 16384;
        if (!sample) samp_tiles = 0;
        const uint32_t parts = sample ? 4 * G : 2 * G;
        if ((rc = ws_reserve(s, WS_PART, (size_t)nq * parts * k * 8))) return rc;
        if ((rc = ws_reserve(s, WS_TC_KEYS, (size_t)nq * 2 * G * cap * 8))) return rc;
        if ((rc = ws_reserve(s, WS_TC_QNORM, (size_t)nq * 4))) return rc;
        if ((rc = ws_reserve(s, WS_TC_TAU, (size_t)nq * 8))) return rc;
        row_norms_i8_kernel<<<std::max(1u, std::min((nq + 63) / 64, 1024u)), 256, 0, st>>>(dq, s->row_stride, s->row_stride / 16, nq,
                                                                                           (int*)s->d_ws[WS_TC_QNORM]);
        LAUNCHED();
        CUtensorMap mapQ;
        if ((rc = make_tile_map(&mapQ, CU_TENSOR_MAP_DATA_TYPE_UINT8, s->dims, 128, dq, nq, s->row_stride, TC_M, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        TciParams tp{};
        tp.n_rows = s->rows;
        tp.nq = nq;
        tp.nk = (s->dims + 127) / 128;
        tp.k = k;
        tp.QT = QT;
        tp.G = G;
        tp.norms = (const int*)s->d_norms;
        tp.qnorms = (const int*)s->d_ws[WS_TC_QNORM];
        tp.skip = s->n_skip ? s->d_skip : nullptr;
        tp.out_keys = (uint64_t*)s->d_ws[WS_PART];
        tp.buf_keys = (uint64_t*)s->d_ws[WS_TC_KEYS];
        tp.cap = cap;
        tp.parts_total = parts;
        tp.debug = env_u32("VECGPU_TCI_DEBUG", 0);
        // measured: with 5 us int8 tiles the coupling costs more than the L2 misses it saves (off by default, VECGPU_TCI_LOCKSTEP=1 enables)
        const bool lockstep = QT > 1 && QT <= 32 && env_u32("VECGPU_TCI_LOCKSTEP", 0);
        if (lockstep && (rc = ws_reserve(s, WS_TC_LOCK, (size_t)G * 32 * 4))) return rc;
        tp.lockstep = lockstep ? (uint32_t*)s->d_ws[WS_TC_LOCK] : nullptr;
        tp.lock_slack = env_u32("VECGPU_TC_LOCKSLACK", 1);
        if (sample) {
            if (lockstep) CU(cudaMemsetAsync(s->d_ws[WS_TC_LOCK], 0, (size_t)G * 32 * 4, st));
            tp.tile_begin = 0;
            tp.tile_end = samp_tiles;
            tp.part_base = 0;
            tp.tau_init = nullptr;
            tci8_scan_kernel<<<QT * G, TCI_THREADS, smem, st>>>(mapQ, mapX, tp);
            LAUNCHED();
            const uint32_t n_keys = 2 * G * k, np2 = std::max(2u, next_pow2(n_keys));
            tci8_tau_kernel<<<nq, std::min(1024u, std::max(32u, np2 / 2)), (size_t)np2 * 8, st>>>(tp.out_keys, n_keys, (uint64_t)parts * k, k, np2,
                                                                                                  (uint64_t*)s->d_ws[WS_TC_TAU]);
            LAUNCHED();
        }
        tp.tile_begin = samp_tiles;
        tp.tile_end = n_xt;
        tp.part_base = sample ? 2 * G : 0;
        tp.tau_init = sample ? (const uint64_t*)s->d_ws[WS_TC_TAU] : nullptr;
        if (lockstep) CU(cudaMemsetAsync(s->d_ws[WS_TC_LOCK], 0, (size_t)G * 32 * 4, st));
        tci8_scan_kernel<<<QT * G, TCI_THREADS, smem, st>>>(mapQ, mapX, tp);
        LAUNCHED();
        g_tc_queries.fetch_add(nq, std::memory_order_relaxed);
        MergeParams mp{};
        mp.keys = tp.out_keys;
        mp.n_cand = (uint64_t)parts * k;
        mp.k = k;
        mp.kp2 = next_pow2(k);
        mp.rowids = s->dense ? nullptr : s->d_rowids;
        mp.first_rowid = s->first_rowid;
        mp.out_rowids = d_out_rowids + (size_t)qoff * k;
        mp.out_dists = d_out_dists + (size_t)qoff * k;
        mp.out_counts = d_out_counts ? d_out_counts + qoff : nullptr;
        mp.pad_rowid = pad_rowid;
        if ((rc = launch_merge(s, mp, nq, st))) return rc;
    }
    return 0;
}

static int knn_core(vecgpu_slab* s, const uint8_t* d_q, uint32_t nq, uint32_t k, int metric, int64_t* d_out_rowids,
                    float* d_out_dists, uint32_t* d_out_counts, int64_t pad_rowid, cudaStream_t st) {
    int rc;
    // a staged host query that the streaming scan can carry in its parameters is not uploaded at all
    const bool inline_q = !s->q_uploaded && nq == 1 && s->row_stride <= SCAN_INLINE_Q_MAX && k != 0 && k <= K_FUSED_MAX && s->rows != 0 &&
                          !(metric_strict(s->elem, metric) && s->row_stride >= 128 && k <= 256) /* f32 L1 has its own TMA kernel */ &&
                          env_u32("VECGPU_INLINE_QUERY", 1) != 0;
    if (!inline_q && (rc = query_upload_if_needed(s, st))) return rc;
    int slot = 0;
    if ((rc = slab_pick_slot(s, st, &slot))) return rc;
    const bool tci8 = nq && k && tci8_eligible(s, nq, k, metric), tc = !tci8 && nq && k && tc_eligible(s, nq, k, metric);
    // the plain streaming scan has a scratch set per stream; everything else runs exclusively
    const bool plain = !tci8 && !tc && k <= K_FUSED_MAX && s->rows != 0 &&
                       !(metric_strict(s->elem, metric) && s->row_stride >= 128 && k <= 256) &&
                       !(s->elem == VECGPU_BIT && nq >= 16 && k <= 32 && s->row_stride <= 128 && s->rows >= 4096);
    if (!plain && (rc = slab_exclusive_begin(s, st, slot))) return rc;
    if (tci8)
        rc = knn_tci8(s, d_q, nq, k, metric, d_out_rowids, d_out_dists, d_out_counts, pad_rowid, st);
    else if (tc)
        rc = knn_tc(s, d_q, nq, k, metric, d_out_rowids, d_out_dists, d_out_counts, pad_rowid, st);
    else
        rc = knn_exact(s, d_q, nq, k, metric, d_out_rowids, d_out_dists, d_out_counts, pad_rowid, st, plain ? slot : 0);
    if (!plain && !rc) rc = slab_exclusive_end(s, st, slot);
    s->q_uploaded = true;  // whatever was staged has been consumed
    s->h_q_staged = nullptr;
    return rc;
}

// stage nq host queries (row_bytes each), padded to row_stride, in pinned memory; upload = also copy them into the slab's
// device query buffer now (otherwise query_upload_if_needed does it when a path asks for device-resident queries)
static int stage_queries(vecgpu_slab* s, const void* queries, uint32_t nq, bool upload = true) {
    const size_t bytes = (size_t)nq * s->row_stride;
    int rc = pin_reserve(s, 0, bytes);
    if (rc) return rc;
    rc = ws_reserve(s, WS_QUERY, bytes);
    if (rc) return rc;
    uint8_t* h = (uint8_t*)s->h_pin[0];
    if (s->row_bytes == s->row_stride) {
        memcpy(h, queries, bytes);
    } else {
        memset(h, 0, bytes);
        for (uint32_t q = 0; q < nq; ++q)
            memcpy(h + (size_t)q * s->row_stride, (const uint8_t*)queries + (size_t)q * s->row_bytes, s->row_bytes);
    }
    s->h_q_staged = h;
    s->h_q_bytes = bytes;
    s->q_uploaded = false;
    if (upload) {
        CU(cudaMemcpyAsync(s->d_ws[WS_QUERY], h, bytes, cudaMemcpyHostToDevice, s->stream));
        s->q_uploaded = true;
    }
    return 0;
}
static int query_upload_if_needed(vecgpu_slab* s, cudaStream_t st) {
    if (s->q_uploaded) return 0;
    CU(cudaMemcpyAsync(s->d_ws[WS_QUERY], s->h_q_staged, s->h_q_bytes, cudaMemcpyHostToDevice, st));
    s->q_uploaded = true;
    return 0;
}

extern "C" int vecgpu_knn(vecgpu_slab* s, const void* queries, uint32_t nq, uint32_t k, int metric, int64_t* out_rowids,
                          float* out_dists, uint32_t* out_counts) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    int rc = check_pair(s->elem, metric);
    if (rc) return rc;
    if (nq == 0 || k == 0) {
        if (out_counts)
            for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return 0;
    }
    if (!queries || !out_rowids || !out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    rc = use_device(s->device);
    if (rc) return rc;
    const size_t n_out = (size_t)nq * k;
    rc = stage_queries(s, queries, nq, /*upload=*/false);
    if (rc) return rc;
    const size_t pin_bytes = n_out * 12 + (size_t)nq * 4;
    if ((rc = pin_reserve(s, 1, pin_bytes))) return rc;
    uint8_t* h = (uint8_t*)s->h_pin[1];
    if (n_out <= DIRECT_OUT_MAX && k <= K_FUSED_MAX) {
        // small results: the merge writes them straight into this pinned (device-mapped) buffer over PCIe — no device->host
        // copies in the stream, the host only waits for the stream (was three dependent cudaMemcpyAsync, ~25 us)
        rc = knn_core(s, (const uint8_t*)s->d_ws[WS_QUERY], nq, k, metric, (int64_t*)h, (float*)(h + n_out * 8), (uint32_t*)(h + n_out * 12), -1,
                      s->stream);
        if (rc) return rc;
        CU(cudaStreamSynchronize(s->stream));
    } else {
        if ((rc = ws_reserve(s, WS_OUT_ROWID, n_out * 8))) return rc;
        if ((rc = ws_reserve(s, WS_OUT_DIST, n_out * 4))) return rc;
        if ((rc = ws_reserve(s, WS_OUT_CNT, (size_t)nq * 4))) return rc;
        rc = knn_core(s, (const uint8_t*)s->d_ws[WS_QUERY], nq, k, metric, (int64_t*)s->d_ws[WS_OUT_ROWID],
                      (float*)s->d_ws[WS_OUT_DIST], (uint32_t*)s->d_ws[WS_OUT_CNT], -1, s->stream);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h, s->d_ws[WS_OUT_ROWID], n_out * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h + n_out * 8, s->d_ws[WS_OUT_DIST], n_out * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h + n_out * 12, s->d_ws[WS_OUT_CNT], (size_t)nq * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    memcpy(out_rowids, h, n_out * 8);
    memcpy(out_dists, h + n_out * 8, n_out * 4);
    if (out_counts) memcpy(out_counts, h + n_out * 12, (size_t)nq * 4);
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_knn_device(vecgpu_slab* s, const void* d_queries, uint32_t nq, uint32_t k, int metric,
                                 int64_t* d_out_rowids, float* d_out_dists, void* stream) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    int rc = check_pair(s->elem, metric);
    if (rc) return rc;
    if (nq == 0 || k == 0) return 0;
    if (!d_queries || !d_out_rowids || !d_out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    rc = use_device(s->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;  // NULL == the CUDA default stream, as everywhere in CUDA
    const uint8_t* dq = (const uint8_t*)d_queries;
    if (s->row_bytes != s->row_stride) {
        int slot = 0;
        if ((rc = slab_pick_slot(s, st, &slot))) return rc;
        const int ws_q = slot ? WS_QUERY1 : WS_QUERY;
        rc = ws_reserve(s, ws_q, (size_t)nq * s->row_stride);
        if (rc) return rc;
        pad_rows_kernel<<<std::min<uint32_t>(1024, (nq * s->row_stride + 255) / 256), 256, 0, st>>>(
            dq, s->row_bytes, (uint8_t*)s->d_ws[ws_q], s->row_stride, nq);
        LAUNCHED();
        dq = (const uint8_t*)s->d_ws[ws_q];
    }
    return knn_core(s, dq, nq, k, metric, d_out_rowids, d_out_dists, nullptr, INT64_MAX, st);
    VG_CATCH
}

extern "C" int vecgpu_merge_device(int device, const int64_t* d_rowids, const float* d_dists, uint32_t nlists, uint32_t nq,
                                   uint32_t k, int64_t* d_out_rowids, float* d_out_dists, void* stream) {
    VG_TRY
    if (!d_rowids || !d_dists || !d_out_rowids || !d_out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (nq == 0 || k == 0 || nlists == 0) return 0;
    int rc = use_device(device);
    if (rc) return rc;
    XMergeParams p{};
    p.rowids = d_rowids;
    p.dists = d_dists;
    p.nlists = nlists;
    p.nq = nq;
    p.k = k;
    if ((uint64_t)nlists * k > 16384) return fail(VECGPU_ERR_INVALID_PARAM, "nlists*k must be <= 16384");
    p.kp2 = std::max(2u, next_pow2(nlists * k));
    p.out_rowids = d_out_rowids;
    p.out_dists = d_out_dists;
    const size_t smem = (size_t)p.kp2 * 12;
    CU(cudaFuncSetAttribute(xmerge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    xmerge_kernel<<<nq, 256, smem, (cudaStream_t)stream>>>(p);
    LAUNCHED();
    return 0;
    VG_CATCH
}

// ---------------------------------------------------------------------------
// pair scoring
// ---------------------------------------------------------------------------
template <class T>
static int launch_pairs_t(const PairParams& p, int num_sms, cudaStream_t st) {
    const uint64_t gpb = 256 / T::LPR;
    const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>((p.n_pairs + gpb - 1) / gpb, (uint64_t)num_sms * 8));
    pair_kernel<T><<<(uint32_t)blocks, 256, 0, st>>>(p);
    LAUNCHED();
    return 0;
}
static int launch_pairs(int elem, int metric, const PairParams& p, int num_sms, cudaStream_t st) {
    if (elem == VECGPU_F32) {
        if (metric == VECGPU_L2) return launch_pairs_t<F32L2<1>>(p, num_sms, st);
        if (metric == VECGPU_L1) return launch_pairs_t<F32L1<1>>(p, num_sms, st);
        return launch_pairs_t<F32Cos<1>>(p, num_sms, st);
    }
    if (elem == VECGPU_I8) {
        if (metric == VECGPU_L2) return launch_pairs_t<I8Dot<1, false>>(p, num_sms, st);
        if (metric == VECGPU_L1) return launch_pairs_t<I8L1<1>>(p, num_sms, st);
        return launch_pairs_t<I8Dot<1, true>>(p, num_sms, st);
    }
    return launch_pairs_t<BitHamming<1>>(p, num_sms, st);
}

template <class T>
static int launch_score_small_t(const ScoreSmallParams& p, cudaStream_t st) {
    constexpr uint32_t gpb = 256 / T::LPR;
    score_small_kernel<T><<<(p.np + gpb - 1) / gpb, 256, p.in_bytes, st>>>(p);
    LAUNCHED();
    return 0;
}
static int launch_score_small(int elem, int metric, const ScoreSmallParams& p, cudaStream_t st) {
    if (elem == VECGPU_F32) {
        if (metric == VECGPU_L2) return launch_score_small_t<F32L2<1>>(p, st);
        if (metric == VECGPU_L1) return launch_score_small_t<F32L1<1>>(p, st);
        return launch_score_small_t<F32Cos<1>>(p, st);
    }
    if (elem == VECGPU_I8) {
        if (metric == VECGPU_L2) return launch_score_small_t<I8Dot<1, false>>(p, st);
        if (metric == VECGPU_L1) return launch_score_small_t<I8L1<1>>(p, st);
        return launch_score_small_t<I8Dot<1, true>>(p, st);
    }
    return launch_score_small_t<BitHamming<1>>(p, st);
}

extern "C" int vecgpu_score(vecgpu_slab* s, const void* queries, uint32_t nq, const int64_t* cand_rowids,
                            const uint32_t* cand_offsets, int metric, float* out_dists) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    int rc = check_pair(s->elem, metric);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!queries || !cand_offsets) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    for (uint32_t q = 0; q < nq; ++q)
        if (cand_offsets[q + 1] < cand_offsets[q]) return fail(VECGPU_ERR_INVALID_PARAM, "cand_offsets must be non-decreasing");
    if (cand_offsets[0] != 0) return fail(VECGPU_ERR_INVALID_PARAM, "cand_offsets[0] must be 0");
    const uint64_t np = cand_offsets[nq];
    if (np == 0) return 0;
    if (!cand_rowids || !out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    rc = use_device(s->device);
    if (rc) return rc;
    {
        // Small calls — one expansion of search_layer is a query and <= 64 rowids — run as ONE launch that reads its input
        // from pinned host memory and stores the distances there (score_small_kernel): ~3x less fixed cost than the
        // upload / resolve / score / download sequence below.
        const size_t q_off = ((size_t)np * 8 + (size_t)(nq + 1) * 4 + 15) & ~(size_t)15;  // the queries start 16-byte aligned
        const size_t in_bytes = q_off + (size_t)nq * s->row_stride, out_off = (in_bytes + 255) & ~(size_t)255;
        if (np <= 256 && in_bytes <= 16384) {
            if (env_u32("VECGPU_SCORE_SMALL", 1)) {
                if ((rc = pin_reserve(s, 1, out_off + (size_t)np * 4))) return rc;
                uint8_t* h = (uint8_t*)s->h_pin[1];
                memset(h, 0, in_bytes);
                memcpy(h, cand_rowids, np * 8);
                memcpy(h + np * 8, cand_offsets, (size_t)(nq + 1) * 4);
                const uint32_t rbytes = vecgpu_row_bytes(s->elem, s->dims);
                for (uint32_t q = 0; q < nq; ++q) memcpy(h + q_off + (size_t)q * s->row_stride, (const uint8_t*)queries + (size_t)q * rbytes, rbytes);
                if ((rc = slab_sync_rowids(s))) return rc;
                if ((rc = slab_sync_skip(s))) return rc;
                ScoreSmallParams sp{};
                sp.in = h;
                sp.in_bytes = (uint32_t)((in_bytes + 15) & ~(size_t)15);
                sp.np = (uint32_t)np;
                sp.nq = nq;
                sp.q_off = (uint32_t)q_off;
                sp.rowids = s->dense ? nullptr : s->d_rowids;
                sp.n_rows = s->rows;
                sp.first_rowid = s->first_rowid;
                sp.skip = s->n_skip ? s->d_skip : nullptr;
                sp.b_base = s->d_vec;
                sp.stride = s->row_stride;
                sp.units = s->row_stride / 16;
                sp.qc_kind = metric_qc_kind(s->elem);
                sp.out = (float*)(h + out_off);
                if ((rc = launch_score_small(s->elem, metric, sp, s->stream))) return rc;
                CU(cudaStreamSynchronize(s->stream));
                memcpy(out_dists, h + out_off, np * 4);
                return 0;
            }
        }
    }
    rc = stage_queries(s, queries, nq);
    if (rc) return rc;
    // candidates + offsets in one pinned upload: [rowids np*8][offsets (nq+1)*4]
    const size_t up_bytes = np * 8 + (size_t)(nq + 1) * 4;
    if ((rc = pin_reserve(s, 1, std::max(up_bytes, (size_t)np * 4)))) return rc;
    if ((rc = ws_reserve(s, WS_TMP, up_bytes))) return rc;
    if ((rc = ws_reserve(s, WS_TMP2, np * 12))) return rc;   // [pos np*8][qidx np*4]
    if ((rc = ws_reserve(s, WS_OUT_DIST, np * 4))) return rc;
    uint8_t* h = (uint8_t*)s->h_pin[1];
    memcpy(h, cand_rowids, np * 8);
    memcpy(h + np * 8, cand_offsets, (size_t)(nq + 1) * 4);
    CU(cudaMemcpyAsync(s->d_ws[WS_TMP], h, up_bytes, cudaMemcpyHostToDevice, s->stream));
    int64_t* d_pos = (int64_t*)s->d_ws[WS_TMP2];
    uint32_t* d_qidx = (uint32_t*)((uint8_t*)s->d_ws[WS_TMP2] + np * 8);
    const uint32_t rb = (uint32_t)std::min<uint64_t>((np + 255) / 256, 1024);
    resolve_kernel<<<rb, 256, 0, s->stream>>>((const int64_t*)s->d_ws[WS_TMP], np,
                                              (const uint32_t*)((uint8_t*)s->d_ws[WS_TMP] + np * 8), nq,
                                              s->dense ? nullptr : s->d_rowids, s->rows, s->first_rowid,
                                              s->n_skip ? s->d_skip : nullptr, d_qidx, d_pos);
    LAUNCHED();
    PairParams p{};
    p.a_base = (const uint8_t*)s->d_ws[WS_QUERY];
    p.b_base = s->d_vec;
    p.a_stride = p.b_stride = s->row_stride;
    p.units = s->row_stride / 16;
    p.a_index = d_qidx;
    p.b_index = d_pos;
    p.n_pairs = np;
    p.out = (float*)s->d_ws[WS_OUT_DIST];
    p.qc_kind = metric_qc_kind(s->elem);
    rc = launch_pairs(s->elem, metric, p, s->num_sms, s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h, s->d_ws[WS_OUT_DIST], np * 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(out_dists, h, np * 4);
    return 0;
    VG_CATCH
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

extern "C" int vecgpu_distance_pairs(int elem, uint32_t dims_a, uint32_t dims_b, const void* a, const void* b, uint64_t n,
                                     int metric, int device, float* out) {
    VG_TRY
    if (elem < 0 || elem > 2) return fail(VECGPU_ERR_UNSUPPORTED, "invalid vector type %d", elem);
    // order of checks follows src/distance/mod.rs:57-83: dimensions, (types), then the metric match
    if (dims_a != dims_b) return fail(VECGPU_ERR_DIM_MISMATCH, "Dimension mismatch: expected %u, got %u", dims_a, dims_b);
    int rc = check_pair(elem, metric);
    if (rc) return rc;
    if (dims_a == 0 || dims_a > DIMS_MAX) return fail(VECGPU_ERR_INVALID_PARAM, "dims must be in 1..%u", DIMS_MAX);
    if (n == 0) return 0;
    if (!a || !b || !out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    rc = use_device(device);
    if (rc) return rc;
    const uint32_t rbytes = vecgpu_row_bytes(elem, dims_a), stride = (rbytes + 15u) & ~15u;
    DevBuf da, db, dout;
    CU(cudaMalloc(&da.p, n * stride));
    CU(cudaMalloc(&db.p, n * stride));
    CU(cudaMalloc(&dout.p, n * 4));
    if (stride == rbytes) {
        CU(cudaMemcpy(da.p, a, n * stride, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(db.p, b, n * stride, cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemset(da.p, 0, n * stride));
        CU(cudaMemset(db.p, 0, n * stride));
        CU(cudaMemcpy2D(da.p, stride, a, rbytes, rbytes, n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy2D(db.p, stride, b, rbytes, rbytes, n, cudaMemcpyHostToDevice));
    }
    PairParams p{};
    p.a_base = (const uint8_t*)da.p;
    p.b_base = (const uint8_t*)db.p;
    p.a_stride = p.b_stride = stride;
    p.units = stride / 16;
    p.n_pairs = n;
    p.out = (float*)dout.p;
    p.qc_kind = metric_qc_kind(elem);
    rc = launch_pairs(elem, metric, p, 148, 0);
    if (rc) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
    VG_CATCH
}

// ---------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------
static int producer_common(const float* in, uint64_t n, uint32_t dims, int device, DevBuf& din) {
    if (!in) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (dims == 0 || dims > DIMS_MAX) return fail(VECGPU_ERR_INVALID_PARAM, "dims must be in 1..%u", DIMS_MAX);
    int rc = use_device(device);
    if (rc) return rc;
    CU(cudaMalloc(&din.p, n * dims * 4));
    CU(cudaMemcpy(din.p, in, n * dims * 4, cudaMemcpyHostToDevice));
    return 0;
}
static uint32_t blocks_for(uint64_t threads) { return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((threads + 255) / 256, 148 * 16)); }

extern "C" int vecgpu_normalize_f32(const float* in, uint64_t n, uint32_t dims, int device, float* out) {
    VG_TRY
    if (n == 0) return 0;
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    DevBuf din, dout, dflag;
    int rc = producer_common(in, n, dims, device, din);
    if (rc) return rc;
    CU(cudaMalloc(&dout.p, n * dims * 4));
    CU(cudaMalloc(&dflag.p, 4));
    CU(cudaMemset(dflag.p, 0, 4));
    normalize_kernel<<<blocks_for(n), 256>>>((const float*)din.p, n, dims, (float*)dout.p, (int*)dflag.p);
    LAUNCHED();
    int flag = 0;
    CU(cudaMemcpy(&flag, dflag.p, 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out, dout.p, n * dims * 4, cudaMemcpyDeviceToHost));
    if (flag) return fail(VECGPU_ERR_INVALID_PARAM, "Cannot normalize zero vector");  // src/vector.rs:451-455
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_quantize_int8(const float* in, uint64_t n, uint32_t dims, int device, int8_t* out) {
    VG_TRY
    if (n == 0) return 0;
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    DevBuf din, dout;
    int rc = producer_common(in, n, dims, device, din);
    if (rc) return rc;
    CU(cudaMalloc(&dout.p, n * dims));
    quantize_int8_kernel<<<blocks_for(n * 32), 256>>>((const float*)din.p, n, dims, (int8_t*)dout.p, dims);
    LAUNCHED();
    CU(cudaMemcpy(out, dout.p, n * dims, cudaMemcpyDeviceToHost));
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_quantize_int8_for_index(const float* in, uint64_t n, uint32_t dims, int device, int8_t* out) {
    VG_TRY
    if (n == 0) return 0;
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    DevBuf din, dout;
    int rc = producer_common(in, n, dims, device, din);
    if (rc) return rc;
    CU(cudaMalloc(&dout.p, n * dims));
    quantize_index_kernel<<<blocks_for(n * dims), 256>>>((const float*)din.p, n * dims, (int8_t*)dout.p);
    LAUNCHED();
    CU(cudaMemcpy(out, dout.p, n * dims, cudaMemcpyDeviceToHost));
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_quantize_binary(const float* in, uint64_t n, uint32_t dims, int device, uint8_t* out) {
    VG_TRY
    if (n == 0) return 0;
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    DevBuf din, dout;
    int rc = producer_common(in, n, dims, device, din);
    if (rc) return rc;
    const uint32_t nb = (dims + 7) / 8;
    CU(cudaMalloc(&dout.p, n * nb));
    quantize_binary_kernel<<<blocks_for(n), 256>>>((const float*)din.p, n, dims, (uint8_t*)dout.p);
    LAUNCHED();
    CU(cudaMemcpy(out, dout.p, n * nb, cudaMemcpyDeviceToHost));
    return 0;
    VG_CATCH
}

// ---------------------------------------------------------------------------
// synthetic fill + device view
// ---------------------------------------------------------------------------
extern "C" int vecgpu_slab_fill_synthetic(vecgpu_slab* s, uint64_t seed, int64_t first_rowid, uint64_t n, int kind) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    if (kind != VECGPU_SYNTH_UNIFORM && !(kind == VECGPU_SYNTH_GAUSS4 && s->elem == VECGPU_F32))
        return fail(VECGPU_ERR_INVALID_PARAM, "unsupported synthetic kind %d for this element type", kind);
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    slab_reset_rows(s);  // bumps layout_gen: an HNSW index built over the old rows fails loudly until rebuilt
    rc = slab_reserve_rows(s, n);
    if (rc) return rc;
    s->first_rowid = first_rowid;
    s->rows = n;
    if (n == 0) return 0;
    const uint32_t grid = (uint32_t)s->num_sms * 16;
    if (s->elem == VECGPU_F32)
        synth_f32_kernel<<<grid, 256, 0, s->stream>>>((float*)s->d_vec, s->row_stride / 4, s->dims, n, seed, first_rowid, kind);
    else if (s->elem == VECGPU_I8)
        synth_i8_kernel<<<grid, 256, 0, s->stream>>>((int8_t*)s->d_vec, s->row_stride, s->dims, n, seed, first_rowid);
    else
        synth_bit_kernel<<<grid, 256, 0, s->stream>>>(s->d_vec, s->row_stride, s->dims, n, seed, first_rowid);
    LAUNCHED();
    CU(cudaStreamSynchronize(s->stream));
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_slab_device_view(vecgpu_slab* s, void** d_vectors, uint32_t* row_stride, uint64_t* rows) {
    VG_TRY
    if (!s) return fail(VECGPU_ERR_INVALID_PARAM, "slab is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    if (d_vectors) *d_vectors = s->d_vec;
    if (row_stride) *row_stride = s->row_stride;
    if (rows) *rows = s->rows;
    return 0;
    VG_CATCH
}

// ---------------------------------------------------------------------------
// batched HNSW driver (config 5)
// ---------------------------------------------------------------------------
#include "hnsw.inl"

// ---------------------------------------------------------------------------
// sharded slabs: exchange endpoints, per-rank and one-process entry points
// ---------------------------------------------------------------------------
#include "xchg_host.inl"
