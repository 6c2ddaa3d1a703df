// xchg_host.inl — host side of the sharded exact KNN behind the C ABI (included by vecgpu.cu).
//
//   vecgpu_xchg_*      one exchange endpoint per rank (GPU): gather buffer, peer mapping (same process or CUDA IPC)
//   vecgpu_shard_knn*  one rank's part of a sharded query: local scan -> push into every peer's buffer -> merge
//   vecgpu_sharded_*   ONE-PROCESS form for the Rust extension (the reference is one process per connection,
//                      src/lib.rs:26-34, src/vtab.rs:2286-2305): a slab cut by contiguous rowid range over several
//                      devices, one host worker thread + stream per device, a single call returns the global top-k.
// Device code: xchg.cuh.
#include <condition_variable>
#include <functional>
#include <thread>

struct vecgpu_xchg {
    int device = 0;
    uint32_t rank = 0, world = 1, cap_q = 0, cap_k = 0;
    XLayout lay{};
    uint8_t* d_buf = nullptr;               // this rank's gather buffer (+ flags)
    uint8_t* peer_base[XCHG_MAX_WORLD] = {nullptr};
    bool ipc_open[XCHG_MAX_WORLD] = {false};
    XPeerTable* d_tab = nullptr;
    uint32_t* d_done = nullptr;             // last-CTA counter of xpush
    uint32_t* d_err = nullptr;
    uint32_t* h_err = nullptr;              // pinned
    uint32_t epoch = 0;
    bool attached = false;
    std::mutex mu;
};

static constexpr uint32_t XCHG_HANDLE_MAGIC = 0x58434847u;  // "XCHG"
struct XchgHandleBlob {  // what vecgpu_xchg_export writes (VECGPU_XCHG_HANDLE_BYTES = 128)
    uint32_t magic, rank, world, cap_q;
    uint64_t cap_entries, total_bytes;
    int32_t device, pid;
    cudaIpcMemHandle_t ipc;  // 64 bytes
    uint8_t pad[128 - 40 - 64];
};
static_assert(sizeof(XchgHandleBlob) == 128, "handle blob must be 128 bytes");

extern "C" int vecgpu_xchg_create(int device, uint32_t rank, uint32_t world, uint32_t max_queries, uint32_t max_k, vecgpu_xchg** out) {
    VG_TRY
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "out is NULL");
    *out = nullptr;
    if (world == 0 || world > XCHG_MAX_WORLD || rank >= world) return fail(VECGPU_ERR_INVALID_PARAM, "need rank < world <= %u", XCHG_MAX_WORLD);
    if (max_queries == 0 || max_k == 0) return fail(VECGPU_ERR_INVALID_PARAM, "max_queries and max_k must be positive");
    int rc = use_device(device);
    if (rc) return rc;
    vecgpu_xchg* x = new (std::nothrow) vecgpu_xchg();
    if (!x) return fail(VECGPU_ERR_CUDA, "out of host memory");
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->cap_q = max_queries;
    x->cap_k = max_k;
    x->lay = xlayout(world, max_queries, (uint64_t)max_queries * max_k);
    auto bail = [&](const char* what) {
        cudaError_t e = cudaGetLastError();
        if (x->d_buf) cudaFree(x->d_buf);
        if (x->d_tab) cudaFree(x->d_tab);
        if (x->d_done) cudaFree(x->d_done);
        if (x->h_err) cudaFreeHost(x->h_err);
        delete x;
        return fail(VECGPU_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    if (cudaMalloc((void**)&x->d_buf, x->lay.total_bytes) != cudaSuccess) return bail("cudaMalloc(gather buffer)");
    if (cudaMemset(x->d_buf, 0, x->lay.total_bytes) != cudaSuccess) return bail("cudaMemset");
    if (cudaMalloc((void**)&x->d_tab, sizeof(XPeerTable)) != cudaSuccess) return bail("cudaMalloc");
    if (cudaMalloc((void**)&x->d_done, 8) != cudaSuccess) return bail("cudaMalloc");
    if (cudaMemset(x->d_done, 0, 8) != cudaSuccess) return bail("cudaMemset");
    if (cudaMallocHost((void**)&x->h_err, 64) != cudaSuccess) return bail("cudaMallocHost");
    *x->h_err = 0;
    x->d_err = x->h_err;  // pinned + device-mapped (UVA): the merge kernel stores into it directly, the host reads it after a sync
    x->peer_base[rank] = x->d_buf;
    if (world == 1) {
        XPeerTable t{};
        t.base[0] = x->d_buf;
        if (cudaMemcpy(x->d_tab, &t, sizeof(t), cudaMemcpyHostToDevice) != cudaSuccess) return bail("cudaMemcpy");
        x->attached = true;
    }
    CU(cudaDeviceSynchronize());
    *out = x;
    return 0;
    VG_CATCH
}

extern "C" void vecgpu_xchg_destroy(vecgpu_xchg* x) {
    VG_TRY
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (uint32_t p = 0; p < x->world; ++p)
        if (x->ipc_open[p] && x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    cudaFree(x->d_buf);
    cudaFree(x->d_tab);
    cudaFree(x->d_done);
    if (x->h_err) cudaFreeHost(x->h_err);
    cudaGetLastError();
    delete x;
    VG_CATCH_VOID
}

extern "C" int vecgpu_xchg_export(vecgpu_xchg* x, void* handle128) {
    VG_TRY
    if (!x || !handle128) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    int rc = use_device(x->device);
    if (rc) return rc;
    XchgHandleBlob b{};
    b.magic = XCHG_HANDLE_MAGIC;
    b.rank = x->rank;
    b.world = x->world;
    b.cap_q = x->cap_q;
    b.cap_entries = x->lay.cap_entries;
    b.total_bytes = x->lay.total_bytes;
    b.device = x->device;
    b.pid = (int32_t)getpid();
    CU(cudaIpcGetMemHandle(&b.ipc, x->d_buf));
    memcpy(handle128, &b, sizeof(b));
    return 0;
    VG_CATCH
}

static int xchg_finish_attach(vecgpu_xchg* x) {
    XPeerTable t{};
    for (uint32_t p = 0; p < x->world; ++p) t.base[p] = x->peer_base[p];
    CU(cudaMemcpy(x->d_tab, &t, sizeof(t), cudaMemcpyHostToDevice));
    x->attached = true;
    return 0;
}

// Peers in OTHER processes (torchrun: one process per GPU): handles[p] is what rank p's vecgpu_xchg_export wrote,
// gathered by the caller over any channel (bench.py uses torch.distributed).  Every rank must attach before the first
// exchange; the caller separates attach and first use with a barrier.
extern "C" int vecgpu_xchg_attach_ipc(vecgpu_xchg* x, const void* handles, uint32_t n_handles) {
    VG_TRY
    if (!x || !handles) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (n_handles != x->world) return fail(VECGPU_ERR_INVALID_PARAM, "expected %u handles, got %u", x->world, n_handles);
    std::lock_guard<std::mutex> lk(x->mu);
    int rc = use_device(x->device);
    if (rc) return rc;
    const XchgHandleBlob* hb = (const XchgHandleBlob*)handles;
    for (uint32_t p = 0; p < x->world; ++p) {
        XchgHandleBlob b;
        memcpy(&b, hb + p, sizeof(b));
        if (b.magic != XCHG_HANDLE_MAGIC || b.rank != p || b.world != x->world || b.cap_q != x->cap_q ||
            b.cap_entries != x->lay.cap_entries || b.total_bytes != x->lay.total_bytes)
            return fail(VECGPU_ERR_INVALID_PARAM, "handle %u does not describe rank %u of a %u-rank exchange with the same capacities", p, p, x->world);
        if (p == x->rank) continue;
        if (x->ipc_open[p]) continue;
        if (b.pid == (int32_t)getpid()) {
            return fail(VECGPU_ERR_INVALID_PARAM, "rank %u lives in this process: use vecgpu_xchg_attach_local", p);
        }
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess));
        x->peer_base[p] = (uint8_t*)ptr;
        x->ipc_open[p] = true;
    }
    return xchg_finish_attach(x);
    VG_CATCH
}

// Peers in THIS process: all[p] is rank p's endpoint.  Enables peer access between the devices.
extern "C" int vecgpu_xchg_attach_local(vecgpu_xchg* const* all, uint32_t n) {
    VG_TRY
    if (!all || n == 0) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    for (uint32_t p = 0; p < n; ++p) {
        if (!all[p] || all[p]->world != n || all[p]->rank != p) return fail(VECGPU_ERR_INVALID_PARAM, "all[%u] is not rank %u of %u", p, p, n);
        if (all[p]->lay.total_bytes != all[0]->lay.total_bytes || all[p]->cap_q != all[0]->cap_q)
            return fail(VECGPU_ERR_INVALID_PARAM, "endpoints were created with different capacities");
    }
    for (uint32_t a = 0; a < n; ++a) {
        vecgpu_xchg* x = all[a];
        std::lock_guard<std::mutex> lk(x->mu);
        int rc = use_device(x->device);
        if (rc) return rc;
        for (uint32_t p = 0; p < n; ++p) {
            if (p == a) continue;
            if (all[p]->device != x->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, x->device, all[p]->device));
                if (!can) return fail(VECGPU_ERR_CUDA, "device %d cannot access device %d (no peer path)", x->device, all[p]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(VECGPU_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", x->device, all[p]->device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            x->peer_base[p] = all[p]->d_buf;
        }
        rc = xchg_finish_attach(x);
        if (rc) return rc;
    }
    return 0;
    VG_CATCH
}

// The exchange of ONE batch = xchg_push (local top-k -> the peers' gather buffers, flags published) followed on the receiving
// ranks by xchg_merge (flag wait + k-way merge).  Both only enqueue on `st`.
static int xchg_push_prepare(vecgpu_xchg* x, const int64_t* d_rowids, const float* d_dists, const uint32_t* d_counts, uint32_t nq, uint32_t k,
                             int root, XPushParams* pp_out) {
    if (!x->attached) return fail(VECGPU_ERR_INVALID_PARAM, "exchange endpoint is not attached to its peers yet");
    if (nq > x->cap_q || (uint64_t)nq * k > x->lay.cap_entries)
        return fail(VECGPU_ERR_INVALID_PARAM, "exchange of %u queries x k=%u exceeds the endpoint's capacity (%u queries, %llu entries)", nq, k,
                    x->cap_q, (unsigned long long)x->lay.cap_entries);
    if ((uint64_t)x->world * k > 16384) return fail(VECGPU_ERR_INVALID_PARAM, "world*k must be <= 16384");
    const uint32_t all = (1u << x->world) - 1u;
    XPushParams pp{};
    pp.tab = x->d_tab;
    pp.lay = x->lay;
    pp.src_rowids = d_rowids;
    pp.src_dists = d_dists;
    pp.src_counts = d_counts;
    pp.nq = nq;
    pp.k = k;
    pp.rank = x->rank;
    pp.epoch = ++x->epoch;
    pp.peer_mask = root < 0 ? all : (1u << root);
    pp.done = x->d_done;
    *pp_out = pp;
    return 0;
}
static int xchg_push_launch(const XPushParams& pp, cudaStream_t st) {
    xpush_kernel<<<std::min<uint32_t>(pp.nq, 296), 128, 0, st>>>(pp);
    LAUNCHED();
    return 0;
}
static int xchg_push(vecgpu_xchg* x, const int64_t* d_rowids, const float* d_dists, const uint32_t* d_counts, uint32_t nq, uint32_t k,
                     int root, uint32_t* epoch_out, cudaStream_t st) {
    XPushParams pp{};
    int rc = xchg_push_prepare(x, d_rowids, d_dists, d_counts, nq, k, root, &pp);
    if (rc) return rc;
    *epoch_out = pp.epoch;
    return xchg_push_launch(pp, st);
}

static int xchg_merge(vecgpu_xchg* x, uint32_t epoch, uint32_t nq, uint32_t k, int64_t pad_rowid, int64_t* d_out_rowids, float* d_out_dists,
                      uint32_t* d_out_counts, cudaStream_t st) {
    XWaitMergeParams wp{};
    wp.tab = x->d_tab;
    wp.lay = x->lay;
    wp.nq = nq;
    wp.k = k;
    wp.rank = x->rank;
    wp.epoch = epoch;
    wp.src_mask = (1u << x->world) - 1u;
    wp.np2 = std::max(2u, next_pow2(x->world * k));
    wp.pad_rowid = pad_rowid;
    wp.out_rowids = d_out_rowids;
    wp.out_dists = d_out_dists;
    wp.out_counts = d_out_counts;
    wp.err = x->d_err;
    wp.timeout_cycles = (unsigned long long)env_u32("VECGPU_XCHG_TIMEOUT_MS", 20000) * 1800000ull;  // ~1.8 GHz
    const size_t smem = x->world * k <= 256 ? 0 : (size_t)wp.np2 * 8;
    if (smem > 48 * 1024) {
        static int cfg_dev = -1;
        int dev = 0;
        CU(cudaGetDevice(&dev));
        if (cfg_dev != dev) {
            CU(cudaFuncSetAttribute(xwait_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            cfg_dev = dev;
        }
    }
    // one warp does all the work when world * k <= 256: a 32-thread CTA keeps the polling footprint per SM small
    xwait_merge_kernel<<<nq, x->world * k <= 256 ? 32 : 256, smem, st>>>(wp);
    LAUNCHED();
    return 0;
}

static int xchg_exchange(vecgpu_xchg* x, const int64_t* d_rowids, const float* d_dists, const uint32_t* d_counts, uint32_t nq, uint32_t k,
                         int64_t pad_rowid, int64_t* d_out_rowids, float* d_out_dists, uint32_t* d_out_counts, int root, cudaStream_t st) {
    uint32_t epoch = 0;
    int rc = xchg_push(x, d_rowids, d_dists, d_counts, nq, k, root, &epoch, st);
    if (rc) return rc;
    if (root >= 0 && (uint32_t)root != x->rank) return 0;  // gather-to-root: only the root merges
    return xchg_merge(x, epoch, nq, k, pad_rowid, d_out_rowids, d_out_dists, d_out_counts, st);
}

// In ONE process the shards' host threads can order the streams with CUDA events instead of letting the merge kernel poll:
// every worker records an event after its push and counts itself in; a worker that is about to merge waits (on the host,
// microseconds) until all pushes of this piece are ENQUEUED, then makes its stream wait for those events.  The flags are
// then already published when xwait_merge_kernel starts, so it never spins — which matters when several shards share one
// device (tests on a 1-GPU box): a polling kernel there can deadlock against a peer whose next kernel is held back by an
// implicit device-wide serialisation point (cudaMalloc / cudaFree of a growing workspace).
struct ShardSync {
    std::atomic<uint32_t>* pushes;  // total pushes enqueued in this call, all workers
    cudaEvent_t* events;            // [n] one per worker
    uint32_t n, me;
};

static int xchg_check_err(vecgpu_xchg* x, cudaStream_t st) {  // synchronises the stream, then reads the (host-resident) error word
    CU(cudaStreamSynchronize(st));
    if (*(volatile uint32_t*)x->h_err) {
        *x->h_err = 0;
        return fail(VECGPU_ERR_CUDA, "exchange timed out: a peer rank never delivered its top-k (did every rank issue the same call?)");
    }
    return 0;
}

// Exchange + merge of per-rank results that are already on the device (e.g. 16 vecgpu_knn_device calls, one exchange):
// d_counts may be NULL, in which case trailing (INT64_MAX, +inf) entries are padding.  Collective over the ranks.
extern "C" int vecgpu_xchg_merge_device(vecgpu_xchg* x, const int64_t* d_rowids, const float* d_dists, const uint32_t* d_counts, uint32_t nq,
                                        uint32_t k, int64_t* d_out_rowids, float* d_out_dists, void* stream) {
    VG_TRY
    if (!x || !d_rowids || !d_dists || !d_out_rowids || !d_out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (nq == 0 || k == 0) return 0;
    std::lock_guard<std::mutex> lk(x->mu);
    int rc = use_device(x->device);
    if (rc) return rc;
    return xchg_exchange(x, d_rowids, d_dists, d_counts, nq, k, INT64_MAX, d_out_rowids, d_out_dists, nullptr, -1, (cudaStream_t)stream);
    VG_CATCH
}

extern "C" int vecgpu_xchg_check(vecgpu_xchg* x, void* stream) {
    VG_TRY
    if (!x) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(x->mu);
    int rc = use_device(x->device);
    if (rc) return rc;
    return xchg_check_err(x, (cudaStream_t)stream);
    VG_CATCH
}

enum { XWS_ROWID = 0, XWS_DIST = 1, XWS_CNT = 2 };

// One rank's part of a sharded exact KNN, host buffers in and out: the call the Rust side (or one torchrun rank) makes.
// `root` < 0: every rank receives the global top-k; otherwise only rank `root` copies it out (the others' out arrays are
// untouched).  With a root and a batch that fits ONE exchange the other ranks only push (gather-to-root, nobody waits for
// anybody but the root); a batch that needs several exchanges is all-gathered so that the merges pace every rank (a rank
// that only pushed could overwrite a buffer half the root has not merged yet).
static int shard_knn_locked(vecgpu_slab* s, vecgpu_xchg* x, const void* queries, uint32_t nq, uint32_t k, int metric, int root,
                            int64_t* out_rowids, float* out_dists, uint32_t* out_counts, const ShardSync* sync = nullptr) {
    const bool receives = root < 0 || (uint32_t)root == x->rank;
    const uint32_t chunk = std::max(1u, std::min(x->cap_q, (uint32_t)std::min<uint64_t>(x->lay.cap_entries / k, 0xFFFFFFFFull)));
    if (nq > chunk) root = -1;
    const bool merges = root < 0 || (uint32_t)root == x->rank;
    const size_t n_out = (size_t)nq * k;
    // a worker that fails before its pushes still counts them in, so that no peer waits for it forever
    auto bail = [&](int code) {
        if (sync) sync->pushes->fetch_add((nq + chunk - 1) / chunk, std::memory_order_release);
        return code;
    };
    int rc = use_device(s->device);
    if (rc) return bail(rc);
    if ((rc = stage_queries(s, queries, nq, /*upload=*/false))) return bail(rc);
    if ((rc = ws_reserve(s, WS_OUT_ROWID, n_out * 8))) return bail(rc);
    if ((rc = ws_reserve(s, WS_OUT_DIST, n_out * 4))) return bail(rc);
    if ((rc = ws_reserve(s, WS_OUT_CNT, (size_t)nq * 4))) return bail(rc);
    if ((rc = ws_reserve(s, WS_X_ROWID, n_out * 8))) return bail(rc);
    if ((rc = ws_reserve(s, WS_X_DIST, n_out * 4))) return bail(rc);
    if ((rc = ws_reserve(s, WS_X_CNT, (size_t)nq * 4))) return bail(rc);
    int64_t* lr = (int64_t*)s->d_ws[WS_X_ROWID];
    float* ld = (float*)s->d_ws[WS_X_DIST];
    uint32_t* lc = (uint32_t*)s->d_ws[WS_X_CNT];
    // final results: small ones are written by the merge kernel straight into mapped pinned memory (no D2H copies)
    const size_t pin_bytes = n_out * 12 + (size_t)nq * 4;
    if ((rc = pin_reserve(s, 1, pin_bytes))) return bail(rc);
    uint8_t* h = (uint8_t*)s->h_pin[1];
    const bool direct = n_out <= DIRECT_OUT_MAX;
    int64_t* fr = direct ? (int64_t*)h : (int64_t*)s->d_ws[WS_OUT_ROWID];
    float* fd = direct ? (float*)(h + n_out * 8) : (float*)s->d_ws[WS_OUT_DIST];
    uint32_t* fc = direct ? (uint32_t*)(h + n_out * 12) : (uint32_t*)s->d_ws[WS_OUT_CNT];
    // local scan of every query first (one launch sequence), then the exchanges piece by piece.  A batch that is ONE piece and
    // ONE query pass lets the scan's last CTA carry the push (scan_kernel's fused tail): no separate merge / push launches.
    XPushParams pp0{};
    const bool one_piece = nq <= chunk;
    if (one_piece) {
        if ((rc = xchg_push_prepare(x, lr, ld, lc, nq, k, root, &pp0))) return bail(rc);
        s->fuse_push = &pp0;
        s->fuse_push_done = false;
    }
    rc = knn_core(s, (const uint8_t*)s->d_ws[WS_QUERY], nq, k, metric, lr, ld, lc, -1, s->stream);
    const bool pushed_by_scan = s->fuse_push_done;
    s->fuse_push = nullptr;
    s->fuse_push_done = false;
    uint32_t piece = 0;
    for (uint32_t q0 = 0; q0 < nq; q0 += chunk, ++piece) {
        const uint32_t m = std::min(chunk, nq - q0);
        uint32_t epoch = 0;
        if (one_piece) {
            epoch = pp0.epoch;
            if (!rc && !pushed_by_scan) rc = xchg_push_launch(pp0, s->stream);
        } else if (!rc) {
            rc = xchg_push(x, lr + (size_t)q0 * k, ld + (size_t)q0 * k, lc + q0, m, k, root, &epoch, s->stream);
        }
        if (sync) {
            // count in even after a failure: the peers must not wait for this worker forever
            if (!rc && cudaEventRecord(sync->events[sync->me], s->stream) != cudaSuccess) rc = fail(VECGPU_ERR_CUDA, "cudaEventRecord failed");
            sync->pushes->fetch_add(1, std::memory_order_release);
            if (merges) {
                const uint32_t want = (piece + 1) * sync->n;
                while (sync->pushes->load(std::memory_order_acquire) < want) std::this_thread::yield();
                for (uint32_t p = 0; p < sync->n && !rc; ++p)
                    if (p != sync->me && cudaStreamWaitEvent(s->stream, sync->events[p], 0) != cudaSuccess)
                        rc = fail(VECGPU_ERR_CUDA, "cudaStreamWaitEvent failed");
            }
        }
        if (!rc && merges) rc = xchg_merge(x, epoch, m, k, -1, fr + (size_t)q0 * k, fd + (size_t)q0 * k, fc + q0, s->stream);
    }
    if (rc) {
        cudaStreamSynchronize(s->stream);
        return rc;
    }
    if (!receives) return xchg_check_err(x, s->stream);
    if (!direct) {
        CU(cudaMemcpyAsync(h, s->d_ws[WS_OUT_ROWID], n_out * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h + n_out * 8, s->d_ws[WS_OUT_DIST], n_out * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h + n_out * 12, s->d_ws[WS_OUT_CNT], (size_t)nq * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    if ((rc = xchg_check_err(x, s->stream))) return rc;
    memcpy(out_rowids, h, n_out * 8);
    memcpy(out_dists, h + n_out * 8, n_out * 4);
    if (out_counts) memcpy(out_counts, h + n_out * 12, (size_t)nq * 4);
    return 0;
}

extern "C" int vecgpu_shard_knn(vecgpu_slab* s, vecgpu_xchg* x, const void* queries, uint32_t nq, uint32_t k, int metric,
                                int64_t* out_rowids, float* out_dists, uint32_t* out_counts) {
    VG_TRY
    if (!s || !x) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    int rc = check_pair(s->elem, metric);
    if (rc) return rc;
    if (x->device != s->device) return fail(VECGPU_ERR_INVALID_PARAM, "the slab and the exchange endpoint live on different devices");
    if (nq == 0 || k == 0) {
        if (out_counts)
            for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return 0;
    }
    if (!queries || !out_rowids || !out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lk2(x->mu);
    return shard_knn_locked(s, x, queries, nq, k, metric, -1, out_rowids, out_dists, out_counts);
    VG_CATCH
}

// Device-resident form (bench `value`): d_queries / outputs on the slab's device, only enqueues on `stream`.
extern "C" int vecgpu_shard_knn_device(vecgpu_slab* s, vecgpu_xchg* x, const void* d_queries, uint32_t nq, uint32_t k, int metric,
                                       int64_t* d_out_rowids, float* d_out_dists, void* stream) {
    VG_TRY
    if (!s || !x) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    int rc = check_pair(s->elem, metric);
    if (rc) return rc;
    if (nq == 0 || k == 0) return 0;
    if (!d_queries || !d_out_rowids || !d_out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (nq > x->cap_q || (uint64_t)nq * k > x->lay.cap_entries) return fail(VECGPU_ERR_INVALID_PARAM, "batch exceeds the exchange capacity");
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lk2(x->mu);
    rc = use_device(s->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t* dq = (const uint8_t*)d_queries;
    if (s->row_bytes != s->row_stride) {
        if ((rc = ws_reserve(s, WS_QUERY, (size_t)nq * s->row_stride))) return rc;
        pad_rows_kernel<<<std::min<uint32_t>(1024, (nq * s->row_stride + 255) / 256), 256, 0, st>>>(dq, s->row_bytes, (uint8_t*)s->d_ws[WS_QUERY],
                                                                                                  s->row_stride, nq);
        LAUNCHED();
        dq = (const uint8_t*)s->d_ws[WS_QUERY];
    }
    const size_t n_out = (size_t)nq * k;
    if ((rc = ws_reserve(s, WS_X_ROWID, n_out * 8))) return rc;
    if ((rc = ws_reserve(s, WS_X_DIST, n_out * 4))) return rc;
    if ((rc = ws_reserve(s, WS_X_CNT, (size_t)nq * 4))) return rc;
    rc = knn_core(s, dq, nq, k, metric, (int64_t*)s->d_ws[WS_X_ROWID], (float*)s->d_ws[WS_X_DIST], (uint32_t*)s->d_ws[WS_X_CNT], INT64_MAX, st);
    if (rc) return rc;
    return xchg_exchange(x, (int64_t*)s->d_ws[WS_X_ROWID], (float*)s->d_ws[WS_X_DIST], (uint32_t*)s->d_ws[WS_X_CNT], nq, k, INT64_MAX,
                         d_out_rowids, d_out_dists, nullptr, -1, st);
    VG_CATCH
}

// =====================================================================================================
// ONE-PROCESS sharded slab: the form the Rust extension would use on an 8-GPU box.
// =====================================================================================================

struct ShardWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    std::atomic<uint32_t> posted{0}, finished{0};
    bool quit = false;
    int rc = 0;
    char err[512] = "";
};

struct vecgpu_sharded {
    int elem = 0;
    uint32_t dims = 0, n = 0;
    std::vector<int> devices;
    std::vector<vecgpu_slab*> slabs;
    std::vector<vecgpu_xchg*> xs;
    std::vector<ShardWorker*> workers;
    std::vector<int64_t> lo_rowid;  // first rowid of shard i (INT64_MAX while empty); routes upserts / deletes
    std::vector<cudaEvent_t> events;  // one per shard: "my push of this piece is enqueued up to here"
    std::atomic<uint32_t> pushes{0};
    uint32_t cap_q = 0, cap_k = 0;
    std::mutex mu;
};

static void shard_worker_main(ShardWorker* w) {
    uint32_t seen = 0;
    for (;;) {
        // short spin first: a query costs ~0.6 ms at 8 GPUs, a condition-variable wake-up alone is 20-50 us
        bool got = false;
        for (int i = 0; i < 20000 && !got; ++i) got = w->posted.load(std::memory_order_acquire) != seen;
        if (!got) {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->quit || w->posted.load(std::memory_order_acquire) != seen; });
            if (w->quit && w->posted.load(std::memory_order_acquire) == seen) return;
        }
        seen = w->posted.load(std::memory_order_acquire);
        int rc;
        try {
            rc = w->job();
        } catch (...) {
            rc = vg_caught();
        }
        w->rc = rc;
        if (rc) snprintf(w->err, sizeof(w->err), "%s", g_err);  // the message is thread-local: hand it to the caller's thread
        w->finished.store(seen, std::memory_order_release);
    }
}

// run fn(i) on worker i for every shard, wait for all; first failure wins
static int sharded_run(vecgpu_sharded* g, const std::function<int(uint32_t)>& fn) {
    for (uint32_t i = 0; i < g->n; ++i) {
        ShardWorker* w = g->workers[i];
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->job = [&fn, i] { return fn(i); };
            w->posted.fetch_add(1, std::memory_order_release);
        }
        w->cv.notify_one();
    }
    int rc = 0;
    for (uint32_t i = 0; i < g->n; ++i) {
        ShardWorker* w = g->workers[i];
        const uint32_t want = w->posted.load(std::memory_order_acquire);
        while (w->finished.load(std::memory_order_acquire) != want) std::this_thread::yield();
        if (w->rc && !rc) {
            rc = w->rc;
            snprintf(g_err, sizeof(g_err), "shard %u (device %d): %.400s", i, g->devices[i], w->err);
        }
    }
    return rc;
}

extern "C" void vecgpu_sharded_destroy(vecgpu_sharded* g) {
    VG_TRY
    if (!g) return;
    for (ShardWorker* w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_one();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t i = 0; i < g->events.size(); ++i) {
        cudaSetDevice(g->devices[i]);
        cudaEventDestroy(g->events[i]);
    }
    for (vecgpu_xchg* x : g->xs) vecgpu_xchg_destroy(x);
    for (vecgpu_slab* s : g->slabs) vecgpu_slab_destroy(s);
    cudaGetLastError();
    delete g;
    VG_CATCH_VOID
}

extern "C" int vecgpu_sharded_create(int elem, uint32_t dims, uint64_t capacity_hint_total, const int* devices, uint32_t n_devices,
                                     uint32_t max_queries, uint32_t max_k, vecgpu_sharded** out) {
    VG_TRY
    if (!out) return fail(VECGPU_ERR_INVALID_PARAM, "out is NULL");
    *out = nullptr;
    const int avail = vecgpu_device_count();
    if (avail <= 0) return fail(VECGPU_ERR_CUDA, "no CUDA device is available (libvecgpu has no CPU fallback)");
    if (n_devices == 0) n_devices = (uint32_t)avail;  // devices == NULL / n == 0: every visible device
    if (n_devices > XCHG_MAX_WORLD) return fail(VECGPU_ERR_INVALID_PARAM, "at most %u shards", XCHG_MAX_WORLD);
    if (max_queries == 0) max_queries = 1024;
    if (max_k == 0) max_k = 128;
    vecgpu_sharded* g = new vecgpu_sharded();
    g->elem = elem;
    g->dims = dims;
    g->n = n_devices;
    g->cap_q = max_queries;
    g->cap_k = max_k;
    int rc = 0;
    for (uint32_t i = 0; i < n_devices && !rc; ++i) {
        const int dev = devices ? devices[i] : (int)i;
        g->devices.push_back(dev);
        vecgpu_slab* s = nullptr;
        rc = vecgpu_slab_create(elem, dims, (capacity_hint_total + n_devices - 1) / n_devices, dev, &s);
        if (rc) break;
        g->slabs.push_back(s);
        vecgpu_xchg* x = nullptr;
        rc = vecgpu_xchg_create(dev, i, n_devices, max_queries, max_k, &x);
        if (rc) break;
        g->xs.push_back(x);
        g->lo_rowid.push_back(INT64_MAX);
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            rc = fail(VECGPU_ERR_CUDA, "cudaEventCreate failed");
            break;
        }
        g->events.push_back(ev);
    }
    if (!rc && n_devices > 1) rc = vecgpu_xchg_attach_local(g->xs.data(), n_devices);
    if (rc) {
        char keep[512];
        snprintf(keep, sizeof(keep), "%s", g_err);
        vecgpu_sharded_destroy(g);
        snprintf(g_err, sizeof(g_err), "%s", keep);
        return rc;
    }
    for (uint32_t i = 0; i < n_devices; ++i) {
        ShardWorker* w = new ShardWorker();
        g->workers.push_back(w);
        w->th = std::thread(shard_worker_main, w);
    }
    *out = g;
    return 0;
    VG_CATCH
}

static void sharded_range(uint64_t n, uint32_t i, uint32_t world, uint64_t* lo, uint64_t* hi) {  // == dist.shard_range
    const uint64_t base = n / world, rem = n % world;
    *lo = i * base + std::min<uint64_t>(i, rem);
    *hi = *lo + base + (i < rem ? 1 : 0);
}

// Bulk load: rows are cut into n_devices contiguous ranges in ascending rowid order (every rowid of shard i is below
// every rowid of shard i+1: the global (distance, rowid) order restricted to a shard is the shard's own order).
extern "C" int vecgpu_sharded_load(vecgpu_sharded* g, const int64_t* rowids, const void* vectors, uint64_t n) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    if (n && !vectors) return fail(VECGPU_ERR_INVALID_PARAM, "vectors is NULL");
    std::lock_guard<std::mutex> lk(g->mu);
    const uint32_t rb = vecgpu_row_bytes(g->elem, g->dims);
    std::vector<std::vector<int64_t>> dense_ids(g->n);
    for (uint32_t i = 0; i < g->n; ++i) {
        uint64_t lo, hi;
        sharded_range(n, i, g->n, &lo, &hi);
        g->lo_rowid[i] = hi > lo ? (rowids ? rowids[lo] : (int64_t)lo + 1) : INT64_MAX;
    }
    return sharded_run(g, [&](uint32_t i) {
        uint64_t lo, hi;
        sharded_range(n, i, g->n, &lo, &hi);
        if (rowids) return vecgpu_slab_load(g->slabs[i], rowids + lo, (const uint8_t*)vectors + lo * rb, hi - lo);
        // dense 1..n globally: shard i starts at rowid lo + 1
        int rc = vecgpu_slab_load(g->slabs[i], nullptr, nullptr, 0);
        if (rc || hi == lo) return rc;
        const int64_t first = (int64_t)lo + 1;
        std::vector<int64_t>& ids = dense_ids[i];
        ids.resize(hi - lo);
        for (uint64_t j = 0; j < hi - lo; ++j) ids[j] = first + (int64_t)j;
        return vecgpu_slab_append(g->slabs[i], ids.data(), (const uint8_t*)vectors + lo * rb, hi - lo);
    });
    VG_CATCH
}

extern "C" int vecgpu_sharded_fill_synthetic(vecgpu_sharded* g, uint64_t seed, int64_t first_rowid, uint64_t n, int kind) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    for (uint32_t i = 0; i < g->n; ++i) {
        uint64_t lo, hi;
        sharded_range(n, i, g->n, &lo, &hi);
        g->lo_rowid[i] = hi > lo ? first_rowid + (int64_t)lo : INT64_MAX;
    }
    return sharded_run(g, [&](uint32_t i) {
        uint64_t lo, hi;
        sharded_range(n, i, g->n, &lo, &hi);
        return vecgpu_slab_fill_synthetic(g->slabs[i], seed, first_rowid + (int64_t)lo, hi - lo, kind);
    });
    VG_CATCH
}

static uint32_t sharded_route(const vecgpu_sharded* g, int64_t rowid) {  // last shard whose first rowid is <= rowid
    uint32_t best = 0;
    bool any = false;
    for (uint32_t i = 0; i < g->n; ++i)
        if (g->lo_rowid[i] != INT64_MAX && g->lo_rowid[i] <= rowid) {
            best = i;
            any = true;
        }
    if (any) return best;
    for (uint32_t i = 0; i < g->n; ++i)
        if (g->lo_rowid[i] != INT64_MAX) return i;  // below every shard: the first non-empty one
    return 0;
}

// Vec0Tab::insert / update / delete hooks (src/vtab.rs:1409, 1684, 1326), routed to the shard that owns the rowid range.
extern "C" int vecgpu_sharded_upsert(vecgpu_sharded* g, int64_t rowid, const void* vec, uint32_t nbytes) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    const uint32_t i = sharded_route(g, rowid);
    int rc = vecgpu_slab_upsert(g->slabs[i], rowid, vec, nbytes);
    if (!rc && rowid < g->lo_rowid[i]) g->lo_rowid[i] = rowid;
    return rc;
    VG_CATCH
}
extern "C" int vecgpu_sharded_delete(vecgpu_sharded* g, int64_t rowid) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    return vecgpu_slab_delete(g->slabs[sharded_route(g, rowid)], rowid);
    VG_CATCH
}
extern "C" int vecgpu_sharded_count(vecgpu_sharded* g, uint64_t* rows, uint64_t* live) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    uint64_t r = 0, l = 0;
    for (uint32_t i = 0; i < g->n; ++i) {
        uint64_t a = 0, b = 0;
        int rc = vecgpu_slab_count(g->slabs[i], &a, &b);
        if (rc) return rc;
        r += a;
        l += b;
    }
    if (rows) *rows = r;
    if (live) *live = l;
    return 0;
    VG_CATCH
}
extern "C" int vecgpu_sharded_shard(vecgpu_sharded* g, uint32_t i, vecgpu_slab** slab, int* device) {
    VG_TRY
    if (!g || i >= g->n) return fail(VECGPU_ERR_INVALID_PARAM, "shard index out of range");
    if (slab) *slab = g->slabs[i];
    if (device) *device = g->devices[i];
    return 0;
    VG_CATCH
}
extern "C" uint32_t vecgpu_sharded_num_shards(vecgpu_sharded* g) { return g ? g->n : 0; }

// brute_force_search over the whole sharded slab (src/vtab.rs:2573-2623 semantics): every device scans its range, pushes its
// top-k into shard 0's gather buffer over NVLink, shard 0 merges and its result is returned.  One call, host buffers.
extern "C" int vecgpu_sharded_knn(vecgpu_sharded* g, const void* queries, uint32_t nq, uint32_t k, int metric, int64_t* out_rowids,
                                  float* out_dists, uint32_t* out_counts) {
    VG_TRY
    if (!g) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    int rc = check_pair(g->elem, metric);
    if (rc) return rc;
    if (nq == 0 || k == 0) {
        if (out_counts)
            for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return 0;
    }
    if (!queries || !out_rowids || !out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->n == 1) return vecgpu_knn(g->slabs[0], queries, nq, k, metric, out_rowids, out_dists, out_counts);
    if (k > g->cap_k) {
        // the gather buffers were sized for max_k: larger k goes through in query chunks as long as ONE query fits
        if ((uint64_t)k > g->xs[0]->lay.cap_entries || (uint64_t)g->n * k > 16384)
            return fail(VECGPU_ERR_INVALID_PARAM, "k=%u exceeds the sharded slab's exchange capacity (create it with a larger max_k)", k);
    }
    g->pushes.store(0, std::memory_order_release);
    return sharded_run(g, [&](uint32_t i) {
        vecgpu_slab* s = g->slabs[i];
        vecgpu_xchg* x = g->xs[i];
        std::lock_guard<std::mutex> l1(s->mu);
        std::lock_guard<std::mutex> l2(x->mu);
        ShardSync sync{&g->pushes, g->events.data(), g->n, i};
        return shard_knn_locked(s, x, queries, nq, k, metric, /*root=*/0, out_rowids, out_dists, out_counts, &sync);
    });
    VG_CATCH
}
