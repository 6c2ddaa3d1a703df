/*
 * vecgpu.h — C ABI of libvecgpu.so, the B200 (sm_100a) replacement for the
 * distance-scoring hot path of sqlite-vec-hnsw.
 *
 * Every entry point below states the reference interface (file:line under the
 * reference tree) it replaces.  Signatures are plain C: pointers + sizes, no
 * torch / C++ types.  All calls are synchronous unless a `stream` is passed
 * (the *_device variants), return 0 on success or a vecgpu_status code, and
 * never throw / unwind across the boundary (cf. catch_unwind, src/lib.rs:149).
 * The message for the last failing call on the calling thread is available
 * from vecgpu_last_error().
 *
 * There is NO CPU fallback behind this ABI: if no CUDA device is usable every
 * compute entry point fails with VECGPU_ERR_CUDA.
 */
#ifndef VECGPU_H
#define VECGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == VectorType order, src/vector.rs:9-16 */
enum vecgpu_elem { VECGPU_F32 = 0, VECGPU_I8 = 1, VECGPU_BIT = 2 };

/* == DistanceMetric order, src/distance/mod.rs:12-21 */
enum vecgpu_metric { VECGPU_L2 = 0, VECGPU_L1 = 1, VECGPU_COSINE = 2, VECGPU_HAMMING = 3 };

/* Error codes; the mapping to src/error.rs:5-36 is in INTEGRATION.md. */
enum vecgpu_status {
    VECGPU_OK = 0,
    VECGPU_ERR_INVALID_PARAM = 1, /* Error::InvalidParameter                       */
    VECGPU_ERR_DIM_MISMATCH = 2,  /* Error::DimensionMismatch (distance/mod.rs:57)  */
    VECGPU_ERR_UNSUPPORTED = 3,   /* Error::InvalidDistanceMetric / InvalidVectorType (distance/mod.rs:64-82) */
    VECGPU_ERR_CUDA = 4           /* Error::InvalidState: no device, OOM, launch failure */
};

/* Synthetic value distributions for vecgpu_slab_fill_synthetic. */
enum vecgpu_synth {
    VECGPU_SYNTH_UNIFORM = 0, /* f32: U[-1,1) on a 2^-23 grid; i8: quantize_int8(U[-1,1)); bit: Bernoulli(1/2) */
    VECGPU_SYNTH_GAUSS4 = 1   /* f32 only: Irwin-Hall(4) bell curve, exact integer arithmetic, std ~1.155   */
};

/* One slab = the HBM-resident, rowid-indexed copy of one vector column of one
 * vec0 table ({table}_data.vecNN, src/shadow.rs:111-129).  Opaque. */
typedef struct vecgpu_slab vecgpu_slab;

/* ---- library ---------------------------------------------------------- */

/* Thread-local message of the last failing call ("" if none). */
const char* vecgpu_last_error(void);
/* "vecgpu x.y.z sm_100a"; mirrors vec_version(), src/sql_functions.rs:13-48. */
const char* vecgpu_version(void);
/* Number of visible CUDA devices (0 when none; never fails). */
int vecgpu_device_count(void);
/* Bytes one stored row occupies in a blob: dims*4 | dims | ceil(dims/8)
 * (src/vector.rs:223-242, 592-600).  0 on bad elem type. */
uint32_t vecgpu_row_bytes(int elem_type, uint32_t dims);
/* 1 if (elem_type, metric) is one of the seven pairs distance() dispatches
 * (src/distance/mod.rs:70-83), else 0. */
int vecgpu_metric_supported(int elem_type, int metric);

/* ---- slab lifecycle: replaces per-row shadow::read_vector (src/shadow.rs:721-740)
 *      + get_all_rowids (src/shadow.rs:853-868) as the scan's data source ---- */

int vecgpu_slab_create(int elem_type, uint32_t dims, uint64_t capacity_hint, int device, vecgpu_slab** out);
void vecgpu_slab_destroy(vecgpu_slab* slab);

/* Bulk (re)load: replaces the slab contents with n rows.  `rowids` must be
 * strictly ascending (the order of "SELECT rowid FROM _data ORDER BY rowid",
 * src/shadow.rs:856) or NULL for dense 1..n.  `vectors` is row-major with
 * vecgpu_row_bytes() bytes per row, host memory. */
int vecgpu_slab_load(vecgpu_slab* slab, const int64_t* rowids, const void* vectors, uint64_t n);

/* Append n more rows whose rowids are all greater than every rowid already in
 * the slab (bulk staging in pieces).  rowids NULL => continue densely. */
int vecgpu_slab_append(vecgpu_slab* slab, const int64_t* rowids, const void* vectors, uint64_t n);

/* Hooks for Vec0Tab::insert / update / delete (src/vtab.rs:1409, 1684, 1326).
 * nbytes != vecgpu_row_bytes() marks the row as "skipped by scans", which is
 * what brute_force_search does with empty / wrong-length blobs
 * (src/vtab.rs:2596-2613).  Deleting an absent rowid is not an error. */
int vecgpu_slab_upsert(vecgpu_slab* slab, int64_t rowid, const void* vec, uint32_t nbytes);
int vecgpu_slab_delete(vecgpu_slab* slab, int64_t rowid);

/* Physically drop the skipped rows (tombstones after vecgpu_slab_delete, rows whose blob had the wrong length): SURVEY §8(f)-4,
 * the compaction that goes with Vec0Tab::delete (src/vtab.rs:1326).  Kept rows keep their order; *removed (optional) = rows
 * dropped.  Row positions change, so an HNSW index built over the slab must be rebuilt (vecgpu_hnsw_search then fails with
 * status 4 until vecgpu_hnsw_build is called again).  A dropped rowid can be re-inserted later with vecgpu_slab_upsert. */
int vecgpu_slab_compact(vecgpu_slab* s, uint64_t* removed);
/* rows = stored rows including skipped/tombstoned; live = rows scans visit. */
int vecgpu_slab_count(vecgpu_slab* slab, uint64_t* rows, uint64_t* live);
/* Copy one row back (debug / tests).  *found = 0 if absent or skipped. */
int vecgpu_slab_get(vecgpu_slab* slab, int64_t rowid, void* out_vec, int* found);

/* ---- the hot path ------------------------------------------------------ */

/* Exact KNN: replaces the body of brute_force_search (src/vtab.rs:2586-2622)
 * for nq queries at once.  Order of results per query is (distance_f32, rowid)
 * ascending == the reference's stable sort over ascending rowids
 * (src/vtab.rs:2619-2620).  out_counts[q] = min(k, live rows); result slots
 * past out_counts[q] are filled with rowid -1 / distance +inf.
 * `queries`: nq rows of vecgpu_row_bytes() bytes, host memory.
 * Errors: unsupported (type, metric) pair -> VECGPU_ERR_UNSUPPORTED exactly
 * like src/distance/mod.rs:78-82. */
int vecgpu_knn(vecgpu_slab* slab, const void* queries, uint32_t nq, uint32_t k, int metric,
               int64_t* out_rowids, float* out_dists, uint32_t* out_counts);

/* Candidate scoring: replaces the neighbour loop of search_layer
 * (src/hnsw/search.rs:501-513, and the entry-point distance :385-389) for nq
 * queries with CSR candidate lists: candidates of query q are
 * cand_rowids[cand_offsets[q] .. cand_offsets[q+1]).  out_dists has
 * cand_offsets[nq] entries; a rowid that is absent or skipped yields NaN.
 * A call of at most 256 pairs whose input fits 16 KB (one expansion: a query and its <= 64 neighbours) is a single
 * kernel launch that reads the arguments from, and stores the distances to, pinned host memory (~15 us + the caller's
 * own overhead); larger calls upload, resolve, score and download. */
int vecgpu_score(vecgpu_slab* slab, const void* queries, uint32_t nq, const int64_t* cand_rowids,
                 const uint32_t* cand_offsets, int metric, float* out_dists);

/* n independent pairs a[i] vs b[i] (host memory, row-major): replaces
 * distance::distance (src/distance/mod.rs:52-84) for the vec_distance_* SQL
 * functions (src/sql_functions.rs:153-215).  dims_a != dims_b ->
 * VECGPU_ERR_DIM_MISMATCH, type mismatch is inexpressible here by design. */
int vecgpu_distance_pairs(int elem_type, uint32_t dims_a, uint32_t dims_b, const void* a, const void* b, uint64_t n,
                          int metric, int device, float* out_dists);

/* ---- producers (src/vector.rs:444-608), on device, host in / host out ---- */

/* normalize: out = v / sqrtf(sum v^2) per row; a zero row -> VECGPU_ERR_INVALID_PARAM
 * (src/vector.rs:444-466). */
int vecgpu_normalize_f32(const float* in, uint64_t n, uint32_t dims, int device, float* out);
/* quantize_int8: per-vector min/max -> [-128,127] (src/vector.rs:514-545). */
int vecgpu_quantize_int8(const float* in, uint64_t n, uint32_t dims, int device, int8_t* out);
/* quantize_int8_for_index: clamp(v,-1,1)*127 rounded (src/vector.rs:554-575). */
int vecgpu_quantize_int8_for_index(const float* in, uint64_t n, uint32_t dims, int device, int8_t* out);
/* quantize_binary: bit = v >= mean, LSB-first (src/vector.rs:579-608). */
int vecgpu_quantize_binary(const float* in, uint64_t n, uint32_t dims, int device, uint8_t* out);

/* ---- device-resident variants (bench `value`, multi-GPU shards) ---------
 * Pointers prefixed d_ are device pointers on the slab's device; `stream` is a
 * cudaStream_t passed as void* (NULL = the CUDA default stream).  These calls
 * only enqueue work on that stream; the caller synchronises.  Calls on the same
 * slab share its scratch buffers in stream order; the streaming-scan path keeps
 * TWO scratch sets, so independent queries may be pipelined on two streams (the
 * tail of one scan then overlaps the start of the next); other paths, and a
 * third stream, are serialised against the earlier work with events. */

/* Make the slab hold n rows with dense rowids first_rowid.. generated on the
 * device by the counter-based generator value(seed, rowid, j) that
 * oracle/vecgpu_oracle.c restates, so the CPU can regenerate any row. */
int vecgpu_slab_fill_synthetic(vecgpu_slab* slab, uint64_t seed, int64_t first_rowid, uint64_t n, int kind);
/* Device addresses of the slab arrays (vectors: rows * row_stride bytes). */
int vecgpu_slab_device_view(vecgpu_slab* slab, void** d_vectors, uint32_t* row_stride, uint64_t* rows);

/* Same computation as vecgpu_knn, queries and results in device memory.
 * d_out_rowids / d_out_dists: nq*k, padded with INT64_MAX / +inf. */
int vecgpu_knn_device(vecgpu_slab* slab, const void* d_queries, uint32_t nq, uint32_t k, int metric,
                      int64_t* d_out_rowids, float* d_out_dists, void* stream);

/* k-way merge of `nlists` per-shard result lists (as produced by
 * vecgpu_knn_device and gathered from the other GPUs): for each of nq queries
 * the inputs are d_rowids/d_dists[list][q][k]; output is the global top-k in
 * (distance, rowid) order.  Replaces nothing in the reference (it is single
 * process); it is the exchange step of SURVEY §8e. */
int vecgpu_merge_device(int device, const int64_t* d_rowids, const float* d_dists, uint32_t nlists, uint32_t nq,
                        uint32_t k, int64_t* d_out_rowids, float* d_out_dists, void* stream);

/* ---- sharded slabs: one box, several GPUs ---------------------------------------------------
 * The reference is ONE process calling brute_force_search inline (src/vtab.rs:2286-2305, src/lib.rs:26-34); sharding
 * is this library's addition and has no reference line to replace.  Rows are cut by contiguous rowid range (every
 * rowid of shard i is below every rowid of shard i+1, so the global (distance, rowid) order restricted to a shard is
 * the shard's own order).  Per query: every GPU scans its range and writes its packed top-k (12 bytes per entry +
 * a count) STRAIGHT INTO THE PEERS' GATHER BUFFERS over NVLink (peer stores + a system-scope release flag), and the
 * receiving GPU merges as soon as the flags are in: no NCCL call and no pack/unpack kernels on the query path
 * (csrc/xchg.cuh).  Results are identical to one slab holding all rows.
 *
 * (1) ONE-PROCESS form — what the Rust extension would link: a sharded slab handle with the same calls as a slab. */
typedef struct vecgpu_sharded vecgpu_sharded;
/* devices NULL / n_devices 0: every visible device.  max_queries / max_k (0: 1024 / 128) size the gather buffers; a
 * larger batch is exchanged in pieces.  One host worker thread + stream per device. */
int vecgpu_sharded_create(int elem_type, uint32_t dims, uint64_t capacity_hint_total, const int* devices, uint32_t n_devices,
                          uint32_t max_queries, uint32_t max_k, vecgpu_sharded** out);
void vecgpu_sharded_destroy(vecgpu_sharded* g);
/* As vecgpu_slab_load / _fill_synthetic / _upsert / _delete / _count, over all shards (upserts and deletes are routed
 * to the shard that owns the rowid's range). */
int vecgpu_sharded_load(vecgpu_sharded* g, const int64_t* rowids, const void* vectors, uint64_t n);
int vecgpu_sharded_fill_synthetic(vecgpu_sharded* g, uint64_t seed, int64_t first_rowid, uint64_t n, int kind);
int vecgpu_sharded_upsert(vecgpu_sharded* g, int64_t rowid, const void* vec, uint32_t nbytes);
int vecgpu_sharded_delete(vecgpu_sharded* g, int64_t rowid);
int vecgpu_sharded_count(vecgpu_sharded* g, uint64_t* rows, uint64_t* live);
uint32_t vecgpu_sharded_num_shards(vecgpu_sharded* g);
/* The slab (and device) behind shard i, e.g. for vecgpu_slab_compact. */
int vecgpu_sharded_shard(vecgpu_sharded* g, uint32_t i, vecgpu_slab** slab, int* device);
/* brute_force_search (src/vtab.rs:2573-2623) over all shards: one call, host buffers, same contract as vecgpu_knn. */
int vecgpu_sharded_knn(vecgpu_sharded* g, const void* queries, uint32_t nq, uint32_t k, int metric, int64_t* out_rowids,
                       float* out_dists, uint32_t* out_counts);

/* (2) ONE-PROCESS-PER-GPU form (torchrun / MPI style launchers): each rank owns a slab with its rowid range and an
 * exchange endpoint; the endpoints are introduced to each other once (same process: _attach_local; other processes:
 * _export + _attach_ipc, the 128-byte handles travel over whatever channel the launcher has). */
typedef struct vecgpu_xchg vecgpu_xchg;
#define VECGPU_XCHG_HANDLE_BYTES 128
int vecgpu_xchg_create(int device, uint32_t rank, uint32_t world, uint32_t max_queries, uint32_t max_k, vecgpu_xchg** out);
void vecgpu_xchg_destroy(vecgpu_xchg* x);
int vecgpu_xchg_export(vecgpu_xchg* x, void* handle /* VECGPU_XCHG_HANDLE_BYTES */);
int vecgpu_xchg_attach_ipc(vecgpu_xchg* x, const void* handles /* world x VECGPU_XCHG_HANDLE_BYTES, rank order */, uint32_t n_handles);
int vecgpu_xchg_attach_local(vecgpu_xchg* const* all /* world endpoints, rank order */, uint32_t n);
/* One rank's part of a sharded query, host buffers in and out (collective: every rank calls it with the same nq, k,
 * metric, in the same order; rank r's slab must hold rowids below rank r+1's).  Every rank receives the global top-k. */
int vecgpu_shard_knn(vecgpu_slab* slab, vecgpu_xchg* x, const void* queries, uint32_t nq, uint32_t k, int metric,
                     int64_t* out_rowids, float* out_dists, uint32_t* out_counts);
/* Device-resident form: only enqueues on `stream`; outputs padded with INT64_MAX / +inf. */
int vecgpu_shard_knn_device(vecgpu_slab* slab, vecgpu_xchg* x, const void* d_queries, uint32_t nq, uint32_t k, int metric,
                            int64_t* d_out_rowids, float* d_out_dists, void* stream);
/* Exchange + merge of per-rank top-k lists that are already on the device (d_counts NULL: trailing INT64_MAX / +inf
 * entries are padding).  Collective; enqueues on `stream`. */
int vecgpu_xchg_merge_device(vecgpu_xchg* x, const int64_t* d_rowids, const float* d_dists, const uint32_t* d_counts,
                             uint32_t nq, uint32_t k, int64_t* d_out_rowids, float* d_out_dists, void* stream);
/* Synchronises `stream` and reports VECGPU_ERR_CUDA if an exchange since the last check timed out waiting for a peer. */
int vecgpu_xchg_check(vecgpu_xchg* x, void* stream);

/* Developer hook: d_buf (device, (SMs * 4 + 1) x u64) receives per-CTA globaltimer stamps of the following single-query
 * scans (CTA start, pipeline primed, rows done, CTA end, end of the fused merge); NULL switches it off.  Process-wide. */
void vecgpu_debug_scan_timeline(void* d_buf);

/* Number of kernels this library has launched in this process (bench's
 * gpu_launches claim). */
uint64_t vecgpu_launch_count(void);
/* Batched float32 queries (nq >= 16, L2 / cosine) run as a tcgen05 kind::tf32 contraction that only selects
 * candidates — ONE TF32 pass by default, with the operand-truncation error inside the certified candidate band
 * (VECGPU_TC_TERMS=3 selects the 3xTF32 compensated form) — followed by an exact re-rank; a query whose candidate
 * bound cannot be certified is re-run through the exact scan.  int8 L2 batches run exactly on kind::i8.
 * Counters since process start: queries served by the tensor-core paths / of which fell back. */
void vecgpu_tc_stats(uint64_t* queries, uint64_t* fallbacks);

/* ---- HNSW with GPU-batched candidate scoring (BASELINE config 5) -------------------------------------
 * The graph (levels, adjacency, stored edge distances) lives in HBM (the search kernel walks it there and rebuild batches
 * are linked there); a host copy is refreshed on demand (export, lockstep driver).  It sits
 * next to a slab that holds the STORED node vectors (normalised for cosine columns, int8 when
 * index_quantization=int8: src/hnsw/insert.rs:300-322); `metric` is the INTERNAL metric (src/hnsw/mod.rs:129-137).
 * Searches (queries, and the search half of every insert of a rebuild batch) run wholly on the device, one warp per
 * query walking all layers (search_layer, src/hnsw/search.rs:340-543); calls of up to 7 queries per SM (1036 on a B200) — SQL
 * issues ONE per MATCH — take a latency form instead, one CTA per query (same results; ef = 200 on 1 M x 384: ~0.6 ms
 * instead of ~4.7 ms for one query, 1.3 ms instead of 4.4 ms for 256).  With VECGPU_HNSW_DEVICE=0 in the environment
 * the lockstep driver is used instead: B inserts or queries advance together and each expansion round scores all
 * their unvisited neighbours in one launch.  Both give identical results. */
typedef struct vecgpu_hnsw vecgpu_hnsw;

/* The STORED representation of a float32 column's node vectors (src/hnsw/insert.rs:300-322): normalised when the column's
 * metric is cosine (HnswMetadata.normalize_vectors), then quantize_int8_for_index'ed with index_quantization=int8
 * (src/vector.rs:554-575).  Built on the device from the column's slab into a new slab with the same rowids and row
 * positions (element type f32 or int8); rows that cannot be stored (zero vectors under normalisation, skipped rows) are
 * skipped there too.  *out = NULL when neither step applies (the column slab is the stored representation).  Queries get
 * the same treatment (src/hnsw/search.rs:285-302) through vecgpu_normalize_f32 / vecgpu_quantize_int8_for_index; the
 * internal metric over the stored slab is L2 (cosine columns) / the column's metric, distances of cosine columns convert
 * as d^2 / 2 (src/hnsw/mod.rs:129-146).  The caller owns *out (vecgpu_slab_destroy). */
int vecgpu_hnsw_stored_slab(vecgpu_slab* column_slab, int normalize, int int8_quantization, vecgpu_slab** out);

/* M in [2,100], ef_construction in [10,2000] as vec_rebuild_hnsw validates (src/sql_functions.rs:442-469);
 * max_m0 = 2M (:489-505).  `seed` makes level assignment reproducible. */
int vecgpu_hnsw_create(vecgpu_slab* slab, int metric, uint32_t M, uint32_t ef_construction, uint64_t seed, vecgpu_hnsw** out);
void vecgpu_hnsw_destroy(vecgpu_hnsw* h);
/* vec_rebuild_hnsw (src/sql_functions.rs:436-534 -> src/hnsw/insert.rs:279-532): rebuild over every live row of
 * the slab, at most `batch` inserts per search launch (0 = 16384; never more than a quarter of the graph built so far). */
int vecgpu_hnsw_build(vecgpu_hnsw* h, uint32_t batch);
/* insert_hnsw (src/hnsw/insert.rs:279-532, called by Vec0Tab::insert when the column has an index) for rows that arrive in
 * rowid order: indexes the rows appended to the slab (vecgpu_slab_append / _upsert of a new highest rowid) since the graph
 * was last built or extended, by continuing the rebuild's insertion loop — a batch of 1 continues the strictly sequential
 * build edge for edge.  *n_inserted (may be NULL) = nodes added.  An empty index is simply built.  Rows inserted OUT of
 * rowid order move row positions: call vecgpu_hnsw_insert_at after each such upsert, or the index fails with status 4
 * (like searches) until vecgpu_hnsw_build. */
int vecgpu_hnsw_insert_appended(vecgpu_hnsw* h, uint32_t batch, uint64_t* n_inserted);
/* Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): the node of `rowid` and every edge from or to it are
 * deleted, then the row — whose vector in the slab the caller has just replaced with vecgpu_slab_upsert — is inserted
 * again (insert_hnsw).  A row that is now deleted or empty only leaves the graph — which makes this call the HNSW side of
 * Vec0Tab::delete too (src/vtab.rs:1340-1397), after vecgpu_slab_delete.  If the node was the entry point, the
 * highest remaining node takes over for the re-insertion.  The edges are removed where the lists live (one thread per
 * adjacency list): ~0.9 ms per update on a 1 M-row graph, removal and re-insertion together. */
int vecgpu_hnsw_reinsert(vecgpu_hnsw* h, int64_t rowid);
/* insert_hnsw (src/hnsw/insert.rs:279-532) for a row inserted OUT of rowid order (Vec0Tab::insert with an explicit or re-used
 * rowid, src/vtab.rs:1409-1682, insert_hnsw at :1667): vecgpu_slab_upsert has just placed `rowid` between existing rows, which moved every later
 * row one position up.  The resident graph is renumbered where it lives (one thread per adjacency list; the per-node arrays
 * shift by one row, device to device), the row gets the level of the next insertion (the level sequence follows insertion
 * order, as in the reference) and is inserted like any other row — the graph equals the sequential build's in insertion
 * order, edge for edge.  Exactly one out-of-order upsert may lie between two calls, with no un-indexed appended rows;
 * otherwise status 4 and the index needs vecgpu_hnsw_build.  (In lockstep mode, VECGPU_HNSW_DEVICE=0, the host lists are
 * renumbered instead.) */
int vecgpu_hnsw_insert_at(vecgpu_hnsw* h, int64_t rowid);
/* search_hnsw (src/hnsw/search.rs:267-335): ef = max(ef_search, k); results closest first, distances in the
 * internal metric (apply convert_distance_for_output for cosine columns); unused slots rowid -1 / +inf. */
int vecgpu_hnsw_search(vecgpu_hnsw* h, const void* queries, uint32_t nq, uint32_t k, uint32_t ef_search,
                       int64_t* out_rowids, float* out_dists, uint32_t* out_counts);
int vecgpu_hnsw_stats(vecgpu_hnsw* h, uint64_t* nodes, uint64_t* edges, int32_t* entry_level, uint64_t* distances_scored,
                      uint64_t* rounds);
/* Entry point of the graph (rowid, level) for the {t}_{c}_hnsw_meta row (HnswMetadata.entry_point_rowid / entry_point_level, src/hnsw/mod.rs:97-103); rowid -1 / level -1
 * when the index is empty. */
int vecgpu_hnsw_entry_point(vecgpu_hnsw* h, int64_t* rowid, int32_t* level);
/* Node list for a bulk write-back into {t}_{c}_hnsw_nodes(rowid, level, vector) (src/shadow.rs:464-474): the rows indexed
 * so far (rebuild + incremental inserts) that have not been deleted since, ascending rowid, with their level.  cap = 0
 * only counts. */
int vecgpu_hnsw_export_nodes(vecgpu_hnsw* h, uint64_t cap, int64_t* rowids, int32_t* levels, uint64_t* n_out);
/* Device-search counters: queries (or inserts) answered by the search kernel, how many of those hit a device capacity
 * limit and were re-run by the lockstep driver, and the number of search launches. */
int vecgpu_hnsw_device_stats(vecgpu_hnsw* h, uint64_t* queries, uint64_t* fallbacks, uint64_t* launches);
/* Expansion batch-size histogram of the device walks since the last rebuild started — the reference's BATCH_SIZE_1_4,
 * _5_16, _17_32, _33_64, _65_PLUS counters (src/hnsw/search.rs:73-85, 443-455): expansions by the number of unvisited
 * neighbours they scored. */
int vecgpu_hnsw_batch_histogram(vecgpu_hnsw* h, uint64_t out5[5]);
/* Edge list for a bulk write-back into {t}_{c}_hnsw_edges (src/shadow.rs:478-487; insert_edges_batch shape,
 * src/hnsw/storage.rs:346-383); edges from or to a row deleted since are left out, as Vec0Tab::delete removes them
 * (src/vtab.rs:1340-1397).  cap = 0 only counts. */
int vecgpu_hnsw_export_edges(vecgpu_hnsw* h, uint64_t cap, int64_t* from_rowids, int64_t* to_rowids, int32_t* levels,
                             float* dists, uint64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* VECGPU_H */
