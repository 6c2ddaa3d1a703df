/*
 * hnsw_seq.c — CPU ORACLE, TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A literal, strictly SEQUENTIAL restatement of the reference's HNSW build and query:
 *   insert_hnsw    src/hnsw/insert.rs:279-532   (one insert at a time, every insert sees all earlier ones)
 *   prune          src/hnsw/insert.rs:144-222   (prune_neighbor_if_needed: keep the closest max_connections by STORED
 *                                                distance; would_survive_prune is `true`, :224-243)
 *   search_layer   src/hnsw/search.rs:340-543
 *   search_hnsw    src/hnsw/search.rs:267-335
 * over an in-memory copy of the two shadow tables (nodes: level + vector; edges: PK (from, level, to) + distance,
 * src/shadow.rs:464-487).  Distances come from vecgpu_oracle.c (orc_distance), so the arithmetic is the scan's.
 *
 * It exists for two things the GPU build cannot show about itself:
 *   1. the product inserts in BATCHES whose members do not see each other; this file is the batch-size-1 truth the
 *      product's graph is compared with edge for edge (tests/test_gpu_hnsw_seq.py), and
 *   2. recall of the reference's OWN procedure at BASELINE cfg5 (1 M x 384, M=16, efc=200), measured here on the CPU
 *      (tools/hnsw_recall_study.py), so that a low recall can be attributed to the algorithm or to the batching.
 *
 * Choices the reference leaves open, fixed here:
 *   - levels are an INPUT (the reference draws them from a time-seeded hash, insert.rs:114-137; the product uses a
 *     counter-based hash so that builds are reproducible — the caller passes that sequence);
 *   - neighbours are visited in ascending neighbour rowid: the reference's "SELECT to_rowid ... WHERE from_rowid = ? AND
 *     level = ?" (src/hnsw/storage.rs:163) has no ORDER BY and walks the primary key (from, level, to);
 *   - equal distances: the reference's heaps compare distances only (search.rs:212-250) and leave ties to BinaryHeap
 *     internals; here ties order by node id.  Irrelevant for continuous data.
 *   - `quirk` != 0 keeps one oddity of the reference: an insert whose level is ABOVE the current entry level still runs
 *     search_layer on those upper layers, finds only the old entry point there and links to it (insert.rs:408-470 has no
 *     guard), so the old entry point gets edges on layers above its own level.  The product does not create those
 *     edges; quirk = 0 leaves them out so the two graphs can be compared.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int orc_distance(int elem, uint32_t dims_a, uint32_t dims_b, const void* a, const void* b, int metric, float* out);
uint32_t orc_row_bytes(int elem, uint32_t dims);

#define MAX_LEVELS 16

typedef struct {
    uint32_t* nb;  /* ascending neighbour id */
    float* d;      /* stored edge distance */
    uint32_t deg, cap;
} adj_t;

typedef struct {
    int elem, metric;
    uint32_t dims, rb, M, max_m0, efc;
    int quirk;
    uint64_t n;
    const uint8_t* vec;
    int8_t* level;
    adj_t* l0;         /* [n] level 0 */
    adj_t** up;        /* [n] -> [MAX_LEVELS] upper levels, allocated on first use */
    int64_t entry;
    int entry_level;
    uint64_t n_nodes, n_dist;
    /* search scratch */
    uint32_t* seen;    /* [n] generation stamps */
    uint32_t gen;
    /* expansion batch-size histogram: 1-4, 5-16, 17-32, 33-64, 65+ (search.rs:443-455) */
    uint64_t hist[5], fetches;
} hnsw_t;

typedef struct {
    float d;
    uint32_t node;
} cand_t;

static adj_t* adj_of(hnsw_t* h, uint32_t node, int lv, int create) {
    if (lv == 0) return &h->l0[node];
    if (!h->up[node]) {
        if (!create) return NULL;
        h->up[node] = calloc(MAX_LEVELS, sizeof(adj_t));
    }
    return &h->up[node][lv];
}

static float dist_to(hnsw_t* h, const void* q, uint32_t node) {
    float d = 0;
    orc_distance(h->elem, h->dims, h->dims, q, h->vec + (size_t)node * h->rb, h->metric, &d);
    ++h->n_dist;
    return d;
}

/* INSERT OR IGNORE INTO edges (from, to, level, distance) — storage.rs:346-383; kept in primary-key order */
static void edge_insert(hnsw_t* h, uint32_t from, uint32_t to, int lv, float d) {
    adj_t* a = adj_of(h, from, lv, 1);
    uint32_t lo = 0, hi = a->deg;
    while (lo < hi) {
        uint32_t mid = (lo + hi) / 2;
        if (a->nb[mid] < to) lo = mid + 1;
        else hi = mid;
    }
    if (lo < a->deg && a->nb[lo] == to) return; /* OR IGNORE */
    if (a->deg == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->nb = realloc(a->nb, a->cap * sizeof(uint32_t));
        a->d = realloc(a->d, a->cap * sizeof(float));
    }
    memmove(a->nb + lo + 1, a->nb + lo, (a->deg - lo) * sizeof(uint32_t));
    memmove(a->d + lo + 1, a->d + lo, (a->deg - lo) * sizeof(float));
    a->nb[lo] = to;
    a->d[lo] = d;
    ++a->deg;
}

/* prune_neighbor_if_needed — insert.rs:144-222: fetch edges with distances (PK order), early return if <= max, stable sort
 * by stored distance, delete everything past max_connections */
static void prune(hnsw_t* h, uint32_t node, int lv, uint32_t maxc) {
    adj_t* a = adj_of(h, node, lv, 0);
    if (!a || a->deg <= maxc) return;
    const uint32_t n = a->deg;
    uint32_t* idx = malloc(n * sizeof(uint32_t));
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    for (uint32_t i = 1; i < n; ++i) { /* insertion sort == stable */
        uint32_t x = idx[i];
        uint32_t j = i;
        while (j > 0 && a->d[idx[j - 1]] > a->d[x]) {
            idx[j] = idx[j - 1];
            --j;
        }
        idx[j] = x;
    }
    uint8_t* keep = calloc(n, 1);
    for (uint32_t i = 0; i < maxc; ++i) keep[idx[i]] = 1;
    uint32_t w = 0;
    for (uint32_t i = 0; i < n; ++i)
        if (keep[i]) {
            a->nb[w] = a->nb[i];
            a->d[w] = a->d[i];
            ++w;
        }
    a->deg = w;
    free(idx);
    free(keep);
}

/* binary heaps on (d, node) */
static int lt(cand_t a, cand_t b) { return a.d < b.d || (a.d == b.d && a.node < b.node); }
static void heap_push(cand_t* h, uint32_t* n, cand_t x, int maxheap) {
    uint32_t i = (*n)++;
    while (i > 0) {
        uint32_t p = (i - 1) / 2;
        if (maxheap ? !lt(h[p], x) : !lt(x, h[p])) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = x;
}
static cand_t heap_pop(cand_t* h, uint32_t* n, int maxheap) {
    cand_t top = h[0], x = h[--(*n)];
    uint32_t i = 0;
    for (;;) {
        uint32_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && (maxheap ? lt(h[c], h[c + 1]) : lt(h[c + 1], h[c]))) ++c;
        if (maxheap ? !lt(x, h[c]) : !lt(h[c], x)) break;
        h[i] = h[c];
        i = c;
    }
    h[i] = x;
    return top;
}

static int cmp_cand(const void* a, const void* b) {
    const cand_t *x = a, *y = b;
    return lt(*x, *y) ? -1 : (lt(*y, *x) ? 1 : 0);
}

/* search_layer — search.rs:340-543.  out must hold ef entries; returns the count, sorted closest first. */
static uint32_t search_layer(hnsw_t* h, const void* q, uint32_t entry, uint32_t ef, int lv, cand_t* out, cand_t* cand, cand_t* res) {
    uint32_t nc = 0, nr = 0;
    if (++h->gen == 0) {
        memset(h->seen, 0, h->n * sizeof(uint32_t));
        h->gen = 1;
    }
    const cand_t e = {dist_to(h, q, entry), entry}; /* :385-389 */
    heap_push(cand, &nc, e, 0);
    heap_push(res, &nr, e, 1);
    h->seen[entry] = h->gen;
    while (nc) {
        const cand_t c = heap_pop(cand, &nc, 0);
        if (nr && c.d > res[0].d) break; /* :406-410 */
        const adj_t* a = adj_of(h, c.node, lv, 0);
        if (!a) continue;
        uint32_t fresh = 0;
        for (uint32_t i = 0; i < a->deg; ++i)
            if (h->seen[a->nb[i]] != h->gen) ++fresh;
        if (!fresh) continue;
        ++h->fetches;
        h->hist[fresh <= 4 ? 0 : fresh <= 16 ? 1 : fresh <= 32 ? 2 : fresh <= 64 ? 3 : 4]++;
        /* :424-434 marks all unvisited neighbours visited first, then scores them in that order (:501-532) */
        uint32_t* todo = malloc(fresh * sizeof(uint32_t));
        uint32_t nt = 0;
        for (uint32_t i = 0; i < a->deg; ++i)
            if (h->seen[a->nb[i]] != h->gen) {
                h->seen[a->nb[i]] = h->gen;
                todo[nt++] = a->nb[i];
            }
        for (uint32_t i = 0; i < nt; ++i) {
            const cand_t x = {dist_to(h, q, todo[i]), todo[i]};
            if (nr < ef || x.d < res[0].d) { /* :516 */
                heap_push(cand, &nc, x, 0);
                heap_push(res, &nr, x, 1);
                while (nr > ef) heap_pop(res, &nr, 1); /* :528-531 */
            }
        }
        free(todo);
    }
    memcpy(out, res, nr * sizeof(cand_t));
    qsort(out, nr, sizeof(cand_t), cmp_cand); /* :540 */
    return nr;
}

hnsw_t* orc_hnsw_new(int elem, uint32_t dims, int metric, const void* vectors, uint64_t n, uint32_t M, uint32_t efc, int quirk) {
    hnsw_t* h = calloc(1, sizeof(hnsw_t));
    h->elem = elem;
    h->dims = dims;
    h->metric = metric;
    h->rb = orc_row_bytes(elem, dims);
    h->M = M;
    h->max_m0 = 2 * M; /* sql_functions.rs:489-505 */
    h->efc = efc;
    h->quirk = quirk;
    h->n = n;
    h->vec = vectors;
    h->level = malloc(n ? n : 1);
    memset(h->level, 0xFF, n ? n : 1); /* -1: not a node (yet) */
    h->l0 = calloc(n ? n : 1, sizeof(adj_t));
    h->up = calloc(n ? n : 1, sizeof(adj_t*));
    h->seen = calloc(n ? n : 1, sizeof(uint32_t));
    h->entry = -1;
    h->entry_level = -1;
    return h;
}

void orc_hnsw_free(hnsw_t* h) {
    if (!h) return;
    for (uint64_t i = 0; i < h->n; ++i) {
        free(h->l0[i].nb);
        free(h->l0[i].d);
        if (h->up[i]) {
            for (int l = 0; l < MAX_LEVELS; ++l) {
                free(h->up[i][l].nb);
                free(h->up[i][l].d);
            }
            free(h->up[i]);
        }
    }
    free(h->level);
    free(h->l0);
    free(h->up);
    free(h->seen);
    free(h);
}

/* insert_hnsw — insert.rs:279-532 for node `node` (position in `vectors`) at level `level` */
void orc_hnsw_insert(hnsw_t* h, uint32_t node, int level) {
    h->level[node] = (int8_t)level;
    if (h->entry < 0) { /* :336-350 */
        h->entry = node;
        h->entry_level = level;
        h->n_nodes = 1;
        return;
    }
    const void* q = h->vec + (size_t)node * h->rb;
    const uint32_t cap = (h->efc > 1 ? h->efc : 1) + 2;
    cand_t* out = malloc(cap * sizeof(cand_t));
    cand_t* res = malloc((cap + 1) * sizeof(cand_t));
    /* the candidate heap can hold everything ever admitted in one layer search */
    cand_t* cand = malloc((h->n_nodes + 2) * sizeof(cand_t));
    uint32_t cur = (uint32_t)h->entry;
    for (int lv = h->entry_level; lv > level; --lv) { /* :396-405, ef = 1 */
        uint32_t c = search_layer(h, q, cur, 1, lv, out, cand, res);
        if (c) cur = out[0].node;
    }
    for (int lv = level; lv >= 0; --lv) { /* :408-498 */
        if (lv > h->entry_level && !h->quirk) continue;
        uint32_t c = search_layer(h, q, cur, h->efc, lv, out, cand, res);
        const uint32_t maxc = lv == 0 ? h->max_m0 : h->M;
        const uint32_t take = c < maxc ? c : maxc; /* :421-430; would_survive_prune == true */
        for (uint32_t i = 0; i < take; ++i) {      /* :463-470 */
            edge_insert(h, node, out[i].node, lv, out[i].d);
            edge_insert(h, out[i].node, node, lv, out[i].d);
        }
        for (uint32_t i = 0; i < take; ++i) prune(h, out[i].node, lv, maxc); /* :476-493 */
        if (take) cur = out[0].node;                                           /* :496-498 */
    }
    if (level > h->entry_level) { /* :502-506 */
        h->entry = node;
        h->entry_level = level;
    }
    ++h->n_nodes;
    free(out);
    free(res);
    free(cand);
}

/* Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): DELETE the node row and every edge from or to it, then
 * insert_hnsw the row again (its vector — read through h->vec — has been replaced by the caller).  The level is the caller's
 * (the product keeps the level of the row position).  When the node was the entry point the search for the re-insertion
 * starts from the highest remaining node (first such position), which is what the meta row would be repaired to. */
void orc_hnsw_reinsert(hnsw_t* h, uint32_t node, int level, int insert_again) {
    for (uint64_t v = 0; v < h->n; ++v) {
        for (int lv = 0; lv < MAX_LEVELS; ++lv) {
            adj_t* a = adj_of(h, (uint32_t)v, lv, 0);
            if (!a) break;
            if (v == node) {
                a->deg = 0;
                continue;
            }
            uint32_t w = 0;
            for (uint32_t i = 0; i < a->deg; ++i)
                if (a->nb[i] != node) {
                    a->nb[w] = a->nb[i];
                    a->d[w] = a->d[i];
                    ++w;
                }
            a->deg = w;
        }
    }
    h->level[node] = -1; /* not a node any more */
    --h->n_nodes;
    if (h->entry == (int64_t)node) {
        h->entry = -1;
        h->entry_level = -1;
        for (uint64_t v = 0; v < h->n; ++v)
            if (h->level[v] > h->entry_level) {
                h->entry = (int64_t)v;
                h->entry_level = h->level[v];
            }
    }
    if (insert_again) orc_hnsw_insert(h, node, level);
}

/* vec_rebuild_hnsw shape: insert rows 0..n-1 in order with the given levels (skip[i] != 0: row not indexed) */
void orc_hnsw_build(hnsw_t* h, const int8_t* levels, const uint8_t* skip) {
    for (uint64_t i = 0; i < h->n; ++i)
        if (!skip || !skip[i]) orc_hnsw_insert(h, (uint32_t)i, levels[i]);
}

/* search_hnsw — search.rs:267-335.  Returns the number of results (<= k). */
uint32_t orc_hnsw_search(hnsw_t* h, const void* query, uint32_t k, uint32_t ef_search, uint32_t* out_nodes, float* out_dists) {
    if (h->entry < 0 || k == 0) return 0;
    const uint32_t ef = ef_search > k ? ef_search : k; /* :282 */
    cand_t* out = malloc((ef + 2) * sizeof(cand_t));
    cand_t* res = malloc((ef + 3) * sizeof(cand_t));
    cand_t* cand = malloc((h->n_nodes + 2) * sizeof(cand_t));
    uint32_t cur = (uint32_t)h->entry;
    for (int lv = h->entry_level; lv >= 1; --lv) { /* :300-310 */
        uint32_t c = search_layer(h, query, cur, 1, lv, out, cand, res);
        if (c) cur = out[0].node;
    }
    uint32_t c = search_layer(h, query, cur, ef, 0, out, cand, res); /* :313 */
    if (c > k) c = k;
    for (uint32_t i = 0; i < c; ++i) {
        out_nodes[i] = out[i].node;
        out_dists[i] = out[i].d;
    }
    free(out);
    free(res);
    free(cand);
    return c;
}

/* graph inspection: edges of (node, level) in primary-key order */
uint32_t orc_hnsw_degree(hnsw_t* h, uint32_t node, int lv) {
    adj_t* a = adj_of(h, node, lv, 0);
    return a ? a->deg : 0;
}
uint32_t orc_hnsw_edges(hnsw_t* h, uint32_t node, int lv, uint32_t* nb, float* d) {
    adj_t* a = adj_of(h, node, lv, 0);
    if (!a) return 0;
    memcpy(nb, a->nb, a->deg * sizeof(uint32_t));
    memcpy(d, a->d, a->deg * sizeof(float));
    return a->deg;
}
/* total edges; every edge as (from, to, level, distance) when the arrays are given */
uint64_t orc_hnsw_export(hnsw_t* h, uint32_t* from, uint32_t* to, int32_t* lvl, float* dist) {
    uint64_t e = 0;
    for (uint64_t i = 0; i < h->n; ++i)
        for (int lv = 0; lv < MAX_LEVELS; ++lv) {
            adj_t* a = adj_of(h, (uint32_t)i, lv, 0);
            if (!a) break;
            for (uint32_t j = 0; j < a->deg; ++j, ++e)
                if (from) {
                    from[e] = (uint32_t)i;
                    to[e] = a->nb[j];
                    lvl[e] = lv;
                    dist[e] = a->d[j];
                }
        }
    return e;
}
void orc_hnsw_info(hnsw_t* h, int64_t* entry, int32_t* entry_level, uint64_t* nodes, uint64_t* distances, uint64_t* hist5, uint64_t* fetches) {
    if (entry) *entry = h->entry;
    if (entry_level) *entry_level = h->entry_level;
    if (nodes) *nodes = h->n_nodes;
    if (distances) *distances = h->n_dist;
    if (hist5) memcpy(hist5, h->hist, sizeof(h->hist));
    if (fetches) *fetches = h->fetches;
}

/* the product's reproducible level sequence (csrc/hnsw.inl hnsw_level_for): floor(-ln(u) / ln(M)) with u from a
 * counter-based hash of (seed, position); formula of insert.rs:129-136, max_level 16 (hnsw/mod.rs:35-47) */
static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_hnsw_levels(uint64_t seed, uint64_t n, uint32_t M, int8_t* out) {
    const double lf = 1.0 / log((double)M);
    for (uint64_t pos = 0; pos < n; ++pos) {
        const uint64_t r = mix64(seed * 0x9E3779B97F4A7C15ull + pos + 1);
        double u = (double)(r % 1000000ull) / 1000000.0;
        if (u < 1e-9) u = 1e-9;
        int level = (int)floor(-log(u) * lf);
        out[pos] = (int8_t)(level < 0 ? 0 : (level > MAX_LEVELS - 1 ? MAX_LEVELS - 1 : level));
    }
}
