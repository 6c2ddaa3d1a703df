// hnsw.inl — host side of the HNSW build / search (BASELINE config 5); the device code is in hnsw_dev.cuh.
// Included by vecgpu.cu (needs vecgpu_slab, launch_pairs, ws_reserve, pin_reserve, CU, LAUNCHED, fail).
//
// What it replaces in the reference: insert_hnsw (src/hnsw/insert.rs:279-532), search_hnsw / search_layer
// (src/hnsw/search.rs:267-543) and the SQL node/edge fetches behind them (src/hnsw/storage.rs), for the
// rebuild (src/sql_functions.rs:436-534) and for queries.  The reference walks the graph one query at a
// time and scores 1-32 neighbours per expansion through SQLite lookups.  Here:
//   - default: the adjacency lists live in HBM; hnsw_search_kernel walks all layers of a query (or of an
//     insert of a rebuild batch) with one warp, and a rebuild batch is linked on the device as well
//     (hnsw_link_*_kernel); the host copy of the lists is refreshed on demand (export, lockstep driver);
//   - lockstep driver (VECGPU_HNSW_DEVICE=0, and the path that answers a query which overflowed a device
//     capacity): B inserts / queries advance together and every round scores ALL their unvisited neighbours
//     in one launch of pair_kernel; heaps, visited sets and linking on the host.
// Both produce the same graph and the same results; the graph can be exported for a bulk write-back into
// {t}_{c}_hnsw_edges (src/shadow.rs:478-487).
//
// Semantics kept from the reference:
//   - layer search: entry scored first (search.rs:385-398); pop closest candidate, stop when it is farther
//     than the worst result (strict >, :406-410); neighbours filtered by a visited set before scoring
//     (:424-434); admission `len < ef || d < worst` (strict <, :516); results trimmed to ef (:528-531).
//   - insert: greedy descent with ef=1 above the node's level (insert.rs:396-405), ef_construction search
//     per level, the closest max_connections results become neighbours (max_m0 = 2M at level 0, M above;
//     :421-430), bidirectional edges carrying the distance (:463-470), overfull neighbours pruned by
//     keeping the closest by STORED distance (simple_prune, :144-222), entry point raised when the new
//     node's level exceeds it.
//   - levels: floor(-ln(u) * 1/ln(M)), capped at max_level-1 = 15 (insert.rs:114-137, hnsw/mod.rs:35-47);
//     u comes from a counter-based hash of (seed, position), so builds are reproducible (the reference's
//     are not: SURVEY F7).
// Deviation: nodes of one batch do not see each other while searching (they are linked afterwards, in
// order); batches start small and grow with the graph (at most a quarter of it).  Judged by recall.

struct HCand {
    float d;
    uint32_t node;
};
struct HMinCmp {  // std heap with this comparator = min-heap on (d, node)
    bool operator()(const HCand& a, const HCand& b) const { return a.d > b.d || (a.d == b.d && a.node > b.node); }
};
struct HMaxCmp {  // max-heap on (d, node)
    bool operator()(const HCand& a, const HCand& b) const { return a.d < b.d || (a.d == b.d && a.node < b.node); }
};

struct HVisited {  // open addressing, cleared by replaying the inserted slots
    std::vector<uint32_t> tab;
    std::vector<uint32_t> used;
    uint32_t mask = 0;
    void init(uint32_t cap_pow2) {
        tab.assign(cap_pow2, 0xFFFFFFFFu);
        mask = cap_pow2 - 1;
        used.clear();
    }
    void clear() {
        for (uint32_t s : used) tab[s] = 0xFFFFFFFFu;
        used.clear();
    }
    void grow() {
        std::vector<uint32_t> old;
        old.reserve(used.size());
        for (uint32_t s : used) old.push_back(tab[s]);
        init((mask + 1) * 2);
        for (uint32_t k : old) insert(k);
    }
    bool insert(uint32_t key) {  // true if newly inserted
        if (used.size() * 2 > mask) grow();
        uint32_t h = (key * 2654435761u) & mask;
        while (true) {
            const uint32_t v = tab[h];
            if (v == key) return false;
            if (v == 0xFFFFFFFFu) {
                tab[h] = key;
                used.push_back(h);
                return true;
            }
            h = (h + 1) & mask;
        }
    }
};

struct HQuery {
    uint32_t a_index = 0;       // row of the query in the a-operand buffer (slab position for inserts)
    int level = 0;              // layer being searched
    int stop_level = 0;         // last layer to search (0)
    int node_level = -1;        // insert: the new node's level; search: -1 (only layer 0 collects ef results)
    uint32_t ef = 1, ef_wide = 1;  // ef of the current layer; ef of the layers that collect results
    uint32_t entry = 0;
    bool started = false, done = false;
    std::vector<HCand> cand, res;
    HVisited visited;
    std::vector<uint32_t> pending;            // nodes submitted for scoring this round
    std::vector<std::vector<HCand>> layers;   // sorted results of the layers <= node_level (insert) / layer 0 (search)
};

struct vecgpu_hnsw {
    vecgpu_slab* slab = nullptr;
    int metric = VECGPU_L2;
    uint32_t M = 16, max_m0 = 32, efc = 200;
    uint64_t seed = 0;
    int max_level = 16;
    double level_factor = 1.0 / std::log(16.0);
    uint64_t n_nodes = 0;
    int64_t entry = -1;
    int entry_level = -1;
    std::vector<int8_t> node_level;
    std::vector<uint8_t> in_graph;  // rows that were inserted by the last rebuild (rows skipped at that time are not nodes)
    // level 0 adjacency: [node][max_m0]; upper levels: node with level L owns L consecutive M-wide lists
    std::vector<uint32_t> nbr0;
    std::vector<float> dist0;
    std::vector<uint16_t> deg0;
    std::vector<uint32_t> upper_base;
    std::vector<uint32_t> nbrU;
    std::vector<float> distU;
    std::vector<uint16_t> degU;
    uint64_t scored = 0, rounds = 0;
    uint64_t batch_hist[5] = {0, 0, 0, 0, 0};  // device walks: expansions by unvisited-neighbour count (search.rs:443-455 buckets)
    uint64_t slab_gen = 0;  // layout generation of the slab the graph was built over (nodes are row positions)
    std::mutex mu;
    // pinned staging + device buffers for the per-round pair lists
    void* h_pin = nullptr;
    size_t pin_cap = 0;
    void* d_buf = nullptr;
    size_t d_cap = 0;
    void* d_queries = nullptr;
    size_t dq_cap = 0;
    // K6: adjacency resident in HBM for the one-warp-per-query search kernel (hnsw_dev.cuh)
    uint32_t* d_nbr0 = nullptr;
    uint16_t* d_deg0 = nullptr;
    uint32_t* d_upper_base = nullptr;
    uint32_t* d_nbrU = nullptr;
    uint16_t* d_degU = nullptr;
    size_t dn_rows = 0, dn_slots = 0;
    size_t dcap_rows = 0, dcap_slots = 0;  // allocated capacity of the device lists (>= dn_rows / dn_slots: incremental inserts grow them)
    bool dev_valid = false;              // device copy exists and matches the host lists except for the dirty ones
    std::vector<uint8_t> dirty0, dirtyU;
    std::vector<uint32_t> dirty0_list, dirtyU_list;
    uint32_t* d_visited = nullptr;
    size_t vis_cap = 0;
    uint8_t* d_sw = nullptr;             // per-launch arrays of the search kernel
    size_t sw_cap = 0;
    uint8_t* h_sw = nullptr;             // pinned twin
    size_t hsw_cap = 0;
    uint64_t dev_queries = 0, dev_fallbacks = 0, dev_launches = 0;
    // device-side linking: stored edge distances on the device, scratch for the operation sort; while a rebuild links on
    // the device the host lists are stale until hnsw_ensure_host() downloads them
    float* d_dist0 = nullptr;
    float* d_distU = nullptr;
    bool host_stale = false;
    void* d_spare[4] = {nullptr, nullptr, nullptr, nullptr};  // vecgpu_hnsw_insert_at: second set of the level-0 arrays (nbr0, dist0, deg0, upper_base)
    size_t spare_rows = 0;
    uint8_t* d_link = nullptr;
    size_t link_cap = 0;
};

static inline uint64_t h_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int hnsw_level_for(const vecgpu_hnsw* h, uint64_t pos) {
    const uint64_t r = h_mix64(h->seed * 0x9E3779B97F4A7C15ull + pos + 1);
    double u = (double)(r % 1000000ull) / 1000000.0;  // [0,1), insert.rs:129
    if (u < 1e-9) u = 1e-9;
    int level = (int)std::floor(-std::log(u) * h->level_factor);
    return std::max(0, std::min(level, h->max_level - 1));
}

static inline uint32_t* h_nbr(vecgpu_hnsw* h, uint32_t node, int level, float** dist, uint16_t** deg, uint32_t* maxc) {
    if (level == 0) {
        *dist = &h->dist0[(size_t)node * h->max_m0];
        *deg = &h->deg0[node];
        *maxc = h->max_m0;
        return &h->nbr0[(size_t)node * h->max_m0];
    }
    const size_t slot = (size_t)h->upper_base[node] + (size_t)(level - 1);
    *dist = &h->distU[slot * h->M];
    *deg = &h->degU[slot];
    *maxc = h->M;
    return &h->nbrU[slot * h->M];
}

// add edge from -> to with stored distance; an overfull list keeps the closest maxc by (distance, node).
// `dirty` (optional) collects the lists that changed, for the incremental upload of the device copy.
static void h_add_edge(vecgpu_hnsw* h, uint32_t from, uint32_t to, float d, int level, std::vector<uint32_t>* dirty0 = nullptr,
                       std::vector<uint32_t>* dirtyU = nullptr) {
    float* dist;
    uint16_t* deg;
    uint32_t maxc;
    uint32_t* nb = h_nbr(h, from, level, &dist, &deg, &maxc);
    if (dirty0) {
        if (level == 0) {
            if (!h->dirty0[from]) {
                h->dirty0[from] = 1;
                dirty0->push_back(from);
            }
        } else {
            const uint32_t slot = h->upper_base[from] + (uint32_t)(level - 1);
            if (!h->dirtyU[slot]) {
                h->dirtyU[slot] = 1;
                dirtyU->push_back(slot);
            }
        }
    }
    for (uint32_t i = 0; i < *deg; ++i)
        if (nb[i] == to) {
            dist[i] = d;
            return;
        }
    if (*deg < maxc) {
        nb[*deg] = to;
        dist[*deg] = d;
        ++*deg;
        return;
    }
    uint32_t worst = 0;
    for (uint32_t i = 1; i < maxc; ++i)
        if (dist[i] > dist[worst] || (dist[i] == dist[worst] && nb[i] > nb[worst])) worst = i;
    if (d < dist[worst] || (d == dist[worst] && to < nb[worst])) {
        nb[worst] = to;
        dist[worst] = d;
    }
}

// ---- one GPU launch for all (query, node) pairs of a round ---------------------------------------------
static int hnsw_score_round(vecgpu_hnsw* h, const uint8_t* a_base, const std::vector<uint32_t>& a_idx,
                            const std::vector<uint32_t>& b_pos, std::vector<float>& out) {
    vecgpu_slab* s = h->slab;
    const size_t n = a_idx.size();
    out.resize(n);
    if (n == 0) return 0;
    const size_t up = n * 4 + n * 8;
    if (up > h->pin_cap) {
        if (h->h_pin) CU(cudaFreeHost(h->h_pin));
        h->h_pin = nullptr;
        h->pin_cap = std::max(up * 2, (size_t)1 << 20);
        CU(cudaMallocHost(&h->h_pin, h->pin_cap));
    }
    const size_t dbytes = n * 16;
    if (dbytes > h->d_cap) {
        if (h->d_buf) CU(cudaFree(h->d_buf));
        h->d_buf = nullptr;
        h->d_cap = std::max(dbytes * 2, (size_t)1 << 20);
        CU(cudaMalloc(&h->d_buf, h->d_cap));
    }
    uint8_t* hp = (uint8_t*)h->h_pin;
    int64_t* hb = (int64_t*)hp;
    uint32_t* ha = (uint32_t*)(hp + n * 8);
    for (size_t i = 0; i < n; ++i) {
        hb[i] = (int64_t)b_pos[i];
        ha[i] = a_idx[i];
    }
    uint8_t* db = (uint8_t*)h->d_buf;
    CU(cudaMemcpyAsync(db, hp, up, cudaMemcpyHostToDevice, s->stream));
    PairParams p{};
    p.a_base = a_base;
    p.b_base = s->d_vec;
    p.a_stride = p.b_stride = s->row_stride;
    p.units = s->row_stride / 16;
    p.b_index = (const int64_t*)db;
    p.a_index = (const uint32_t*)(db + n * 8);
    p.n_pairs = n;
    p.out = (float*)(db + n * 12);
    p.qc_kind = s->elem == VECGPU_I8 ? 1u : 0u;
    int rc = launch_pairs(s->elem, h->metric, p, s->num_sms, s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(hp, p.out, n * 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(out.data(), hp, n * 4);
    h->scored += n;
    h->rounds += 1;
    return 0;
}

// start the search of `q.level` from q.entry
static void hq_start_layer(HQuery& q) {
    q.cand.clear();
    q.res.clear();
    q.visited.clear();
    q.started = false;
}

// finish the current layer: record / descend.  Returns with q.done set when nothing is left.
static void hq_finish_layer(vecgpu_hnsw* h, HQuery& q) {
    std::vector<HCand> sorted = q.res;
    std::sort(sorted.begin(), sorted.end(), [](const HCand& a, const HCand& b) { return a.d < b.d || (a.d == b.d && a.node < b.node); });
    const bool collect = q.node_level < 0 ? q.level == 0 : q.level <= q.node_level;
    if (!sorted.empty()) q.entry = sorted[0].node;  // closest becomes the entry of the next layer (insert.rs / search.rs:318-323)
    if (collect) {
        if ((int)q.layers.size() <= q.level) q.layers.resize(q.level + 1);
        q.layers[q.level] = std::move(sorted);
    }
    if (q.level == q.stop_level) {
        q.done = true;
        return;
    }
    q.level -= 1;
    const bool wide = q.node_level < 0 ? q.level == 0 : q.level <= q.node_level;
    q.ef = wide ? q.ef_wide : 1;
    hq_start_layer(q);
}

// advance all queries until every one is done; a_base = device rows the queries live in
static int hnsw_run_batch(vecgpu_hnsw* h, std::vector<HQuery>& qs, size_t nqs, const uint8_t* a_base) {
    std::vector<uint32_t> a_idx, b_pos, offs(nqs + 1);
    std::vector<float> dists;
    while (true) {
        // ---- prepare: each query pops until it has neighbours to score (or finishes)
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t qi = 0; qi < (int64_t)nqs; ++qi) {
            HQuery& q = qs[qi];
            q.pending.clear();
            while (!q.done && q.pending.empty()) {
                if (!q.started) {
                    q.pending.push_back(q.entry);  // the entry point is scored first
                    q.visited.insert(q.entry);
                    break;
                }
                if (q.cand.empty()) {
                    hq_finish_layer(h, q);
                    continue;
                }
                std::pop_heap(q.cand.begin(), q.cand.end(), HMinCmp());
                const HCand c = q.cand.back();
                q.cand.pop_back();
                if (!q.res.empty() && c.d > q.res.front().d) {  // farther than the worst result: layer done
                    hq_finish_layer(h, q);
                    continue;
                }
                float* nd;
                uint16_t* deg;
                uint32_t maxc;
                const uint32_t* nb = h_nbr(h, c.node, q.level, &nd, &deg, &maxc);
                for (uint32_t i = 0; i < *deg; ++i)
                    if (q.visited.insert(nb[i])) q.pending.push_back(nb[i]);
            }
        }
        // ---- gather
        size_t total = 0;
        for (size_t qi = 0; qi < nqs; ++qi) {
            offs[qi] = (uint32_t)total;
            total += qs[qi].pending.size();
        }
        offs[nqs] = (uint32_t)total;
        if (total == 0) break;
        a_idx.resize(total);
        b_pos.resize(total);
#pragma omp parallel for schedule(static)
        for (int64_t qi = 0; qi < (int64_t)nqs; ++qi) {
            const HQuery& q = qs[qi];
            for (size_t j = 0; j < q.pending.size(); ++j) {
                a_idx[offs[qi] + j] = q.a_index;
                b_pos[offs[qi] + j] = q.pending[j];
            }
        }
        // ---- one launch scores every pending pair of the batch
        int rc = hnsw_score_round(h, a_base, a_idx, b_pos, dists);
        if (rc) return rc;
        // ---- update heaps (admission rule of search.rs:516-531)
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t qi = 0; qi < (int64_t)nqs; ++qi) {
            HQuery& q = qs[qi];
            for (size_t j = 0; j < q.pending.size(); ++j) {
                const HCand c{dists[offs[qi] + j], q.pending[j]};
                if (c.d != c.d) continue;  // NaN: node missing
                if (!q.started || q.res.size() < q.ef || c.d < q.res.front().d) {
                    q.cand.push_back(c);
                    std::push_heap(q.cand.begin(), q.cand.end(), HMinCmp());
                    q.res.push_back(c);
                    std::push_heap(q.res.begin(), q.res.end(), HMaxCmp());
                    while (q.res.size() > q.ef) {
                        std::pop_heap(q.res.begin(), q.res.end(), HMaxCmp());
                        q.res.pop_back();
                    }
                }
                q.started = true;
            }
            if (!q.pending.empty()) q.started = true;
        }
    }
    return 0;
}


// ---- K6: device-resident graph + one-warp-per-query search (hnsw_dev.cuh) ------------------------------------
static bool hnsw_device_enabled(const vecgpu_hnsw* h) {
    const char* e = getenv("VECGPU_HNSW_DEVICE");
    if (e && e[0] == '0') return false;
    return h->node_level.size() < 0x7FFFFFFFull;  // node << 1 must fit the low key word
}

static void hnsw_dev_free_graph(vecgpu_hnsw* h) {
    cudaFree(h->d_nbr0);
    cudaFree(h->d_deg0);
    cudaFree(h->d_upper_base);
    cudaFree(h->d_nbrU);
    cudaFree(h->d_degU);
    cudaFree(h->d_dist0);
    cudaFree(h->d_distU);
    for (void*& sp : h->d_spare) {
        cudaFree(sp);
        sp = nullptr;
    }
    h->spare_rows = 0;
    h->d_nbr0 = h->d_upper_base = h->d_nbrU = nullptr;
    h->d_deg0 = h->d_degU = nullptr;
    h->d_dist0 = h->d_distU = nullptr;
    h->dn_rows = h->dn_slots = 0;
    h->dcap_rows = h->dcap_slots = 0;
    h->dev_valid = false;
}

// (re)create the device copy from the host lists
static int hnsw_dev_upload_all(vecgpu_hnsw* h, bool empty_graph = false) {
    vecgpu_slab* s = h->slab;
    const size_t n = h->node_level.size(), slots = h->degU.size();
    hnsw_dev_free_graph(h);
    cudaGetLastError();
    CU(cudaMalloc(&h->d_nbr0, std::max<size_t>(1, n * h->max_m0) * 4));
    CU(cudaMalloc(&h->d_deg0, std::max<size_t>(1, n) * 2));
    CU(cudaMalloc(&h->d_upper_base, std::max<size_t>(1, n) * 4));
    CU(cudaMalloc(&h->d_nbrU, std::max<size_t>(1, slots * h->M) * 4));
    CU(cudaMalloc(&h->d_degU, std::max<size_t>(1, slots) * 2));
    CU(cudaMalloc(&h->d_dist0, std::max<size_t>(1, n * h->max_m0) * 4));
    CU(cudaMalloc(&h->d_distU, std::max<size_t>(1, slots * h->M) * 4));
    h->dn_rows = h->dcap_rows = n;
    h->dn_slots = h->dcap_slots = slots;
    if (n) CU(cudaMemcpyAsync(h->d_upper_base, h->upper_base.data(), n * 4, cudaMemcpyHostToDevice, s->stream));
    if (empty_graph) {  // start of a rebuild: every list is empty, only the degrees need a defined value
        CU(cudaMemsetAsync(h->d_deg0, 0, std::max<size_t>(1, n) * 2, s->stream));
        CU(cudaMemsetAsync(h->d_degU, 0, std::max<size_t>(1, slots) * 2, s->stream));
    } else {
        if (n) {
            CU(cudaMemcpyAsync(h->d_nbr0, h->nbr0.data(), n * h->max_m0 * 4, cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(h->d_deg0, h->deg0.data(), n * 2, cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(h->d_dist0, h->dist0.data(), n * h->max_m0 * 4, cudaMemcpyHostToDevice, s->stream));
        }
        if (slots) {
            CU(cudaMemcpyAsync(h->d_nbrU, h->nbrU.data(), slots * h->M * 4, cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(h->d_degU, h->degU.data(), slots * 2, cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(h->d_distU, h->distU.data(), slots * h->M * 4, cudaMemcpyHostToDevice, s->stream));
        }
    }
    CU(cudaStreamSynchronize(s->stream));
    h->dirty0.assign(n, 0);
    h->dirtyU.assign(slots, 0);
    h->dirty0_list.clear();
    h->dirtyU_list.clear();
    h->dev_valid = true;
    return 0;
}

// Extend the device lists for rows appended to the slab (incremental inserts): capacity grows geometrically, the existing
// lists stay where they are (or move device to device), the new nodes start with empty lists.  No host round trip of the graph.
static int hnsw_dev_grow(vecgpu_hnsw* h, size_t n_old, size_t slots_old, size_t n, size_t slots) {
    vecgpu_slab* s = h->slab;
    auto regrow = [&](void** p, size_t old_bytes, size_t new_bytes) -> int {
        void* np = nullptr;
        CU(cudaMalloc(&np, std::max<size_t>(16, new_bytes)));
        if (*p && old_bytes) CU(cudaMemcpyAsync(np, *p, old_bytes, cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (*p) cudaFree(*p);
        *p = np;
        return 0;
    };
    int rc;
    if (n > h->dcap_rows) {
        const size_t cap = std::max(n, h->dcap_rows + h->dcap_rows / 2);
        if ((rc = regrow((void**)&h->d_nbr0, n_old * h->max_m0 * 4, cap * h->max_m0 * 4))) return rc;
        if ((rc = regrow((void**)&h->d_dist0, n_old * h->max_m0 * 4, cap * h->max_m0 * 4))) return rc;
        if ((rc = regrow((void**)&h->d_deg0, n_old * 2, cap * 2))) return rc;
        if ((rc = regrow((void**)&h->d_upper_base, n_old * 4, cap * 4))) return rc;
        h->dcap_rows = cap;
    }
    if (slots > h->dcap_slots) {
        const size_t cap = std::max(slots, h->dcap_slots + h->dcap_slots / 2);
        if ((rc = regrow((void**)&h->d_nbrU, slots_old * h->M * 4, cap * h->M * 4))) return rc;
        if ((rc = regrow((void**)&h->d_distU, slots_old * h->M * 4, cap * h->M * 4))) return rc;
        if ((rc = regrow((void**)&h->d_degU, slots_old * 2, cap * 2))) return rc;
        h->dcap_slots = cap;
    }
    if (n > n_old) {
        CU(cudaMemcpyAsync(h->d_upper_base + n_old, h->upper_base.data() + n_old, (n - n_old) * 4, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemsetAsync(h->d_deg0 + n_old, 0, (n - n_old) * 2, s->stream));
    }
    if (slots > slots_old) CU(cudaMemsetAsync(h->d_degU + slots_old, 0, (slots - slots_old) * 2, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    h->dn_rows = n;
    h->dn_slots = slots;
    h->dirty0.resize(n, 0);
    h->dirtyU.resize(slots, 0);
    return 0;
}

// bring the host lists up to date after a rebuild that linked on the device
static int hnsw_ensure_host(vecgpu_hnsw* h) {
    if (!h->host_stale) return 0;
    vecgpu_slab* s = h->slab;
    const size_t n = h->node_level.size(), slots = h->degU.size();
    if (n) {
        CU(cudaMemcpyAsync(h->nbr0.data(), h->d_nbr0, n * h->max_m0 * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h->dist0.data(), h->d_dist0, n * h->max_m0 * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h->deg0.data(), h->d_deg0, n * 2, cudaMemcpyDeviceToHost, s->stream));
    }
    if (slots) {
        CU(cudaMemcpyAsync(h->nbrU.data(), h->d_nbrU, slots * h->M * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h->distU.data(), h->d_distU, slots * h->M * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(h->degU.data(), h->d_degU, slots * 2, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    h->host_stale = false;
    return 0;
}

static int hnsw_sw_reserve(vecgpu_hnsw* h, size_t bytes) {
    if (bytes > h->sw_cap) {
        if (h->d_sw) CU(cudaFree(h->d_sw));
        h->d_sw = nullptr;
        h->sw_cap = std::max(bytes * 2, (size_t)1 << 20);
        CU(cudaMalloc(&h->d_sw, h->sw_cap));
    }
    if (bytes > h->hsw_cap) {
        if (h->h_sw) CU(cudaFreeHost(h->h_sw));
        h->h_sw = nullptr;
        h->hsw_cap = std::max(bytes * 2, (size_t)1 << 20);
        CU(cudaMallocHost(&h->h_sw, h->hsw_cap));
    }
    return 0;
}

// push the lists changed since the last flush: [id][deg][width ids] items, scattered by hnsw_scatter_kernel
static int hnsw_dev_flush_dirty(vecgpu_hnsw* h) {
    vecgpu_slab* s = h->slab;
    for (int upper = 0; upper < 2; ++upper) {
        std::vector<uint32_t>& list = upper ? h->dirtyU_list : h->dirty0_list;
        if (list.empty()) continue;
        const uint32_t width = upper ? h->M : h->max_m0;
        const size_t item = (size_t)width + 2, bytes = list.size() * item * 4;
        int rc = hnsw_sw_reserve(h, bytes);
        if (rc) return rc;
        uint32_t* st = (uint32_t*)h->h_sw;
        const uint32_t* nbr = upper ? h->nbrU.data() : h->nbr0.data();
        const uint16_t* deg = upper ? h->degU.data() : h->deg0.data();
        uint8_t* flags = upper ? h->dirtyU.data() : h->dirty0.data();
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)list.size(); ++i) {
            const uint32_t id = list[i];
            uint32_t* d = st + (size_t)i * item;
            d[0] = id;
            d[1] = deg[id];
            memcpy(d + 2, nbr + (size_t)id * width, (size_t)width * 4);
            flags[id] = 0;
        }
        CU(cudaMemcpyAsync(h->d_sw, st, bytes, cudaMemcpyHostToDevice, s->stream));
        const uint32_t n_items = (uint32_t)list.size();
        const uint32_t blocks = std::max(1u, std::min((n_items + 7) / 8, (uint32_t)s->num_sms * 8));
        hnsw_scatter_kernel<<<blocks, 256, 0, s->stream>>>((const uint32_t*)h->d_sw, n_items, width, upper ? h->d_nbrU : h->d_nbr0,
                                                            upper ? h->d_degU : h->d_deg0);
        LAUNCHED();
        CU(cudaStreamSynchronize(s->stream));  // the staging buffers are reused right away
        list.clear();
    }
    return 0;
}

struct HDevOut {  // host view of one launch's results (pinned memory, valid until the next launch)
    const uint32_t* cnt = nullptr;
    const uint32_t* status = nullptr;
    const uint64_t* keys = nullptr;
    uint32_t take = 0;
    // device twins (valid until the next launch on the stream)
    const uint32_t* d_ai = nullptr;
    const uint32_t* d_off = nullptr;
    const int8_t* d_lvl = nullptr;
    const uint32_t* d_cnt = nullptr;
    const uint64_t* d_keys = nullptr;
    size_t keys_bytes = 0;
};

template <class T>
static int hnsw_dev_launch_t(vecgpu_hnsw* h, HSearchParams& p, size_t per_warp) {
    vecgpu_slab* s = h->slab;
    uint32_t wpb = 8;
    while (wpb > 1 && wpb * per_warp > 200 * 1024) wpb >>= 1;
    if (wpb * per_warp > SMEM_MAX) return fail(VECGPU_ERR_INVALID_PARAM, "ef too large for the device search");
    CU(cudaFuncSetAttribute(hnsw_search_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    int bps = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hnsw_search_kernel<T>, (int)wpb * 32, wpb * per_warp));
    int bps_max = 6;
    if (const char* e = getenv("VECGPU_HNSW_BPS")) bps_max = std::max(1, atoi(e));
    bps = std::max(1, std::min(bps, bps_max));
    uint32_t grid = (uint32_t)s->num_sms * (uint32_t)bps;
    if ((uint64_t)grid * wpb > p.nq) {  // few queries: spread them, one or a few warps per CTA
        grid = std::min(grid, p.nq);
        wpb = (p.nq + grid - 1) / grid;
    }
    // Few queries (the SQL case is ONE per MATCH; up to ~1000 in a call): one CTA per query scores all fresh neighbours of an expansion at once and
    // keeps the visited set in shared memory (hnsw_search_cta_kernel).  Only for walks whose beam fits its tables.
    // Measured crossover (tools/hnsw_crossover.py, profiles/r2_hnsw_crossover.txt): ceil(nq / SMs) rounds of ~0.62 ms against the
    // one-warp walk's ~4.5 ms floor (its latency; it only becomes throughput-bound beyond ~4000 queries) meet at 7 rounds.
    const uint32_t cta_max = env_u32("VECGPU_HNSW_CTA_MAX_NQ", 7u * (uint32_t)s->num_sms);
    // visited slots: a walk touches ~ef x 30 nodes; 16 K slots (at most 3/4 used) cover ef <= 256 and leave room for the second row buffer
    p.cta_vis = p.ef_wide <= 256 ? 16384u : HC_VIS;
    size_t cta_smem = (size_t)p.cta_vis * 4 + (size_t)p.cap * 16 + (size_t)((h->max_m0 + 31u) & ~31u) * 12 + (size_t)s->row_stride * (1 + HC_ROWS) + 64;
    p.spec_rows = 0;
    if (p.prefetch && env_u32("VECGPU_HNSW_SPEC_ROWS", 1) && cta_smem + (size_t)s->row_stride * HC_ROWS <= 220 * 1024) {
        p.spec_rows = 1;
        cta_smem += (size_t)s->row_stride * HC_ROWS;
    }
    // (inserts too: a rebuild's first batches and incremental inserts are a handful of walks)
    if (p.nq <= cta_max && p.q_smem && cta_smem <= 220 * 1024 && p.ef_wide <= 512) {
        static int cfg_dev_cta = -1;
        int dev2 = 0;
        CU(cudaGetDevice(&dev2));
        if (cfg_dev_cta != dev2) {
            CU(cudaFuncSetAttribute(hnsw_search_cta_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            CU(cudaFuncSetAttribute(hnsw_search_cta_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            cfg_dev_cta = dev2;
        }
        const uint32_t cta_grid = std::min<uint32_t>(p.nq, (uint32_t)s->num_sms);
        if (p.prof) hnsw_search_cta_kernel<T, true><<<cta_grid, HC_THREADS, cta_smem, s->stream>>>(p);
        else hnsw_search_cta_kernel<T, false><<<cta_grid, HC_THREADS, cta_smem, s->stream>>>(p);
        LAUNCHED();
        return 0;
    }
    const size_t vis_bytes = (size_t)grid * wpb * p.vis_size * 4;
    if (vis_bytes > h->vis_cap) {
        if (h->d_visited) CU(cudaFree(h->d_visited));
        h->d_visited = nullptr;
        h->vis_cap = 0;
        CU(cudaMalloc(&h->d_visited, vis_bytes));
        h->vis_cap = vis_bytes;
    }
    p.visited = h->d_visited;
    hnsw_search_kernel<T><<<grid, wpb * 32, wpb * per_warp, s->stream>>>(p);
    LAUNCHED();
    return 0;
}

// One launch: every query walks all its layers on the device.  a_index / node_level / out_off are host arrays (or NULL).
static int hnsw_dev_search(vecgpu_hnsw* h, const uint8_t* a_base, const uint32_t* a_index, const int8_t* node_level, uint32_t nq,
                           uint32_t ef_wide, uint32_t take, const uint32_t* out_off, uint32_t n_slots, HDevOut* out, bool fetch_keys = true) {
    vecgpu_slab* s = h->slab;
    int rc;
    if (!h->dev_valid && (rc = hnsw_dev_upload_all(h))) return rc;
    if ((rc = hnsw_dev_flush_dirty(h))) return rc;
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    // upload block: [a_index][out_off][node_level]   download block: [cnt][status][scored][keys]
    const size_t o_ai = 0, o_off = al(o_ai + (size_t)nq * 4), o_lvl = al(o_off + (size_t)nq * 4), up_end = al(o_lvl + nq);
    const size_t o_cnt = up_end, o_status = al(o_cnt + (size_t)n_slots * 4), o_scored = al(o_status + (size_t)nq * 4),
                 o_next = o_scored + 8, o_hist = o_scored + 16, o_prof = o_hist + 40, o_keys = al(o_prof + 160), total = o_keys + (size_t)n_slots * take * 8;
    if ((rc = hnsw_sw_reserve(h, total))) return rc;
    uint8_t* hp = h->h_sw;
    uint8_t* dp = h->d_sw;
    if (a_index) memcpy(hp + o_ai, a_index, (size_t)nq * 4);
    if (out_off) memcpy(hp + o_off, out_off, (size_t)nq * 4);
    if (node_level) memcpy(hp + o_lvl, node_level, nq);
    CU(cudaMemcpyAsync(dp, hp, up_end, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemsetAsync(dp + o_cnt, 0, o_keys - o_cnt, s->stream));

    HSearchParams p{};
    p.g.nbr0 = h->d_nbr0;
    p.g.deg0 = h->d_deg0;
    p.g.upper_base = h->d_upper_base;
    p.g.nbrU = h->d_nbrU;
    p.g.degU = h->d_degU;
    p.g.max_m0 = h->max_m0;
    p.g.M = h->M;
    p.a_base = a_base;
    p.a_stride = s->row_stride;
    p.a_index = a_index ? (const uint32_t*)(dp + o_ai) : nullptr;
    p.b_base = s->d_vec;
    p.b_stride = s->row_stride;
    p.units = s->row_stride / 16;
    p.qc_kind = s->elem == VECGPU_I8 ? 1u : 0u;
    p.nq = nq;
    p.entry = (uint32_t)h->entry;
    p.entry_level = h->entry_level;
    p.node_level = node_level ? (const int8_t*)(dp + o_lvl) : nullptr;
    p.ef_wide = ef_wide;
    p.cap = ((2 * ef_wide + 32 + 31) / 32) * 32;  // the array never holds more than 2 ef - 1 entries (ties behind ef < ef)
    p.take = take;
    p.vis_size = std::min<uint32_t>(1u << 18, std::max<uint32_t>(4096u, next_pow2((uint32_t)std::min<uint64_t>(1u << 18, 4ull * ef_wide * h->max_m0))));
    if (const char* e = getenv("VECGPU_HNSW_VIS_LOG2")) p.vis_size = 1u << std::max(12, std::min(18, atoi(e)));
    p.out_off = out_off ? (const uint32_t*)(dp + o_off) : nullptr;
    p.out_cnt = (uint32_t*)(dp + o_cnt);
    p.status = (uint32_t*)(dp + o_status);
    p.scored = (unsigned long long*)(dp + o_scored);
    p.hist = (unsigned long long*)(dp + o_hist);
    p.batch_admit = env_u32("VECGPU_HNSW_BATCH_ADMIT", 1);
    p.prefetch = env_u32("VECGPU_HNSW_PREFETCH", 1);
    const bool prof = getenv("VECGPU_HNSW_TIMING") != nullptr;
    p.prof = prof ? (unsigned long long*)(dp + o_prof) : nullptr;
    p.next_q = (unsigned int*)(dp + o_next);
    p.out_keys = (uint64_t*)(dp + o_keys);
    p.max_steps = 1u << 20;
    p.q_smem = s->row_stride <= 16384 ? 1u : 0u;
    const size_t per_warp = (size_t)p.cap * 8 + (size_t)((h->max_m0 + 31u) & ~31u) * 4 + (p.q_smem ? (size_t)s->row_stride : 0);
    const int elem = s->elem, metric = h->metric;
    if (elem == VECGPU_F32) {
        if (metric == VECGPU_L2) rc = hnsw_dev_launch_t<F32L2<1>>(h, p, per_warp);
        else if (metric == VECGPU_L1) rc = hnsw_dev_launch_t<F32L1<1>>(h, p, per_warp);
        else rc = hnsw_dev_launch_t<F32Cos<1>>(h, p, per_warp);
    } else if (elem == VECGPU_I8) {
        if (metric == VECGPU_L2) rc = hnsw_dev_launch_t<I8Dot<1, false>>(h, p, per_warp);
        else if (metric == VECGPU_L1) rc = hnsw_dev_launch_t<I8L1<1>>(h, p, per_warp);
        else rc = hnsw_dev_launch_t<I8Dot<1, true>>(h, p, per_warp);
    } else {
        rc = hnsw_dev_launch_t<BitHamming<1>>(h, p, per_warp);
    }
    if (rc) return rc;
    CU(cudaMemcpyAsync(hp + o_cnt, dp + o_cnt, (fetch_keys ? total : o_keys) - o_cnt, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    h->scored += *(const unsigned long long*)(hp + o_scored);
    for (int b = 0; b < 5; ++b) h->batch_hist[b] += ((const unsigned long long*)(hp + o_hist))[b];
    if (prof) {
        const unsigned long long* pf = (const unsigned long long*)(hp + o_prof);
        if (pf[0] | pf[1] | pf[2] | pf[3])
            fprintf(stderr, "[vecgpu] hnsw CTA walk, %u queries, cycles: row fetch %llu, scoring %llu, admission %llu, pop+adjacency+visited %llu; admission batches merged %llu, one by one %llu; expansions with rows staged ahead %llu of %llu; admission split: prediction %llu, ranks %llu, placement %llu, finish %llu; since the end of scoring (summed): helper released %llu, "
                    "helper list ready %llu, helper copies issued %llu, warp 0 at the end barrier %llu, barrier passed %llu, rows landed %llu\n",
                    nq, pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7], pf[8], pf[9], pf[10], pf[11], pf[12], pf[13], pf[14], pf[15], pf[16], pf[17]);
    }
    h->rounds += 1;
    h->dev_launches += 1;
    h->dev_queries += nq;
    out->cnt = (const uint32_t*)(hp + o_cnt);
    out->status = (const uint32_t*)(hp + o_status);
    out->keys = (const uint64_t*)(hp + o_keys);
    out->take = take;
    out->d_ai = (const uint32_t*)(dp + o_ai);
    out->d_off = (const uint32_t*)(dp + o_off);
    out->d_lvl = (const int8_t*)(dp + o_lvl);
    out->d_cnt = (const uint32_t*)(dp + o_cnt);
    out->d_keys = (const uint64_t*)(dp + o_keys);
    out->keys_bytes = total - o_keys;
    return 0;
}

// link a whole batch on the device from the search kernel's results (still in the launch workspace)
static int hnsw_dev_link(vecgpu_hnsw* h, const HDevOut& o, uint32_t nq, uint32_t n_slots, int search_entry_level) {
    vecgpu_slab* s = h->slab;
    const uint32_t n_ops = n_slots * o.take;
    if (n_ops == 0) return 0;
    size_t sort_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint64_t*)nullptr,
                                       (uint64_t*)nullptr, (int)n_ops, 0, 64, s->stream));
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t arr = al((size_t)n_ops * 8), need = 4 * arr + al(sort_bytes);
    if (need > h->link_cap) {
        if (h->d_link) CU(cudaFree(h->d_link));
        h->d_link = nullptr;
        h->link_cap = std::max(need * 2, (size_t)1 << 20);
        CU(cudaMalloc(&h->d_link, h->link_cap));
    }
    HLinkParams p{};
    p.nbr0 = h->d_nbr0;
    p.dist0 = h->d_dist0;
    p.deg0 = h->d_deg0;
    p.upper_base = h->d_upper_base;
    p.nbrU = h->d_nbrU;
    p.distU = h->d_distU;
    p.degU = h->d_degU;
    p.max_m0 = h->max_m0;
    p.M = h->M;
    p.n_rows = (uint32_t)h->node_level.size();
    p.a_index = o.d_ai;
    p.out_off = o.d_off;
    p.node_level = o.d_lvl;
    p.entry_level = search_entry_level;
    p.nq = nq;
    p.take = o.take;
    p.out_keys = o.d_keys;
    p.out_cnt = o.d_cnt;
    p.op_key = (uint64_t*)h->d_link;
    p.op_val = (uint64_t*)(h->d_link + arr);
    p.n_ops = n_ops;
    uint64_t* skey = (uint64_t*)(h->d_link + 2 * arr);
    uint64_t* sval = (uint64_t*)(h->d_link + 3 * arr);
    hnsw_link_forward_kernel<<<std::max(1u, std::min((nq + 7) / 8, (uint32_t)s->num_sms * 8)), 256, 0, s->stream>>>(p);
    LAUNCHED();
    CU(cub::DeviceRadixSort::SortPairs(h->d_link + 4 * arr, sort_bytes, p.op_key, skey, p.op_val, sval, (int)n_ops, 0, 64, s->stream));
    g_launches.fetch_add(4, std::memory_order_relaxed);
    hnsw_link_reverse_kernel<<<std::max(1u, std::min((n_ops + 255) / 256, (uint32_t)s->num_sms * 8)), 256, 0, s->stream>>>(p, skey, sval);
    LAUNCHED();
    h->host_stale = true;
    return 0;
}

static inline HCand h_decode(uint64_t key) { return HCand{order_bits_inv((uint32_t)(key >> 32)), (uint32_t)(key & 0xFFFFFFFFull) >> 1}; }

// ---- C ABI ------------------------------------------------------------------------------------------------
static int64_t h_rowid_of(const vecgpu_slab* s, uint32_t pos) { return s->dense ? s->first_rowid + (int64_t)pos : s->h_rowids[pos]; }

// The slab an HNSW index of a float32 column is built over: the STORED representation of src/hnsw/insert.rs:300-322 —
// normalised for cosine columns, int8-quantised with index_quantization=int8 — produced on the device from the column's
// slab (same rowids, same positions; rows the reference could not store are skipped).  *out = NULL when the column slab
// itself is the stored representation (no normalisation, no quantisation).
extern "C" int vecgpu_hnsw_stored_slab(vecgpu_slab* col, int normalize, int int8_quantization, vecgpu_slab** out) {
    VG_TRY
    if (!col || !out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    *out = nullptr;
    if (!normalize && !int8_quantization) return 0;
    if (col->elem != VECGPU_F32)  // search.rs:291-301 / insert.rs:303-322 apply both steps to Float32 columns only
        return fail(VECGPU_ERR_UNSUPPORTED, "normalisation / int8 index quantisation apply to float32 columns only");
    std::lock_guard<std::mutex> lk(col->mu);
    int rc = use_device(col->device);
    if (rc) return rc;
    vecgpu_slab* st = nullptr;
    rc = vecgpu_slab_create(int8_quantization ? VECGPU_I8 : VECGPU_F32, col->dims, col->rows, col->device, &st);
    if (rc) return rc;
    struct Guard {
        vecgpu_slab* s;
        ~Guard() { if (s) vecgpu_slab_destroy(s); }
    } guard{st};
    const uint64_t n = col->rows;
    st->rows = n;
    st->dense = col->dense;
    st->first_rowid = col->first_rowid;
    st->h_rowids = col->h_rowids;
    st->rowids_synced = 0;
    if ((rc = slab_sync_rowids(st))) return rc;
    if (n) {
        uint8_t* d_flags = nullptr;
        CU(cudaMalloc((void**)&d_flags, n));
        struct FreeDev {
            void* p;
            ~FreeDev() { cudaFree(p); }
        } fd{d_flags};
        hnsw_stored_rows_kernel<<<(uint32_t)std::min<uint64_t>((n + 127) / 128, (uint64_t)col->num_sms * 16), 128, 0, col->stream>>>(
            col->d_vec, col->row_stride, n, col->dims, normalize ? 1 : 0, int8_quantization ? 1 : 0, st->d_vec, st->row_stride,
            col->n_skip ? col->d_skip : nullptr, d_flags);
        LAUNCHED();
        st->h_skip.resize(n);
        CU(cudaMemcpyAsync(st->h_skip.data(), d_flags, n, cudaMemcpyDeviceToHost, col->stream));
        CU(cudaStreamSynchronize(col->stream));
        uint64_t dead = 0;
        for (uint64_t i = 0; i < n; ++i) dead += st->h_skip[i] != 0;
        st->n_skip = dead;
        if (!dead) st->h_skip.clear();
        st->skip_synced = 0;
        if ((rc = slab_sync_skip(st))) return rc;
    }
    guard.s = nullptr;
    *out = st;
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_hnsw_create(vecgpu_slab* slab, int metric, uint32_t M, uint32_t ef_construction, uint64_t seed,
                                  vecgpu_hnsw** out) {
    VG_TRY
    if (!slab || !out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    *out = nullptr;
    int rc = check_pair(slab->elem, metric);
    if (rc) return rc;
    // argument ranges of vec_rebuild_hnsw (src/sql_functions.rs:442-469)
    if (M < 2 || M > 100) return fail(VECGPU_ERR_INVALID_PARAM, "M must be between 2 and 100");
    if (ef_construction < 10 || ef_construction > 2000) return fail(VECGPU_ERR_INVALID_PARAM, "ef_construction must be between 10 and 2000");
    vecgpu_hnsw* h = new (std::nothrow) vecgpu_hnsw();
    if (!h) return fail(VECGPU_ERR_CUDA, "out of host memory");
    h->slab = slab;
    h->metric = metric;
    h->M = M;
    h->max_m0 = 2 * M;  // src/sql_functions.rs:489-505: max_m0 = 2m
    h->efc = ef_construction;
    h->seed = seed;
    h->level_factor = 1.0 / std::log((double)M);
    *out = h;
    return 0;
    VG_CATCH
}

extern "C" void vecgpu_hnsw_destroy(vecgpu_hnsw* h) {
    VG_TRY
    if (!h) return;
    if (h->slab) cudaSetDevice(h->slab->device);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    cudaFree(h->d_buf);
    cudaFree(h->d_queries);
    hnsw_dev_free_graph(h);
    cudaFree(h->d_visited);
    cudaFree(h->d_sw);
    cudaFree(h->d_link);
    if (h->h_sw) cudaFreeHost(h->h_sw);
    cudaGetLastError();
    delete h;
    VG_CATCH_VOID
}

static void hq_init(vecgpu_hnsw* h, HQuery& q, uint32_t a_index, int node_level, uint32_t ef_wide) {
    q.a_index = a_index;
    q.node_level = node_level;
    q.level = h->entry_level;
    q.stop_level = 0;
    q.ef_wide = ef_wide;
    const bool wide = node_level < 0 ? q.level == 0 : q.level <= node_level;
    q.ef = wide ? ef_wide : 1;
    q.entry = (uint32_t)h->entry;
    q.done = false;
    q.layers.clear();
    if (q.visited.tab.empty()) q.visited.init(1024);
    hq_start_layer(q);
}

// vec_rebuild_hnsw: (re)build the graph over every live row of the slab, `batch` inserts in lockstep — or, `incremental`,
// continue the same insertion loop over the rows appended to the slab since the graph was last built or extended
// (insert_hnsw for rows that arrive in rowid order, src/hnsw/insert.rs:279-532: the loop is the rebuild's, so a batch of one
// continues the strictly sequential build).  Caller holds both mutexes.
static int hnsw_build_locked(vecgpu_hnsw* h, uint32_t batch, bool incremental, uint64_t* n_inserted,
                             const std::vector<uint32_t>* only_nodes = nullptr) {
    vecgpu_slab* s = h->slab;
    int rc = use_device(s->device);
    if (rc) return rc;
    if (batch == 0) batch = 16384;  // measured on 1 M x 384: 3.7 / 2.8 / 2.5 / 2.2 s at 4096 / 8192 / 16384 / 32768, same recall
    const uint64_t n = s->rows;
    if (n >= 0xFFFFFFFFull) return fail(VECGPU_ERR_INVALID_PARAM, "too many rows");
    const bool use_dev = hnsw_device_enabled(h);
    uint64_t pos0 = 0;
    if (n_inserted) *n_inserted = 0;
    if (incremental && h->entry >= 0) {
        if (h->slab_gen != s->layout_gen)
            return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction, reload or out-of-order insert) since this HNSW index was built: rebuild it");
        pos0 = h->node_level.size();
        if (pos0 > n) return fail(VECGPU_ERR_CUDA, "the slab has fewer rows than the index: rebuild it");
        if (pos0 == n && !only_nodes) return 0;
        const bool grow_on_device = use_dev && h->dev_valid;
        if (grow_on_device) {
            if ((rc = hnsw_dev_flush_dirty(h))) return rc;  // the device copy is the current one from here on
        } else if ((rc = hnsw_ensure_host(h))) {            // the lists are extended on the host, then uploaded
            return rc;
        }
        const uint64_t slots_old = h->degU.size();
        uint64_t upper_slots = slots_old;
        h->node_level.resize(n, 0);
        h->in_graph.resize(n, 0);
        h->upper_base.resize(n, 0);
        for (uint64_t pos = pos0; pos < n; ++pos) {
            const int L = hnsw_level_for(h, pos);
            h->node_level[pos] = (int8_t)L;
            h->upper_base[pos] = (uint32_t)upper_slots;
            upper_slots += (uint64_t)L;
        }
        h->nbr0.resize((size_t)n * h->max_m0, 0);
        h->dist0.resize((size_t)n * h->max_m0, 0.f);
        h->deg0.resize(n, 0);
        h->nbrU.resize((size_t)upper_slots * h->M, 0);
        h->distU.resize((size_t)upper_slots * h->M, 0.f);
        h->degU.resize(upper_slots, 0);
        if (grow_on_device) {
            if ((rc = hnsw_dev_grow(h, pos0, slots_old, n, upper_slots))) return rc;
        } else if (use_dev) {
            if ((rc = hnsw_dev_upload_all(h, false))) return rc;
        } else {
            hnsw_dev_free_graph(h);
        }
    } else {
        h->n_nodes = 0;
        h->entry = -1;
        h->entry_level = -1;
        h->scored = h->rounds = 0;
        for (int b = 0; b < 5; ++b) h->batch_hist[b] = 0;
        h->slab_gen = s->layout_gen;
        h->node_level.assign(n, 0);
        h->in_graph.assign(n, 0);
        h->upper_base.assign(n, 0);
        uint64_t upper_slots = 0;
        for (uint64_t pos = 0; pos < n; ++pos) {
            const int L = hnsw_level_for(h, pos);
            h->node_level[pos] = (int8_t)L;
            h->upper_base[pos] = (uint32_t)upper_slots;
            upper_slots += (uint64_t)L;
        }
        h->nbr0.assign((size_t)n * h->max_m0, 0);
        h->dist0.assign((size_t)n * h->max_m0, 0.f);
        h->deg0.assign(n, 0);
        h->nbrU.assign((size_t)upper_slots * h->M, 0);
        h->distU.assign((size_t)upper_slots * h->M, 0.f);
        h->degU.assign(upper_slots, 0);
        h->host_stale = false;  // the host lists were just reset
        if (use_dev) {
            if ((rc = hnsw_dev_upload_all(h, true))) return rc;  // empty lists; kept in sync batch by batch
        } else {
            hnsw_dev_free_graph(h);
        }
    }
    // device search + device linking by default; VECGPU_HNSW_LINK=host keeps the ordered edge replay on the host threads
    const char* link_env = getenv("VECGPU_HNSW_LINK");
    const bool dev_link = use_dev && !(link_env && link_env[0] == 'h') && batch <= 32768 && h->max_m0 <= 256;
    const bool timing = getenv("VECGPU_HNSW_TIMING") != nullptr;
    double t_search = 0, t_decode = 0, t_ops = 0, t_link = 0, t_flush = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

    struct HOp {
        uint32_t from, to;
        float d;
        int32_t level;
    };
    std::vector<HQuery> qs, fq;
    std::vector<uint32_t> nodes, off, fb;
    std::vector<int8_t> lv8;
    std::vector<HOp> ops;
    const int nthreads = std::max(1, omp_get_max_threads());
    std::vector<std::vector<uint32_t>> t_dirty0(nthreads), t_dirtyU(nthreads);
    uint64_t pos = only_nodes ? n : pos0;  // only_nodes: (re)insert exactly these row positions, as one batch
    bool explicit_batch = only_nodes != nullptr;
    const uint64_t nodes_before = h->n_nodes;
    while (explicit_batch || pos < n) {
        const uint64_t want = std::min<uint64_t>(batch, std::max<uint64_t>(1, h->n_nodes / 4));
        nodes.clear();
        if (explicit_batch) {
            explicit_batch = false;
            for (uint32_t p : *only_nodes) {
                nodes.push_back(p);
                h->in_graph[p] = 1;
            }
        }
        while (pos < n && nodes.size() < want) {
            if (s->h_skip.empty() || !s->h_skip[pos]) {
                nodes.push_back((uint32_t)pos);
                h->in_graph[pos] = 1;
            }
            ++pos;
        }
        if (nodes.empty()) break;
        size_t first = 0;
        if (h->entry < 0) {  // very first node: it becomes the entry point, nothing to search
            h->entry = nodes[0];
            h->entry_level = h->node_level[nodes[0]];
            h->n_nodes = 1;
            first = 1;
        }
        const size_t nb = nodes.size() - first;
        if (nb == 0) continue;
        if (qs.size() < nb) qs.resize(nb);
        const int search_entry_level = h->entry_level;
        if (use_dev) {
            // ---- one launch: every insert of the batch walks all its layers on the device (K6)
            off.resize(nb);
            lv8.resize(nb);
            uint32_t slots = 0;
            for (size_t i = 0; i < nb; ++i) {
                const int L = h->node_level[nodes[first + i]];
                lv8[i] = (int8_t)L;
                off[i] = slots;
                slots += (uint32_t)std::min(L, search_entry_level) + 1u;
            }
            HDevOut o;
            double t0 = now();
            rc = hnsw_dev_search(h, s->d_vec, nodes.data() + first, lv8.data(), (uint32_t)nb, h->efc, h->max_m0, off.data(), slots, &o, !dev_link);
            if (rc) return rc;
            t_search += now() - t0;
            t0 = now();
            if (dev_link) {
                bool any_fb = false;
                for (size_t i = 0; i < nb && !any_fb; ++i) any_fb = o.status[i] != 0;
                if (!any_fb) {
                    // ---- the common case: link on the device too; the host only tracks the entry point
                    if ((rc = hnsw_dev_link(h, o, (uint32_t)nb, slots, search_entry_level))) return rc;
                    for (size_t i = 0; i < nb; ++i) {
                        const uint32_t node = nodes[first + i];
                        const int L = h->node_level[node];
                        if (L > h->entry_level) {
                            h->entry = node;
                            h->entry_level = L;
                        }
                    }
                    h->n_nodes += nb;
                    t_link += now() - t0;
                    continue;
                }
                // a query overflowed a device capacity: this batch is linked by the host loop (which needs the keys and
                // current host lists); the device copy is rebuilt from the host afterwards
                CU(cudaMemcpyAsync((void*)o.keys, o.d_keys, o.keys_bytes, cudaMemcpyDeviceToHost, s->stream));
                CU(cudaStreamSynchronize(s->stream));
                if ((rc = hnsw_ensure_host(h))) return rc;
            }
            fb.clear();
            for (size_t i = 0; i < nb; ++i) {
                HQuery& q = qs[i];
                q.layers.clear();
                if (o.status[i]) {
                    fb.push_back((uint32_t)i);
                    continue;
                }
                const int nl = std::min((int)lv8[i], search_entry_level) + 1;
                q.layers.resize(nl);
                for (int lv = 0; lv < nl; ++lv) {
                    const uint32_t slot = off[i] + (uint32_t)lv, c = o.cnt[slot];
                    q.layers[lv].resize(c);
                    for (uint32_t j = 0; j < c; ++j) q.layers[lv][j] = h_decode(o.keys[(size_t)slot * o.take + j]);
                }
            }
            if (!fb.empty()) {  // capacity overflow on the device: the lockstep driver answers those
                h->dev_fallbacks += fb.size();
                if (fq.size() < fb.size()) fq.resize(fb.size());
                for (size_t j = 0; j < fb.size(); ++j) hq_init(h, fq[j], nodes[first + fb[j]], h->node_level[nodes[first + fb[j]]], h->efc);
                rc = hnsw_run_batch(h, fq, fb.size(), s->d_vec);
                if (rc) return rc;
                for (size_t j = 0; j < fb.size(); ++j) qs[fb[j]].layers = fq[j].layers;
            }
            t_decode += now() - t0;
        } else {
            for (size_t i = 0; i < nb; ++i) hq_init(h, qs[i], nodes[first + i], h->node_level[nodes[first + i]], h->efc);
            rc = hnsw_run_batch(h, qs, nb, s->d_vec);
            if (rc) return rc;
        }
        // ---- link, in insertion order (insert.rs:408-498).  The edge operations are listed in order, then every
        //      host thread replays the list and applies the operations of the adjacency lists it owns: each list
        //      sees its operations in the sequential order, so the result equals the one-thread loop.
        double t1 = now();
        ops.clear();
        for (size_t i = 0; i < nb; ++i) {
            HQuery& q = qs[i];
            const uint32_t node = nodes[first + i];
            const int L = h->node_level[node];
            for (int lv = std::min(L, search_entry_level); lv >= 0; --lv) {
                if ((int)q.layers.size() <= lv) continue;
                const uint32_t maxc = lv == 0 ? h->max_m0 : h->M;
                const std::vector<HCand>& w = q.layers[lv];
                const size_t take = std::min<size_t>(maxc, w.size());
                for (size_t j = 0; j < take; ++j) {
                    if (w[j].node == node) continue;
                    ops.push_back(HOp{node, w[j].node, w[j].d, lv});
                    ops.push_back(HOp{w[j].node, node, w[j].d, lv});
                }
            }
            if (L > h->entry_level) {
                h->entry = node;
                h->entry_level = L;
            }
        }
        t_ops += now() - t1;
        t1 = now();
#pragma omp parallel num_threads(nthreads)
        {
            const uint32_t tid = (uint32_t)omp_get_thread_num(), nt = (uint32_t)omp_get_num_threads();
            std::vector<uint32_t>* d0 = use_dev ? &t_dirty0[tid] : nullptr;
            std::vector<uint32_t>* dU = use_dev ? &t_dirtyU[tid] : nullptr;
            for (const HOp& op : ops)
                if (((op.from * 2654435761u) >> 12) % nt == tid) h_add_edge(h, op.from, op.to, op.d, op.level, d0, dU);
        }
        if (use_dev)
            for (int t = 0; t < nthreads; ++t) {
                h->dirty0_list.insert(h->dirty0_list.end(), t_dirty0[t].begin(), t_dirty0[t].end());
                h->dirtyU_list.insert(h->dirtyU_list.end(), t_dirtyU[t].begin(), t_dirtyU[t].end());
                t_dirty0[t].clear();
                t_dirtyU[t].clear();
            }
        if (dev_link) h->dev_valid = false;  // (overflow batch) the stored distances on the device are stale: full re-upload
        t_link += now() - t1;
        h->n_nodes += nb;
    }
    double t2 = now();
    if (use_dev && h->dev_valid && (rc = hnsw_dev_flush_dirty(h))) return rc;
    t_flush += now() - t2;
    if (timing)
        fprintf(stderr, "[vecgpu hnsw build] search+flush %.3f s  decode %.3f s  ops %.3f s  link %.3f s  (threads %d, %s linking)\n",
                t_search + t_flush, t_decode, t_ops, t_link, nthreads, dev_link ? "device" : "host");
    if (n_inserted) *n_inserted = h->n_nodes - nodes_before;
    return 0;
}

extern "C" int vecgpu_hnsw_build(vecgpu_hnsw* h, uint32_t batch) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lk(h->slab->mu);
    std::lock_guard<std::mutex> lk2(h->mu);
    return hnsw_build_locked(h, batch, false, nullptr);
    VG_CATCH
}

extern "C" int vecgpu_hnsw_insert_appended(vecgpu_hnsw* h, uint32_t batch, uint64_t* n_inserted) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lk(h->slab->mu);
    std::lock_guard<std::mutex> lk2(h->mu);
    return hnsw_build_locked(h, batch, true, n_inserted);
    VG_CATCH
}

// Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): the node and every edge from or to it are deleted, then the
// row — whose vector in the slab has just been replaced by vecgpu_slab_upsert — is inserted again (insert_hnsw).  The node
// keeps its row position and therefore its level.  If the row is now deleted or empty it only leaves the graph.  The entry
// point, when it is the node itself, passes to the highest remaining node (the first such position) for the re-insertion.
// With the graph resident the edges are removed by hnsw_unlink_kernel; in lockstep mode (VECGPU_HNSW_DEVICE=0) on the host lists.
static int hnsw_reinsert_locked(vecgpu_hnsw* h, int64_t rowid) {
    vecgpu_slab* s = h->slab;
    int rc = use_device(s->device);
    if (rc) return rc;
    if (h->entry < 0) return fail(VECGPU_ERR_INVALID_PARAM, "the index is empty: build it first");
    if (h->slab_gen != s->layout_gen)
        return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction, reload or out-of-order insert) since this HNSW index was built: rebuild it");
    const int64_t p64 = slab_find(s, rowid, nullptr);
    if (p64 < 0) return fail(VECGPU_ERR_INVALID_PARAM, "rowid %lld is not in the slab", (long long)rowid);
    const uint64_t rows = h->node_level.size();
    if ((uint64_t)p64 >= rows) return fail(VECGPU_ERR_INVALID_PARAM, "rowid %lld was appended after the index was built: use vecgpu_hnsw_insert_appended", (long long)rowid);
    const uint32_t pos = (uint32_t)p64;
    const bool on_device = hnsw_device_enabled(h) && h->dev_valid;
    if (on_device) {
        if (h->in_graph[pos]) {
            // the edges are removed where the lists live: one thread per list, no round trip of the graph
            if ((rc = hnsw_dev_flush_dirty(h))) return rc;
            const uint32_t L = (uint32_t)h->node_level[pos];
            hnsw_unlink_kernel<<<(uint32_t)std::min<uint64_t>((rows + 255) / 256, (uint64_t)s->num_sms * 16), 256, 0, s->stream>>>(
                h->d_nbr0, h->d_dist0, h->d_deg0, rows, h->max_m0, pos, pos, 1);
            LAUNCHED();
            const uint64_t slots = h->degU.size();
            if (slots) {
                hnsw_unlink_kernel<<<(uint32_t)std::min<uint64_t>((slots + 255) / 256, (uint64_t)s->num_sms * 16), 256, 0, s->stream>>>(
                    h->d_nbrU, h->d_distU, h->d_degU, slots, h->M, pos, h->upper_base[pos], L);
                LAUNCHED();
            }
            CU(cudaStreamSynchronize(s->stream));
            h->host_stale = true;  // the device copy is the current one
        }
    } else if ((rc = hnsw_ensure_host(h))) {  // (lockstep mode) the lists are edited on the host
        return rc;
    }
    if (h->in_graph[pos] && !on_device) {
        const int64_t nrows = (int64_t)rows;
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < nrows; ++v) {
            for (int lv = 0; lv <= h->node_level[v]; ++lv) {
                float* dist;
                uint16_t* deg;
                uint32_t maxc;
                uint32_t* nb = h_nbr(h, (uint32_t)v, lv, &dist, &deg, &maxc);
                if ((uint32_t)v == pos) {
                    *deg = 0;
                    continue;
                }
                uint32_t w = 0;
                for (uint32_t i = 0; i < *deg; ++i)
                    if (nb[i] != pos) {
                        nb[w] = nb[i];
                        dist[w] = dist[i];
                        ++w;
                    }
                *deg = (uint16_t)w;
            }
        }
    }
    if (h->in_graph[pos]) {
        h->in_graph[pos] = 0;
        h->n_nodes -= 1;
        if ((uint32_t)h->entry == pos) {
            h->entry = -1;
            h->entry_level = -1;
            for (uint64_t v = 0; v < rows; ++v)
                if (h->in_graph[v] && !(v < s->h_skip.size() && s->h_skip[v]) && h->node_level[v] > h->entry_level) {
                    h->entry = (int64_t)v;
                    h->entry_level = h->node_level[v];
                }
        }
    }
    if (!on_device) {
        h->host_stale = false;
        if (hnsw_device_enabled(h) && (rc = hnsw_dev_upload_all(h, false))) return rc;
    }
    if (pos < s->h_skip.size() && s->h_skip[pos]) return 0;  // deleted or emptied: it only leaves the graph
    std::vector<uint32_t> one{pos};
    return hnsw_build_locked(h, 1, true, nullptr, &one);
}

extern "C" int vecgpu_hnsw_reinsert(vecgpu_hnsw* h, int64_t rowid) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lk(h->slab->mu);
    std::lock_guard<std::mutex> lk2(h->mu);
    return hnsw_reinsert_locked(h, rowid);
    VG_CATCH
}

// Vec0Tab::insert of a row whose rowid is NOT the highest (an explicit rowid, or a re-used one): vecgpu_slab_upsert has just
// placed it between existing rows, which moved every later row one position up.  Node ids are row positions: the graph is
// renumbered in place (neighbour ids >= the new position + 1, the per-node arrays shifted by one row), the new row gets the
// level it would have had as an appended row (level_for(old row count)), its upper lists go behind the existing ones, and it
// is inserted like any other row (insert_hnsw).  Exactly ONE such insert may lie between two calls.
extern "C" int vecgpu_hnsw_insert_at(vecgpu_hnsw* h, int64_t rowid) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    vecgpu_slab* s = h->slab;
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lk2(h->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    if (h->entry < 0) return fail(VECGPU_ERR_INVALID_PARAM, "the index is empty: build it first");
    const uint64_t n_old = h->node_level.size();
    const int64_t p64 = slab_find(s, rowid, nullptr);
    if (p64 < 0) return fail(VECGPU_ERR_INVALID_PARAM, "rowid %lld is not in the slab", (long long)rowid);
    // a re-used rowid whose tombstoned row still sits in the slab: no row moved, the node is simply inserted again
    if (s->layout_gen == h->slab_gen && s->rows == n_old && !h->in_graph[(size_t)p64]) return hnsw_reinsert_locked(h, rowid);
    if (s->layout_gen != h->slab_gen + 1 || s->rows != n_old + 1)
        return fail(VECGPU_ERR_CUDA, "more than one change of row positions (or appended rows not yet indexed) since the index was last brought up to date: rebuild it");
    const uint32_t p = (uint32_t)p64;
    const bool use_dev = hnsw_device_enabled(h);
    const uint64_t slots_old = h->degU.size();
    const int L = hnsw_level_for(h, n_old);  // the level the row would have got as row number n_old (appended)
    if (!use_dev || !h->dev_valid) {
        // ---- the graph is not resident (lockstep mode, or a capacity overflow sent the last batch to the host): the same
        //      renumbering and shift on the host lists; the insertion below uploads them again where a device copy is kept
        if ((rc = hnsw_ensure_host(h))) return rc;
        for (uint64_t v = 0; v < n_old; ++v)
            for (uint32_t i = 0; i < h->deg0[v]; ++i)
                if (h->nbr0[v * h->max_m0 + i] >= p) h->nbr0[v * h->max_m0 + i] += 1;
        for (uint64_t sl = 0; sl < slots_old; ++sl)
            for (uint32_t i = 0; i < h->degU[sl]; ++i)
                if (h->nbrU[sl * h->M + i] >= p) h->nbrU[sl * h->M + i] += 1;
        h->nbr0.insert(h->nbr0.begin() + (size_t)p * h->max_m0, h->max_m0, 0u);
        h->dist0.insert(h->dist0.begin() + (size_t)p * h->max_m0, h->max_m0, 0.f);
        h->deg0.insert(h->deg0.begin() + p, (uint16_t)0);
        h->node_level.insert(h->node_level.begin() + p, (int8_t)L);
        h->in_graph.insert(h->in_graph.begin() + p, 0);
        h->upper_base.insert(h->upper_base.begin() + p, (uint32_t)slots_old);
        if (L > 0) {
            h->nbrU.resize((size_t)(slots_old + L) * h->M, 0);
            h->distU.resize((size_t)(slots_old + L) * h->M, 0.f);
            h->degU.resize(slots_old + L, 0);
        }
        if ((uint64_t)h->entry >= p) h->entry += 1;
        h->dev_valid = false;
        h->slab_gen = s->layout_gen;
        if (p < s->h_skip.size() && s->h_skip[p]) return 0;  // an empty blob: a row, but not a node
        std::vector<uint32_t> one_host{p};
        return hnsw_build_locked(h, 1, true, nullptr, &one_host);
    }
    if ((rc = hnsw_dev_flush_dirty(h))) return rc;
    // ---- shift the level-0 arrays by one row from p into the spare set, renumbering the neighbour ids on the way; the
    //      two sets swap roles (no allocation per insert once both exist)
    const bool timing = getenv("VECGPU_HNSW_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    const uint64_t n = n_old + 1;
    const size_t item_bytes[4] = {(size_t)h->max_m0 * 4, (size_t)h->max_m0 * 4, 2, 4};
    if (h->spare_rows < n) {
        const size_t cap = std::max<size_t>(n + n / 8, h->dcap_rows);
        for (int a = 0; a < 4; ++a) {
            if (h->d_spare[a]) CU(cudaFree(h->d_spare[a]));
            h->d_spare[a] = nullptr;
        }
        h->spare_rows = 0;
        for (int a = 0; a < 4; ++a) CU(cudaMalloc(&h->d_spare[a], std::max<size_t>(16, cap * item_bytes[a])));
        h->spare_rows = cap;
    }
    {
        const uint64_t items = n_old * h->max_m0;
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((items + h->max_m0 + 255) / 256, (uint64_t)s->num_sms * 32);
        const uint32_t blocks1 = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)s->num_sms * 32);
        hnsw_shift_rows_kernel<uint32_t, true><<<blocks, 256, 0, s->stream>>>(h->d_nbr0, (uint32_t*)h->d_spare[0], items, h->max_m0, p);
        LAUNCHED();
        hnsw_shift_rows_kernel<uint32_t, false><<<blocks, 256, 0, s->stream>>>((const uint32_t*)h->d_dist0, (uint32_t*)h->d_spare[1], items, h->max_m0, p);
        LAUNCHED();
        hnsw_shift_rows_kernel<uint16_t, false><<<blocks1, 256, 0, s->stream>>>(h->d_deg0, (uint16_t*)h->d_spare[2], n_old, 1, p);
        LAUNCHED();
        hnsw_shift_rows_kernel<uint32_t, false><<<blocks1, 256, 0, s->stream>>>(h->d_upper_base, (uint32_t*)h->d_spare[3], n_old, 1, p);
        LAUNCHED();
        if (slots_old) {
            hnsw_renumber_kernel<<<(uint32_t)std::min<uint64_t>((slots_old + 255) / 256, (uint64_t)s->num_sms * 16), 256, 0, s->stream>>>(
                h->d_nbrU, h->d_degU, slots_old, h->M, p);
            LAUNCHED();
        }
        std::swap(*(void**)&h->d_nbr0, h->d_spare[0]);
        std::swap(*(void**)&h->d_dist0, h->d_spare[1]);
        std::swap(*(void**)&h->d_deg0, h->d_spare[2]);
        std::swap(*(void**)&h->d_upper_base, h->d_spare[3]);
        std::swap(h->dcap_rows, h->spare_rows);
        h->dn_rows = n;
        const uint32_t ub = (uint32_t)slots_old;  // the new node's upper lists (if any) go behind the existing ones
        CU(cudaMemcpyAsync(h->d_upper_base + p, &ub, 4, cudaMemcpyHostToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    const double t_shift = now();
    // ---- host mirrors of the per-node state (the host LISTS are stale from here on: the device copy is the current one)
    h->node_level.insert(h->node_level.begin() + p, (int8_t)L);
    h->in_graph.insert(h->in_graph.begin() + p, 0);
    h->upper_base.insert(h->upper_base.begin() + p, (uint32_t)slots_old);
    h->nbr0.resize((size_t)n * h->max_m0, 0);
    h->dist0.resize((size_t)n * h->max_m0, 0.f);
    h->deg0.resize(n, 0);
    h->dirty0.assign(n, 0);
    h->host_stale = true;
    if ((uint64_t)h->entry >= p) h->entry += 1;
    // ---- the new node's upper lists go behind the existing ones
    if (L > 0) {
        const uint64_t slots = slots_old + (uint64_t)L;
        h->nbrU.resize((size_t)slots * h->M, 0);
        h->distU.resize((size_t)slots * h->M, 0.f);
        h->degU.resize(slots, 0);
        if ((rc = hnsw_dev_grow(h, n, slots_old, n, slots))) return rc;
    }
    h->slab_gen = s->layout_gen;
    if (p < s->h_skip.size() && s->h_skip[p]) return 0;  // an empty blob: a row, but not a node
    std::vector<uint32_t> one{p};
    const double t_host = now();
    rc = hnsw_build_locked(h, 1, true, nullptr, &one);
    if (timing)
        fprintf(stderr, "[vecgpu hnsw insert_at] shift+renumber %.3f ms  host mirrors %.3f ms  insert %.3f ms\n", (t_shift - t_begin) * 1e3,
                (t_host - t_shift) * 1e3, (now() - t_host) * 1e3);
    return rc;
    VG_CATCH
}

// search_hnsw (src/hnsw/search.rs:267-335) for nq queries in lockstep; distances are in the graph's (internal) metric
extern "C" int vecgpu_hnsw_search(vecgpu_hnsw* h, const void* queries, uint32_t nq, uint32_t k, uint32_t ef_search,
                                  int64_t* out_rowids, float* out_dists, uint32_t* out_counts) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    if (nq == 0 || k == 0) return 0;
    if (!queries || !out_rowids || !out_dists) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    vecgpu_slab* s = h->slab;
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lk2(h->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    for (size_t i = 0; i < (size_t)nq * k; ++i) {
        out_rowids[i] = -1;
        out_dists[i] = INFINITY;
    }
    if (out_counts) memset(out_counts, 0, (size_t)nq * 4);
    if (h->entry < 0) return 0;  // empty index: no rows (search.rs:279-281)
    if (h->slab_gen != s->layout_gen)
        return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction or out-of-order insert) since this HNSW index was built: rebuild it");
    const uint32_t ef = std::max(ef_search, k);  // search.rs:282
    const size_t qbytes = (size_t)nq * s->row_stride;
    if (qbytes > h->dq_cap) {
        if (h->d_queries) CU(cudaFree(h->d_queries));
        h->d_queries = nullptr;
        h->dq_cap = std::max(qbytes, (size_t)1 << 16);
        CU(cudaMalloc(&h->d_queries, h->dq_cap));
    }
    CU(cudaMemsetAsync(h->d_queries, 0, qbytes, s->stream));
    CU(cudaMemcpy2DAsync(h->d_queries, s->row_stride, queries, s->row_bytes, s->row_bytes, nq, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    std::vector<HQuery> qs;
    // Rows deleted (or emptied) since the build are tombstones in the slab: the walk still passes through their nodes, but
    // they are never returned (the reference removes the node and its edges, src/vtab.rs:1340-1397; the effect on the result
    // list is the same: a deleted rowid cannot come back).  The beam holds ef >= k entries, so the list is filtered first.
    const bool filter = s->n_skip != 0 && !s->h_skip.empty();
    auto emit = [&](uint32_t qi, const std::vector<HCand>& w) {
        uint32_t cnt = 0;
        for (size_t j = 0; j < w.size() && cnt < k; ++j) {
            if (filter && w[j].node < s->h_skip.size() && s->h_skip[w[j].node]) continue;
            out_rowids[(size_t)qi * k + cnt] = h_rowid_of(s, w[j].node);
            out_dists[(size_t)qi * k + cnt] = w[j].d;
            ++cnt;
        }
        if (out_counts) out_counts[qi] = cnt;
    };
    // lockstep driver (one scoring launch per expansion round) for the queries listed in `which` (NULL: q0..q0+m)
    auto host_loop = [&](const uint32_t* which, uint32_t q0, uint32_t m) -> int {
        int r0 = hnsw_ensure_host(h);  // the lockstep driver walks the host lists
        if (r0) return r0;
        qs.clear();
        qs.resize(m);
        for (uint32_t i = 0; i < m; ++i) hq_init(h, qs[i], which ? which[i] : q0 + i, -1, ef);
        int r = hnsw_run_batch(h, qs, m, (const uint8_t*)h->d_queries);
        if (r) return r;
        for (uint32_t i = 0; i < m; ++i)
            if (!qs[i].layers.empty()) emit(which ? which[i] : q0 + i, qs[i].layers[0]);
        return 0;
    };
    // a beam whose sorted array (2 ef keys) does not fit one warp's share of shared memory goes through the lockstep driver
    const bool ef_fits = ((size_t)2 * ef + 64) * 8 + ((h->max_m0 + 31u) & ~31u) * 4 + s->row_stride <= 200 * 1024;
    if (hnsw_device_enabled(h) && ef_fits) {
        // K6: the whole layered search on the device, one warp per query
        const uint32_t chunk = 1u << 16;
        std::vector<uint32_t> fb;
        std::vector<HCand> w;
        for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
            const uint32_t m = std::min(chunk, nq - q0);
            HDevOut o;
            rc = hnsw_dev_search(h, (const uint8_t*)h->d_queries + (size_t)q0 * s->row_stride, nullptr, nullptr, m, ef, filter ? ef : k, nullptr, m, &o);
            if (rc) return rc;
            for (uint32_t i = 0; i < m; ++i) {
                if (o.status[i]) {
                    fb.push_back(q0 + i);
                    continue;
                }
                w.resize(o.cnt[i]);
                for (uint32_t j = 0; j < o.cnt[i]; ++j) w[j] = h_decode(o.keys[(size_t)i * o.take + j]);
                emit(q0 + i, w);
            }
        }
        if (!fb.empty()) {
            h->dev_fallbacks += fb.size();
            if ((rc = host_loop(fb.data(), 0, (uint32_t)fb.size()))) return rc;
        }
        return 0;
    }
    const uint32_t chunk = 8192;
    for (uint32_t q0 = 0; q0 < nq; q0 += chunk)
        if ((rc = host_loop(nullptr, q0, std::min(chunk, nq - q0)))) return rc;
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_hnsw_stats(vecgpu_hnsw* h, uint64_t* nodes, uint64_t* edges, int32_t* entry_level, uint64_t* distances_scored,
                                 uint64_t* rounds) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lks(h->slab->mu);  // same order as build and search: slab, then index
    std::lock_guard<std::mutex> lk(h->mu);
    if (nodes) *nodes = h->n_nodes;
    if (edges && h->host_stale) {
        int rc = use_device(h->slab->device);
        if (rc || (rc = hnsw_ensure_host(h))) return rc;
    }
    if (edges) {
        uint64_t e = 0;
        for (uint16_t d : h->deg0) e += d;
        for (uint16_t d : h->degU) e += d;
        *edges = e;
    }
    if (entry_level) *entry_level = h->entry_level;
    if (distances_scored) *distances_scored = h->scored;
    if (rounds) *rounds = h->rounds;
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_hnsw_entry_point(vecgpu_hnsw* h, int64_t* rowid, int32_t* level) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lks(h->slab->mu);  // same order as build and search: slab, then index
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->entry >= 0 && h->slab_gen != h->slab->layout_gen)
        return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction or out-of-order insert) since this HNSW index was built: rebuild it");
    if (rowid) *rowid = h->entry < 0 ? -1 : h_rowid_of(h->slab, (uint32_t)h->entry);
    if (level) *level = h->entry < 0 ? -1 : h->entry_level;
    return 0;
    VG_CATCH
}

// Node list for the bulk write-back into {t}_{c}_hnsw_nodes(rowid, level, vector) (src/shadow.rs:464-474): every indexed
// row with its level.  Call with cap = 0 to get the count in *n_out.
extern "C" int vecgpu_hnsw_export_nodes(vecgpu_hnsw* h, uint64_t cap, int64_t* rowids, int32_t* levels, uint64_t* n_out) {
    VG_TRY
    if (!h || !n_out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lks(h->slab->mu);  // same order as build and search: slab, then index
    std::lock_guard<std::mutex> lk(h->mu);
    vecgpu_slab* s = h->slab;
    if (h->entry >= 0 && h->slab_gen != s->layout_gen)
        return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction or out-of-order insert) since this HNSW index was built: rebuild it");
    uint64_t n = 0;
    if (h->entry >= 0) {
        const uint64_t rows = h->node_level.size();
        for (uint64_t pos = 0; pos < rows; ++pos) {
            if (!h->in_graph[pos]) continue;
            if (pos < s->h_skip.size() && s->h_skip[pos]) continue;  // deleted since: Vec0Tab::delete removed the node row (vtab.rs:1340-1397)
            if (n < cap) {
                rowids[n] = h_rowid_of(s, (uint32_t)pos);
                levels[n] = h->node_level[pos];
            }
            ++n;
        }
    }
    *n_out = n;
    return 0;
    VG_CATCH
}

// K6 counters: queries answered by the device search kernel, how many of them overflowed a device capacity and were
// re-run through the lockstep driver, and the number of search launches.
// The reference's expansion batch-size histogram (BATCH_SIZE_1_4 ... BATCH_SIZE_65_PLUS, src/hnsw/search.rs:73-85, 443-455):
// expansions of the device walks (build and queries) by the number of unvisited neighbours they scored.
extern "C" int vecgpu_hnsw_batch_histogram(vecgpu_hnsw* h, uint64_t out5[5]) {
    VG_TRY
    if (!h || !out5) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    for (int b = 0; b < 5; ++b) out5[b] = h->batch_hist[b];
    return 0;
    VG_CATCH
}

extern "C" int vecgpu_hnsw_device_stats(vecgpu_hnsw* h, uint64_t* queries, uint64_t* fallbacks, uint64_t* launches) {
    VG_TRY
    if (!h) return fail(VECGPU_ERR_INVALID_PARAM, "hnsw is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    if (queries) *queries = h->dev_queries;
    if (fallbacks) *fallbacks = h->dev_fallbacks;
    if (launches) *launches = h->dev_launches;
    return 0;
    VG_CATCH
}

// edge list for a bulk write-back into {t}_{c}_hnsw_edges(from_rowid, to_rowid, level, distance) (src/shadow.rs:478-487).
// Call with cap = 0 to get the count in *n_out.
extern "C" int vecgpu_hnsw_export_edges(vecgpu_hnsw* h, uint64_t cap, int64_t* from_rowids, int64_t* to_rowids, int32_t* levels,
                                        float* dists, uint64_t* n_out) {
    VG_TRY
    if (!h || !n_out) return fail(VECGPU_ERR_INVALID_PARAM, "NULL argument");
    std::lock_guard<std::mutex> lks(h->slab->mu);  // same order as build and search: slab, then index
    std::lock_guard<std::mutex> lk(h->mu);
    vecgpu_slab* s = h->slab;
    if (h->entry >= 0 && h->slab_gen != s->layout_gen)
        return fail(VECGPU_ERR_CUDA, "the slab's rows moved (compaction or out-of-order insert) since this HNSW index was built: rebuild it");
    if (h->host_stale) {
        int rc = use_device(s->device);
        if (rc || (rc = hnsw_ensure_host(h))) return rc;
    }
    uint64_t n = 0;
    const uint64_t rows = h->node_level.size();
    auto gone = [&](uint64_t pos) { return pos < s->h_skip.size() && s->h_skip[pos] != 0; };  // deleted since the build
    for (uint64_t node = 0; node < rows; ++node) {
        if (gone(node)) continue;  // Vec0Tab::delete removed its edges in both directions (vtab.rs:1340-1397)
        for (int lv = 0; lv <= h->node_level[node]; ++lv) {
            float* dist;
            uint16_t* deg;
            uint32_t maxc;
            const uint32_t* nb = h_nbr(h, (uint32_t)node, lv, &dist, &deg, &maxc);
            for (uint32_t i = 0; i < *deg; ++i) {
                if (gone(nb[i])) continue;
                if (n < cap) {
                    from_rowids[n] = h_rowid_of(s, (uint32_t)node);
                    to_rowids[n] = h_rowid_of(s, nb[i]);
                    levels[n] = lv;
                    dists[n] = dist[i];
                }
                ++n;
            }
        }
    }
    *n_out = n;
    return 0;
    VG_CATCH
}
