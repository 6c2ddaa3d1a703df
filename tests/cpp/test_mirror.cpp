// C++ host mirror (include/vecgpu.hpp) exercised with the reference's own unit tests, restated.
// Each block cites the Rust test it mirrors.  Exit code 0 = all passed.  On a host without a CUDA device the
// program checks the loud-failure path instead (there is no CPU fallback) and prints NO_DEVICE_OK.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "vecgpu.hpp"

using namespace vecgpu;
static int g_fail = 0, g_pass = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (cond) ++g_pass;                                                         \
        else { ++g_fail; std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static std::vector<uint8_t> blob_f32(std::initializer_list<float> v) { return Vector::from_f32(std::vector<float>(v)).as_bytes(); }

int main() {
    // ---- host-only logic (runs everywhere)
    CHECK(distance_metric_from_str("L2").unwrap() == DistanceMetric::L2);            // src/distance/mod.rs:26-34
    CHECK(distance_metric_from_str("euclidean").unwrap() == DistanceMetric::L2);
    CHECK(distance_metric_from_str("manhattan").unwrap() == DistanceMetric::L1);
    CHECK(distance_metric_from_str("dot").is_err());
    CHECK(std::string(as_str(DistanceMetric::Cosine)) == "cosine");
    CHECK(vector_type_from_str("float").unwrap() == VectorType::Float32);              // src/vector.rs:30-37
    CHECK(vector_type_from_str("binary").unwrap() == VectorType::Bit);
    CHECK(vector_type_from_str("float16").is_err());
    CHECK(row_bytes(VectorType::Float32, 768) == 3072 && row_bytes(VectorType::Bit, 13) == 2);
    {   // src/distance/mod.rs:155-162 test_dimension_mismatch: checked before anything touches the device
        auto r = distance(Vector::from_f32({1, 2, 3}), Vector::from_f32({1, 2}), DistanceMetric::L2);
        CHECK(r.is_err() && r.unwrap_err().kind == Error::Kind::DimensionMismatch && r.unwrap_err().expected == 3 && r.unwrap_err().actual == 2);
        auto t = distance(Vector::from_f32({1, 2}), Vector::from_i8({1, 2}), DistanceMetric::L2);
        CHECK(t.is_err() && t.unwrap_err().kind == Error::Kind::InvalidVectorType);
    }
    CHECK(internal_distance_metric(DistanceMetric::Cosine, true) == DistanceMetric::L2);  // src/hnsw/mod.rs:129-137
    CHECK(convert_distance_for_output(DistanceMetric::Cosine, true, 1.0f) == 0.5f);       // :139-146

    if (vecgpu_device_count() <= 0) {
        auto s = Slab::create(VectorType::Float32, 4);
        CHECK(s.is_err() && s.unwrap_err().kind == Error::Kind::InvalidState);
        auto d = distance(Vector::from_f32({1, 2}), Vector::from_f32({3, 4}), DistanceMetric::L2);
        CHECK(d.is_err() && d.unwrap_err().kind == Error::Kind::InvalidState);
        if (g_fail) std::printf("FAILED %d checks\n", g_fail);
        else std::printf("NO_DEVICE_OK %d checks\n", g_pass);
        return g_fail ? 1 : 0;
    }

    // ---- src/distance/scalar.rs:114-213
    {
        auto a = Vector::from_f32({1, 2, 3}), b = Vector::from_f32({4, 5, 6});
        auto l2 = distance(a, b, DistanceMetric::L2);                                  // test_scalar_l2_f32
        CHECK(l2.is_ok() && std::fabs(l2.unwrap() - 5.196f) < 0.01f);
        auto l1 = distance(a, b, DistanceMetric::L1);                                  // test_scalar_l1_f32
        CHECK(l1.is_ok() && std::fabs(l1.unwrap() - 9.0f) < 0.01f);
        auto co = distance(Vector::from_f32({1, 0, 0}), Vector::from_f32({0, 1, 0}), DistanceMetric::Cosine);  // test_scalar_cosine_f32
        CHECK(co.is_ok() && std::fabs(co.unwrap() - 1.0f) < 0.01f);
        auto cp = distance(a, Vector::from_f32({2, 4, 6}), DistanceMetric::Cosine);   // test_scalar_cosine_parallel
        CHECK(cp.is_ok() && std::fabs(cp.unwrap()) < 0.01f);
        auto i2 = distance(Vector::from_i8({1, 2, 3}), Vector::from_i8({4, 5, 6}), DistanceMetric::L2);  // test_scalar_l2_i8
        CHECK(i2.is_ok() && std::fabs(i2.unwrap() - 5.196f) < 0.01f);
        auto i1 = distance(Vector::from_i8({1, 2, 3}), Vector::from_i8({4, 5, 6}), DistanceMetric::L1);  // test_scalar_l1_i8
        CHECK(i1.is_ok() && std::fabs(i1.unwrap() - 9.0f) < 0.01f);
        const uint8_t x[4] = {1, 0, 1, 0}, y[4] = {0, 1, 1, 0};                        // test_scalar_hamming
        auto hm = distance(VectorRef::from_blob(x, 4, VectorType::Bit, 32), VectorRef::from_blob(y, 4, VectorType::Bit, 32), DistanceMetric::Hamming);
        CHECK(hm.is_ok() && hm.unwrap() >= 0.0f && hm.unwrap() == 2.0f);
        auto bad = distance(a, b, DistanceMetric::Hamming);                           // src/distance/mod.rs:78-82
        CHECK(bad.is_err() && bad.unwrap_err().kind == Error::Kind::InvalidDistanceMetric &&
              bad.unwrap_err().message == "Distance metric Hamming not supported for vector type Float32");
    }
    // ---- src/vector.rs:746-789
    {
        auto n = Vector::from_f32({3, 4}).normalize();
        CHECK(n.is_ok());
        auto v = n.unwrap().as_f32().unwrap();
        CHECK(std::fabs(v[0] - 0.6f) < 1e-4f && std::fabs(v[1] - 0.8f) < 1e-4f);
        CHECK(Vector::from_f32({0, 0}).normalize().is_err());                          // "Cannot normalize zero vector"
        CHECK(Vector::from_i8({1, 2}).normalize().is_err());
        auto q = Vector::from_f32({0.0f, 0.5f, 1.0f}).quantize_int8().unwrap().as_i8().unwrap();
        CHECK(q[0] == -128 && q[2] == 127 && q[0] < q[1] && q[1] < q[2]);
        auto qi = Vector::from_f32({1.0f, -1.0f, 0.5f, 2.0f}).quantize_int8_for_index().unwrap().as_i8().unwrap();
        CHECK(qi[0] == 127 && qi[1] == -127 && qi[2] == 64 && qi[3] == 127);           // round(0.5*127)=64 (half away from zero)
        auto qb = Vector::from_f32({1, -1, 1, -1, 5, -5, 0.5f, -0.5f, 9}).quantize_binary().unwrap();
        CHECK(qb.vec_type() == VectorType::Bit && qb.as_bytes().size() == 2);
    }
    // ---- tests/test_knn_simple.rs:34-53: e1, e2, e3; query e1; k = 2; default metric cosine
    {
        auto slab = Slab::create(VectorType::Float32, 3).unwrap().release();
        const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        CHECK(slab->load(nullptr, eye, 3).is_ok());
        auto res = brute_force_search(*slab, blob_f32({1, 0, 0}), 2, DistanceMetric::Cosine);
        CHECK(res.is_ok() && res.unwrap().size() == 2 && res.unwrap()[0].first == 1);
        CHECK(res.unwrap()[1].first == 2 && res.unwrap()[0].second == 0.0f && res.unwrap()[1].second == 1.0f);  // tie 2 vs 3 -> stable -> 2
        CHECK(brute_force_search(*slab, blob_f32({1, 0, 0}), 0, DistanceMetric::Cosine).unwrap().empty());
        CHECK(brute_force_search(*slab, blob_f32({1, 0, 0}), 100, DistanceMetric::Cosine).unwrap().size() == 3);
        CHECK(brute_force_search(*slab, blob_f32({1, 0}), 2, DistanceMetric::Cosine).unwrap().empty());
        delete slab;
    }
    // ---- tests/integration_test.rs:635-678: rows [i, i+1, i+2], inserted one by one; query [1,2,3]; k = 3
    {
        auto slab = Slab::create(VectorType::Float32, 3).unwrap().release();
        for (int i = 1; i <= 5; ++i) CHECK(slab->upsert(i, blob_f32({(float)i, (float)i + 1, (float)i + 2})).is_ok());
        auto res = brute_force_search(*slab, blob_f32({1, 2, 3}), 3, DistanceMetric::Cosine).unwrap();
        CHECK(res.size() == 3 && res[0].first == 1 && res[0].second < 0.01f);
        auto l2 = brute_force_search(*slab, blob_f32({1, 2, 3}), 3, DistanceMetric::L2).unwrap();
        CHECK(l2.size() == 3 && l2[0].first == 1 && l2[1].first == 2 && l2[2].first == 3);
        CHECK(slab->remove(1).is_ok() && slab->upsert(2, {}).is_ok());               // delete; empty blob => skipped row
        auto after = brute_force_search(*slab, blob_f32({1, 2, 3}), 3, DistanceMetric::L2).unwrap();
        CHECK(after.size() == 3 && after[0].first == 3);
        auto sc = slab->score(blob_f32({1, 2, 3}), {3, 1, 99}, DistanceMetric::L2).unwrap();  // src/hnsw/search.rs:501-513
        CHECK(slab->compact().unwrap() == 2);                                           // rows 1 (deleted) and 2 (empty blob) are gone
        auto compacted = brute_force_search(*slab, blob_f32({1, 2, 3}), 3, DistanceMetric::L2).unwrap();
        CHECK(compacted.size() == 3 && compacted[0].first == 3 && compacted[0].second == after[0].second);
        CHECK(std::fabs(sc[0] - std::sqrt(12.0f)) < 1e-6f && std::isnan(sc[1]) && std::isnan(sc[2]));
        delete slab;
    }
    // ---- src/vtab.rs:3246-3286: HNSW cosine on 3 rows -> rowids [1, 2] in order
    {
        auto slab = Slab::create(VectorType::Float32, 3).unwrap().release();
        const float rows[9] = {1, 0, 0, 0.8f, 0.6f, 0, 0, 0, 1};  // already unit length (stored representation)
        CHECK(slab->load(nullptr, rows, 3).is_ok());
        auto idx = HnswIndex::create(*slab, DistanceMetric::Cosine).unwrap().release();
        CHECK(idx->rebuild().is_ok());
        auto res = idx->search(blob_f32({1, 0, 0}), 2).unwrap();
        CHECK(res.size() == 2 && res[0].first == 1 && res[1].first == 2 && res[0].second < 1e-6f);
        CHECK(std::fabs(res[1].second - 0.2f) < 1e-5f);  // cosine distance of (1,0,0) and (0.8,0.6,0) via d_L2^2/2
        auto ep = idx->entry_point().unwrap();           // -> _hnsw_meta entry_point_rowid / _level (src/hnsw/mod.rs:97-103)
        CHECK(ep.first >= 1 && ep.first <= 3 && ep.second >= 0);
        CHECK(HnswIndex::create(*slab, DistanceMetric::Cosine, 1, 400).is_err());  // M in [2,100] (src/sql_functions.rs:442-469)
        // Vec0Tab::insert -> insert_hnsw: a row that arrives in rowid order joins the graph without a rebuild
        CHECK(slab->upsert(4, blob_f32({0.6f, 0.8f, 0})).is_ok());
        CHECK(idx->insert_appended().unwrap() == 1 && idx->insert_appended().unwrap() == 0);
        auto res4 = idx->search(blob_f32({0.6f, 0.8f, 0}), 2).unwrap();
        CHECK(res4.size() == 2 && res4[0].first == 4 && res4[1].first == 2);
        // Vec0Tab::update: the row gets a new vector, its node is deleted and inserted again
        CHECK(slab->upsert(3, blob_f32({0, 1, 0})).is_ok() && idx->reinsert(3).is_ok());
        auto res3 = idx->search(blob_f32({0, 1, 0}), 2).unwrap();
        CHECK(res3.size() == 2 && res3[0].first == 3 && res3[0].second < 1e-6f && res3[1].first == 4);
        CHECK(idx->reinsert(77).is_err());
        // Vec0Tab::insert with an explicit rowid BETWEEN existing ones: row positions move, the resident graph is renumbered
        CHECK(slab->upsert(10, blob_f32({0.8f, 0, 0.6f})).is_ok() && idx->insert_appended().unwrap() == 1);
        CHECK(slab->upsert(7, blob_f32({0, 0.6f, 0.8f})).is_ok() && idx->insert_at(7).is_ok());
        auto res7 = idx->search(blob_f32({0, 0.6f, 0.8f}), 1).unwrap();
        CHECK(res7.size() == 1 && res7[0].first == 7 && res7[0].second < 1e-6f);
        auto res10 = idx->search(blob_f32({0.8f, 0, 0.6f}), 1).unwrap();
        CHECK(res10.size() == 1 && res10[0].first == 10);
        CHECK(idx->insert_at(8).is_err());  // not in the slab
        delete idx;
        delete slab;
    }
    // ---- stored representation (src/hnsw/insert.rs:300-322) + index_quantization=int8 (src/vector.rs:554-575)
    {
        auto col = Slab::create(VectorType::Float32, 3).unwrap().release();
        const float rows[12] = {2, 0, 0, 4, 3, 0, 0, 0, 5, 0, 0, 0};  // raw column values; row 4 is a zero vector
        CHECK(col->load(nullptr, rows, 4).is_ok());
        auto none = col->stored_for_hnsw(false, false);
        CHECK(none.is_ok() && none.unwrap() == nullptr);
        auto st = col->stored_for_hnsw(true, false).unwrap().release();  // normalised; the zero vector cannot be stored
        CHECK(st && st->vec_type == VectorType::Float32 && st->live_rows() == 3);
        auto idx = HnswIndex::create(*st, DistanceMetric::Cosine).unwrap().release();
        CHECK(idx->rebuild().is_ok());
        auto res = idx->search(blob_f32({1, 0, 0}), 2).unwrap();
        CHECK(res.size() == 2 && res[0].first == 1 && res[1].first == 2 && std::fabs(res[1].second - 0.2f) < 1e-5f);
        auto hist = idx->batch_histogram().unwrap();
        CHECK(hist[0] + hist[1] + hist[2] + hist[3] + hist[4] > 0 && hist[4] == 0);
        auto q8 = col->stored_for_hnsw(true, true).unwrap().release();
        CHECK(q8 && q8->vec_type == VectorType::Int8 && q8->row_bytes() == 3);
        auto sc = q8->score({127, 0, 0}, {1, 2}, DistanceMetric::L2).unwrap();       // (127,0,0) and round((.8,.6,0)*127) = (102,76,0)
        CHECK(sc[0] == 0.0f && std::fabs(sc[1] - std::sqrt(25.0f * 25.0f + 76.0f * 76.0f)) < 1e-4f);
        delete idx;
        delete q8;
        delete st;
        delete col;
    }
    // ---- a column sharded by rowid range (two shards on device 0): same answers as one slab
    {
        auto sh = ShardedSlab::create(VectorType::Float32, 3, {0, 0}).unwrap().release();
        auto one = Slab::create(VectorType::Float32, 3).unwrap().release();
        std::vector<float> rows;
        for (int i = 1; i <= 40; ++i) { rows.push_back((float)(i % 7)); rows.push_back((float)(i % 5) + 1); rows.push_back((float)i * 0.25f); }
        CHECK(sh->load(nullptr, rows.data(), 40).is_ok() && one->load(nullptr, rows.data(), 40).is_ok());
        CHECK(sh->num_shards() == 2 && sh->live_rows() == 40);
        CHECK(sh->remove(7).is_ok() && one->remove(7).is_ok());
        CHECK(sh->upsert(33, blob_f32({1, 2, 3})).is_ok() && one->upsert(33, blob_f32({1, 2, 3})).is_ok());
        for (auto m : {DistanceMetric::L2, DistanceMetric::Cosine, DistanceMetric::L1}) {
            auto a = sh->brute_force_search(blob_f32({1, 2, 3}), 12, m).unwrap();
            auto b = brute_force_search(*one, blob_f32({1, 2, 3}), 12, m).unwrap();
            CHECK(a.size() == 12 && a == b);
        }
        CHECK(sh->brute_force_search(blob_f32({1, 2}), 5, DistanceMetric::L2).unwrap().empty());
        delete one;
        delete sh;
    }
    if (g_fail) std::printf("FAILED %d checks (%d passed)\n", g_fail, g_pass);
    else std::printf("ALL_PASSED %d checks\n", g_pass);
    return g_fail ? 1 : 0;
}
