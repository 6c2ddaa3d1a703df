// vecgpu.hpp — header-only C++17 mirror of the reference's interface for the distance-scoring path,
// layered on the C ABI of vecgpu.h.  The reference is Rust; no cargo/rustc exists in the build image, so this
// is the compiled-language host side: same names, argument meaning and error behaviour as the crate, so tests
// read like the reference's own (tests/cpp/test_mirror.cpp).  Citations are into the reference tree.
//
//   vecgpu::VectorType / DistanceMetric       src/vector.rs:9-46, src/distance/mod.rs:12-44
//   vecgpu::Error, Result<T>                  src/error.rs:5-38
//   vecgpu::Vector, VectorRef                 src/vector.rs:126-320, 444-608 (producers run on the GPU)
//   vecgpu::distance(a, b, metric)            src/distance/mod.rs:52-84
//   vecgpu::Slab                              the HBM copy of {t}_data.vecNN (src/shadow.rs:111-129)
//   vecgpu::brute_force_search(...)           src/vtab.rs:2573-2623
//   vecgpu::HnswIndex                         src/hnsw/{insert,search,rebuild}.rs, src/hnsw/mod.rs:129-146
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <variant>
#include <array>
#include <vector>

#include "vecgpu.h"

namespace vecgpu {

// ---- src/error.rs:5-36 ------------------------------------------------------------------------------------
struct Error {
    enum class Kind { InvalidVectorFormat, DimensionMismatch, InvalidVectorType, InvalidDistanceMetric, NotImplemented,
                      InvalidParameter, InvalidState };
    Kind kind;
    std::string message;
    size_t expected = 0, actual = 0;  // DimensionMismatch { expected, actual }
    static Error from_status(int code) {
        const std::string msg = vecgpu_last_error();
        switch (code) {
            case VECGPU_ERR_DIM_MISMATCH: return {Kind::DimensionMismatch, msg};
            case VECGPU_ERR_UNSUPPORTED:
                return {msg.rfind("invalid vector type", 0) == 0 ? Kind::InvalidVectorType : Kind::InvalidDistanceMetric, msg};
            case VECGPU_ERR_INVALID_PARAM: return {Kind::InvalidParameter, msg};
            default: return {Kind::InvalidState, msg};  // no device / OOM / launch failure: there is no CPU fallback
        }
    }
};

template <class T>
class Result {  // std::result::Result<T, Error>
    std::variant<T, Error> v_;

public:
    Result(T value) : v_(std::move(value)) {}
    Result(Error e) : v_(std::move(e)) {}
    bool is_ok() const { return v_.index() == 0; }
    bool is_err() const { return !is_ok(); }
    T& unwrap() { return std::get<0>(v_); }
    const T& unwrap() const { return std::get<0>(v_); }
    const Error& unwrap_err() const { return std::get<1>(v_); }
};
struct Unit {};

// ---- enums ------------------------------------------------------------------------------------------------
enum class VectorType { Float32 = VECGPU_F32, Int8 = VECGPU_I8, Bit = VECGPU_BIT };          // src/vector.rs:9-16
enum class DistanceMetric { L2 = VECGPU_L2, L1 = VECGPU_L1, Cosine = VECGPU_COSINE, Hamming = VECGPU_HAMMING };  // distance/mod.rs:12-21

inline std::string to_lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}
inline Result<VectorType> vector_type_from_str(const std::string& s) {  // src/vector.rs:30-37
    const std::string t = to_lower(s);
    if (t == "float32" || t == "float") return VectorType::Float32;
    if (t == "int8") return VectorType::Int8;
    if (t == "bit" || t == "binary") return VectorType::Bit;
    return Error{Error::Kind::InvalidVectorType, s};
}
inline const char* as_str(VectorType t) { return t == VectorType::Float32 ? "float32" : t == VectorType::Int8 ? "int8" : "bit"; }
inline Result<DistanceMetric> distance_metric_from_str(const std::string& s) {  // src/distance/mod.rs:26-34
    const std::string t = to_lower(s);
    if (t == "l2" || t == "euclidean") return DistanceMetric::L2;
    if (t == "l1" || t == "manhattan") return DistanceMetric::L1;
    if (t == "cosine") return DistanceMetric::Cosine;
    if (t == "hamming") return DistanceMetric::Hamming;
    return Error{Error::Kind::InvalidDistanceMetric, s};
}
inline const char* as_str(DistanceMetric m) {  // src/distance/mod.rs:37-44
    switch (m) {
        case DistanceMetric::L2: return "l2";
        case DistanceMetric::L1: return "l1";
        case DistanceMetric::Cosine: return "cosine";
        default: return "hamming";
    }
}
inline size_t row_bytes(VectorType t, size_t dims) { return vecgpu_row_bytes((int)t, (uint32_t)dims); }

// ---- VectorData / VectorRef / Vector (src/vector.rs:90-320) --------------------------------------------------
struct VectorRef {  // zero-copy borrowed view (src/vector.rs:126-184)
    VectorType vec_type;
    size_t dimensions;
    const uint8_t* data;
    size_t len;
    static VectorRef from_blob(const uint8_t* blob, size_t len, VectorType t, size_t dims) { return {t, dims, blob, len}; }
    const uint8_t* as_bytes() const { return data; }
};

class Vector {  // owned (src/vector.rs:215-320)
public:
    VectorType vec_type_;
    size_t dimensions_;
    std::vector<uint8_t> data_;

    static Vector from_f32(const std::vector<float>& v) {  // :217-227
        Vector out{VectorType::Float32, v.size(), std::vector<uint8_t>(v.size() * 4)};
        if (!v.empty()) std::memcpy(out.data_.data(), v.data(), v.size() * 4);
        return out;
    }
    static Vector from_i8(const std::vector<int8_t>& v) {  // :230-236
        Vector out{VectorType::Int8, v.size(), std::vector<uint8_t>(v.size())};
        if (!v.empty()) std::memcpy(out.data_.data(), v.data(), v.size());
        return out;
    }
    static Result<Vector> from_blob(const uint8_t* blob, size_t len, VectorType t, size_t dims) {  // :259-266
        return Vector{t, dims, std::vector<uint8_t>(blob, blob + len)};
    }
    VectorType vec_type() const { return vec_type_; }
    size_t dimensions() const { return dimensions_; }
    const std::vector<uint8_t>& as_bytes() const { return data_; }
    VectorRef as_ref() const { return {vec_type_, dimensions_, data_.data(), data_.size()}; }
    Result<std::vector<float>> as_f32() const {  // :304-320
        if (vec_type_ != VectorType::Float32) return Error{Error::Kind::InvalidVectorType, "Vector is not Float32 type"};
        std::vector<float> out(dimensions_);
        if (dimensions_) std::memcpy(out.data(), data_.data(), dimensions_ * 4);
        return out;
    }
    Result<std::vector<int8_t>> as_i8() const {
        if (vec_type_ != VectorType::Int8) return Error{Error::Kind::InvalidVectorType, "Vector is not Int8 type"};
        std::vector<int8_t> out(dimensions_);
        if (dimensions_) std::memcpy(out.data(), data_.data(), dimensions_);
        return out;
    }
    // producers, on the GPU (src/vector.rs:444-608)
    Result<Vector> normalize(int device = 0) const {
        if (vec_type_ == VectorType::Int8) return Error{Error::Kind::InvalidVectorType, "Cannot normalize Int8 vectors (would lose precision)"};
        if (vec_type_ == VectorType::Bit) return Error{Error::Kind::InvalidVectorType, "Cannot normalize binary vectors"};
        Vector out{VectorType::Float32, dimensions_, std::vector<uint8_t>(data_.size())};
        int rc = vecgpu_normalize_f32((const float*)data_.data(), 1, (uint32_t)dimensions_, device, (float*)out.data_.data());
        if (rc) return Error::from_status(rc);
        return out;
    }
    Result<Vector> quantize_int8(int device = 0) const {
        if (vec_type_ != VectorType::Float32) return Error{Error::Kind::InvalidVectorType, "Can only quantize Float32 vectors"};
        Vector out{VectorType::Int8, dimensions_, std::vector<uint8_t>(dimensions_)};
        int rc = vecgpu_quantize_int8((const float*)data_.data(), 1, (uint32_t)dimensions_, device, (int8_t*)out.data_.data());
        if (rc) return Error::from_status(rc);
        return out;
    }
    Result<Vector> quantize_int8_for_index(int device = 0) const {
        if (vec_type_ != VectorType::Float32) return Error{Error::Kind::InvalidVectorType, "Can only quantize Float32 vectors"};
        Vector out{VectorType::Int8, dimensions_, std::vector<uint8_t>(dimensions_)};
        int rc = vecgpu_quantize_int8_for_index((const float*)data_.data(), 1, (uint32_t)dimensions_, device, (int8_t*)out.data_.data());
        if (rc) return Error::from_status(rc);
        return out;
    }
    Result<Vector> quantize_binary(int device = 0) const {
        if (vec_type_ != VectorType::Float32) return Error{Error::Kind::InvalidVectorType, "Can only quantize Float32 vectors to binary"};
        Vector out{VectorType::Bit, dimensions_, std::vector<uint8_t>((dimensions_ + 7) / 8)};
        int rc = vecgpu_quantize_binary((const float*)data_.data(), 1, (uint32_t)dimensions_, device, out.data_.data());
        if (rc) return Error::from_status(rc);
        return out;
    }
};

// ---- distance() (src/distance/mod.rs:52-84): dimensions, then types, then the (type, metric) match -------------
inline Result<float> distance(const VectorRef& a, const VectorRef& b, DistanceMetric metric, int device = 0) {
    if (a.dimensions != b.dimensions) {
        Error e{Error::Kind::DimensionMismatch,
                "Dimension mismatch: expected " + std::to_string(a.dimensions) + ", got " + std::to_string(b.dimensions)};
        e.expected = a.dimensions;
        e.actual = b.dimensions;
        return e;
    }
    if (a.vec_type != b.vec_type) return Error{Error::Kind::InvalidVectorType, "Vector types must match for distance calculation"};
    const size_t rb = row_bytes(a.vec_type, a.dimensions);
    if (vecgpu_metric_supported((int)a.vec_type, (int)metric) && (a.len != rb || b.len != rb))
        return Error{Error::Kind::InvalidParameter, "distance calculation failed"};  // simsimd returns None (scalar.rs:18)
    float out = 0.f;
    int rc = vecgpu_distance_pairs((int)a.vec_type, (uint32_t)a.dimensions, (uint32_t)b.dimensions, a.data, b.data, 1, (int)metric, device, &out);
    if (rc) return Error::from_status(rc);
    return out;
}
inline Result<float> distance(const Vector& a, const Vector& b, DistanceMetric metric, int device = 0) {
    return distance(a.as_ref(), b.as_ref(), metric, device);
}

// ---- src/hnsw/mod.rs:129-146 ---------------------------------------------------------------------------------
inline DistanceMetric internal_distance_metric(DistanceMetric m, bool normalize_vectors) {
    return (m == DistanceMetric::Cosine && normalize_vectors) ? DistanceMetric::L2 : m;
}
inline float convert_distance_for_output(DistanceMetric m, bool normalize_vectors, float internal_dist) {
    return (m == DistanceMetric::Cosine && normalize_vectors) ? (internal_dist * internal_dist) / 2.0f : internal_dist;
}

// ---- Slab -------------------------------------------------------------------------------------------------
class Slab {
    vecgpu_slab* h_ = nullptr;

public:
    VectorType vec_type;
    size_t dims;
    Slab(VectorType t, size_t d) : vec_type(t), dims(d) {}
    Slab(const Slab&) = delete;
    Slab& operator=(const Slab&) = delete;
    ~Slab() { if (h_) vecgpu_slab_destroy(h_); }
    static Result<std::unique_ptr<Slab>> create(VectorType t, size_t dims, uint64_t capacity_hint = 0, int device = 0) {
        auto s = std::make_unique<Slab>(t, dims);
        int rc = vecgpu_slab_create((int)t, (uint32_t)dims, capacity_hint, device, &s->h_);
        if (rc) return Error::from_status(rc);
        return s;
    }
    vecgpu_slab* raw() { return h_; }
    size_t row_bytes() const { return vecgpu::row_bytes(vec_type, dims); }
    // The STORED representation of this column for an HNSW index (src/hnsw/insert.rs:300-322): normalised for cosine
    // columns, quantize_int8_for_index'ed with index_quantization=int8.  nullptr when neither applies (this slab is it).
    Result<std::unique_ptr<Slab>> stored_for_hnsw(bool normalize_vectors, bool int8_quantization) {
        vecgpu_slab* out = nullptr;
        int rc = vecgpu_hnsw_stored_slab(h_, normalize_vectors ? 1 : 0, int8_quantization ? 1 : 0, &out);
        if (rc) return Error::from_status(rc);
        if (!out) return std::unique_ptr<Slab>();
        auto s = std::make_unique<Slab>(int8_quantization ? VectorType::Int8 : VectorType::Float32, dims);
        s->h_ = out;
        return s;
    }
    Result<Unit> load(const int64_t* rowids, const void* vectors, uint64_t n) {
        int rc = vecgpu_slab_load(h_, rowids, vectors, n);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    Result<Unit> upsert(int64_t rowid, const std::vector<uint8_t>& blob) {  // Vec0Tab::insert / update hook
        int rc = vecgpu_slab_upsert(h_, rowid, blob.empty() ? (const void*)"" : blob.data(), (uint32_t)blob.size());
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    Result<Unit> remove(int64_t rowid) {  // Vec0Tab::delete hook
        int rc = vecgpu_slab_delete(h_, rowid);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    Result<uint64_t> compact() {  // drop the tombstoned rows physically; HNSW indexes over the slab must be rebuilt afterwards
        uint64_t removed = 0;
        int rc = vecgpu_slab_compact(h_, &removed);
        if (rc) return Error::from_status(rc);
        return removed;
    }
    uint64_t live_rows() const {
        uint64_t rows = 0, live = 0;
        vecgpu_slab_count(h_, &rows, &live);
        return live;
    }
    // candidate scoring of search_layer (src/hnsw/search.rs:501-513): NaN for absent rowids
    Result<std::vector<float>> score(const std::vector<uint8_t>& query, const std::vector<int64_t>& cand_rowids, DistanceMetric metric) {
        std::vector<float> out(cand_rowids.size());
        const uint32_t offs[2] = {0, (uint32_t)cand_rowids.size()};
        int rc = vecgpu_score(h_, query.data(), 1, cand_rowids.data(), offs, (int)metric, out.data());
        if (rc) return Error::from_status(rc);
        return out;
    }
};

// ---- brute_force_search (src/vtab.rs:2573-2623) ---------------------------------------------------------------
// `k` follows `k as usize` (vtab.rs:2292): 0 -> empty; larger than the table -> every live row.
inline Result<std::vector<std::pair<int64_t, float>>> brute_force_search(Slab& slab, const std::vector<uint8_t>& query_vector, size_t k,
                                                                         DistanceMetric distance_metric) {
    std::vector<std::pair<int64_t, float>> out;
    if (query_vector.size() != slab.row_bytes()) return out;  // every row fails distance() and is skipped (vtab.rs:2610-2613)
    const uint64_t live = slab.live_rows();
    const uint32_t kk = (uint32_t)std::min<uint64_t>(k, live);
    if (!vecgpu_metric_supported((int)slab.vec_type, (int)distance_metric)) return out;  // distance() errors are swallowed per row
    if (kk == 0) return out;
    std::vector<int64_t> rowids(kk);
    std::vector<float> dists(kk);
    uint32_t count = 0;
    int rc = vecgpu_knn(slab.raw(), query_vector.data(), 1, kk, (int)distance_metric, rowids.data(), dists.data(), &count);
    if (rc) return Error::from_status(rc);
    for (uint32_t i = 0; i < count; ++i) out.emplace_back(rowids[i], dists[i]);
    return out;
}

// ---- ShardedSlab: one column split by contiguous rowid range over the GPUs of the box (one process) ------------
// Same calls as Slab; a query scans every shard and the shards exchange + merge their top-k over NVLink peer memory.
class ShardedSlab {
    vecgpu_sharded* h_ = nullptr;

public:
    VectorType vec_type;
    size_t dims;
    ShardedSlab(VectorType t, size_t d) : vec_type(t), dims(d) {}
    ShardedSlab(const ShardedSlab&) = delete;
    ShardedSlab& operator=(const ShardedSlab&) = delete;
    ~ShardedSlab() { if (h_) vecgpu_sharded_destroy(h_); }
    // devices empty: every visible device; a device may appear more than once (several shards on one GPU)
    static Result<std::unique_ptr<ShardedSlab>> create(VectorType t, size_t dims, const std::vector<int>& devices = {},
                                                       uint64_t capacity_hint_total = 0, uint32_t max_queries = 0, uint32_t max_k = 0) {
        auto s = std::make_unique<ShardedSlab>(t, dims);
        int rc = vecgpu_sharded_create((int)t, (uint32_t)dims, capacity_hint_total, devices.empty() ? nullptr : devices.data(),
                                       (uint32_t)devices.size(), max_queries, max_k, &s->h_);
        if (rc) return Error::from_status(rc);
        return s;
    }
    size_t row_bytes() const { return vecgpu::row_bytes(vec_type, dims); }
    uint32_t num_shards() const { return vecgpu_sharded_num_shards(h_); }
    Result<Unit> load(const int64_t* rowids, const void* vectors, uint64_t n) {
        int rc = vecgpu_sharded_load(h_, rowids, vectors, n);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    Result<Unit> upsert(int64_t rowid, const std::vector<uint8_t>& blob) {
        int rc = vecgpu_sharded_upsert(h_, rowid, blob.empty() ? (const void*)"" : blob.data(), (uint32_t)blob.size());
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    Result<Unit> remove(int64_t rowid) {
        int rc = vecgpu_sharded_delete(h_, rowid);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    uint64_t live_rows() const {
        uint64_t rows = 0, live = 0;
        vecgpu_sharded_count(h_, &rows, &live);
        return live;
    }
    // brute_force_search (src/vtab.rs:2573-2623) over all shards
    Result<std::vector<std::pair<int64_t, float>>> brute_force_search(const std::vector<uint8_t>& query_vector, size_t k, DistanceMetric distance_metric) {
        std::vector<std::pair<int64_t, float>> out;
        if (query_vector.size() != row_bytes()) return out;
        const uint32_t kk = (uint32_t)std::min<uint64_t>(k, live_rows());
        if (!vecgpu_metric_supported((int)vec_type, (int)distance_metric) || kk == 0) return out;
        std::vector<int64_t> rowids(kk);
        std::vector<float> dists(kk);
        uint32_t count = 0;
        int rc = vecgpu_sharded_knn(h_, query_vector.data(), 1, kk, (int)distance_metric, rowids.data(), dists.data(), &count);
        if (rc) return Error::from_status(rc);
        for (uint32_t i = 0; i < count; ++i) out.emplace_back(rowids[i], dists[i]);
        return out;
    }
};

// ---- HnswIndex ------------------------------------------------------------------------------------------------
class HnswIndex {
    vecgpu_hnsw* h_ = nullptr;
    DistanceMetric metric_, internal_;
    bool normalize_;

public:
    HnswIndex(DistanceMetric m, bool normalize_vectors) : metric_(m), internal_(internal_distance_metric(m, normalize_vectors)), normalize_(normalize_vectors) {}
    HnswIndex(const HnswIndex&) = delete;
    ~HnswIndex() { if (h_) vecgpu_hnsw_destroy(h_); }
    // defaults of HnswParams (src/hnsw/mod.rs:35-47): M=32, ef_construction=400
    static Result<std::unique_ptr<HnswIndex>> create(Slab& slab, DistanceMetric metric, uint32_t M = 32, uint32_t ef_construction = 400,
                                                     uint64_t seed = 42, bool normalize_vectors = true) {
        auto idx = std::make_unique<HnswIndex>(metric, normalize_vectors);
        int rc = vecgpu_hnsw_create(slab.raw(), (int)idx->internal_, M, ef_construction, seed, &idx->h_);
        if (rc) return Error::from_status(rc);
        return idx;
    }
    Result<Unit> rebuild(uint32_t batch = 0) {  // vec_rebuild_hnsw; 0 = library default (16384 inserts per launch)
        int rc = vecgpu_hnsw_build(h_, batch);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    // insert_hnsw for rows appended to the slab in rowid order since the last rebuild / insert (src/hnsw/insert.rs:279-532)
    Result<uint64_t> insert_appended(uint32_t batch = 0) {
        uint64_t n = 0;
        int rc = vecgpu_hnsw_insert_appended(h_, batch, &n);
        if (rc) return Error::from_status(rc);
        return n;
    }
    // Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): node and edges deleted, row inserted again
    Result<Unit> reinsert(int64_t rowid) {
        int rc = vecgpu_hnsw_reinsert(h_, rowid);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    // insert_hnsw for a row the slab has just received OUT of rowid order (row positions moved by one): renumbered in place
    Result<Unit> insert_at(int64_t rowid) {
        int rc = vecgpu_hnsw_insert_at(h_, rowid);
        if (rc) return Error::from_status(rc);
        return Unit{};
    }
    // search_hnsw (src/hnsw/search.rs:267-335); the query must already be in the stored representation
    Result<std::vector<std::pair<int64_t, float>>> search(const std::vector<uint8_t>& query, uint32_t k, uint32_t ef_search = 200) {
        std::vector<int64_t> rowids(k);
        std::vector<float> dists(k);
        uint32_t count = 0;
        int rc = vecgpu_hnsw_search(h_, query.data(), 1, k, ef_search, rowids.data(), dists.data(), &count);
        if (rc) return Error::from_status(rc);
        std::vector<std::pair<int64_t, float>> out;
        for (uint32_t i = 0; i < count; ++i) out.emplace_back(rowids[i], convert_distance_for_output(metric_, normalize_, dists[i]));
        return out;
    }
    // BATCH_SIZE_1_4 / _5_16 / _17_32 / _33_64 / _65_PLUS (src/hnsw/search.rs:73-85): expansions by neighbours scored
    Result<std::array<uint64_t, 5>> batch_histogram() {
        std::array<uint64_t, 5> h{};
        int rc = vecgpu_hnsw_batch_histogram(h_, h.data());
        if (rc) return Error::from_status(rc);
        return h;
    }
    // HnswMetadata.entry_point_rowid / entry_point_level (src/hnsw/mod.rs:97-103); (-1, -1) when empty
    Result<std::pair<int64_t, int32_t>> entry_point() {
        int64_t rowid = -1;
        int32_t level = -1;
        int rc = vecgpu_hnsw_entry_point(h_, &rowid, &level);
        if (rc) return Error::from_status(rc);
        return std::make_pair(rowid, level);
    }
};

}  // namespace vecgpu
