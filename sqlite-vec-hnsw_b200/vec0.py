"""Host-side mirror of the reference's interface for the distance-scoring path.

Same names, argument meaning and error behaviour as the Rust crate, so the
parity tests read like the reference's own tests:

  VectorType / DistanceMetric      src/vector.rs:9-46, src/distance/mod.rs:12-44
  Vector                           src/vector.rs:215-608 (the parts on the path)
  distance(a, b, metric)           src/distance/mod.rs:52-84
  brute_force_search(...)          src/vtab.rs:2573-2623
  Slab.score(...)                  the neighbour loop of src/hnsw/search.rs:501-513
  internal_distance_metric /
  convert_distance_for_output      src/hnsw/mod.rs:129-146

All arithmetic happens in libvecgpu.so on the GPU.  Nothing here computes a
distance on the CPU; with the library or the device missing every call raises.
"""
import ctypes as C
import enum
import json
import struct

import numpy as np

from . import _lib


# ---------------------------------------------------------------- errors (src/error.rs:5-36)
class VecError(Exception):
    pass


class InvalidVectorFormat(VecError):
    pass


class DimensionMismatch(VecError):
    def __init__(self, message, expected=None, actual=None):
        super().__init__(message)
        self.expected, self.actual = expected, actual


class InvalidVectorType(VecError):
    pass


class InvalidDistanceMetric(VecError):
    pass


class NotImplementedVec(VecError):
    pass


class InvalidParameter(VecError):
    pass


class InvalidState(VecError):
    pass


def _raise(code):
    msg = _lib.last_error()
    if code == _lib.ERR_DIM_MISMATCH:
        raise DimensionMismatch(msg)
    if code == _lib.ERR_UNSUPPORTED:
        if msg.startswith("invalid vector type"):
            raise InvalidVectorType(msg)
        raise InvalidDistanceMetric(msg)
    if code == _lib.ERR_INVALID_PARAM:
        raise InvalidParameter(msg)
    raise InvalidState(msg)


def _check(code):
    if code != _lib.OK:
        _raise(code)


# ---------------------------------------------------------------- enums
class VectorType(enum.IntEnum):
    Float32 = 0
    Int8 = 1
    Bit = 2

    @staticmethod
    def from_str(s):  # src/vector.rs:30-37
        t = s.lower()
        if t in ("float32", "float"):
            return VectorType.Float32
        if t == "int8":
            return VectorType.Int8
        if t in ("bit", "binary"):
            return VectorType.Bit
        raise InvalidVectorType(s)

    def as_str(self):  # src/vector.rs:40-46
        return ("float32", "int8", "bit")[int(self)]

    def row_bytes(self, dims):
        return (dims * 4, dims, (dims + 7) // 8)[int(self)]


class DistanceMetric(enum.IntEnum):
    L2 = 0
    L1 = 1
    Cosine = 2
    Hamming = 3

    @staticmethod
    def from_str(s):  # src/distance/mod.rs:26-34
        t = s.lower()
        if t in ("l2", "euclidean"):
            return DistanceMetric.L2
        if t in ("l1", "manhattan"):
            return DistanceMetric.L1
        if t == "cosine":
            return DistanceMetric.Cosine
        if t == "hamming":
            return DistanceMetric.Hamming
        raise InvalidDistanceMetric(s)

    def as_str(self):  # src/distance/mod.rs:37-44
        return ("l2", "l1", "cosine", "hamming")[int(self)]


_NP = {VectorType.Float32: np.dtype("<f4"), VectorType.Int8: np.dtype("i1"), VectorType.Bit: np.dtype("u1")}


def _ptr(a):
    # keeps `a` alive for the duration of the call (the ctypes object references the array).  The per-query hot paths
    # (Slab.knn, Exchange.shard_knn) pass `named_array.ctypes.data` instead: data_as() costs several microseconds per argument.
    return a.ctypes.data_as(C.c_void_p)


def _as_raw(x, vec_type):
    """bytes / uint8 arrays are taken as raw little-endian blobs; anything else is
    converted to the element type of the column."""
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype="u1")
    a = np.asarray(x)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a, dtype=_NP[vec_type])


# ---------------------------------------------------------------- Vector (src/vector.rs:215-608)
class Vector:
    """Owned vector: (vec_type, dimensions, little-endian bytes)."""

    __slots__ = ("vec_type", "dimensions", "data")

    def __init__(self, vec_type, dimensions, data):
        self.vec_type = VectorType(vec_type)
        self.dimensions = int(dimensions)
        self.data = bytes(data)

    @staticmethod
    def from_f32(values):  # :217-227
        a = np.asarray(values, dtype="<f4")
        return Vector(VectorType.Float32, a.size, a.tobytes())

    @staticmethod
    def from_i8(values):  # :230-236
        a = np.asarray(values, dtype="i1")
        return Vector(VectorType.Int8, a.size, a.tobytes())

    @staticmethod
    def from_json(text, vec_type):  # :239-256
        try:
            values = json.loads(text)
            values = [float(v) for v in values]
        except Exception as e:  # serde_json error
            raise InvalidVectorFormat(f"JSON parsing error: {e}")
        vec_type = VectorType(vec_type)
        if vec_type == VectorType.Float32:
            return Vector.from_f32(np.asarray(values, dtype=np.float64).astype("<f4"))
        if vec_type == VectorType.Int8:
            # Rust `f64 as i8` saturates and truncates toward zero
            return Vector.from_i8(np.clip(np.trunc(np.asarray(values, dtype=np.float64)), -128, 127).astype("i1"))
        raise NotImplementedVec("Binary vector from JSON not yet implemented")

    @staticmethod
    def from_blob(blob, vec_type, dimensions):  # :259-266 (no size validation there either)
        return Vector(vec_type, dimensions, blob)

    def as_bytes(self):
        return self.data

    def as_f32(self):
        if self.vec_type != VectorType.Float32:
            raise InvalidVectorType("Vector is not Float32 type")
        return np.frombuffer(self.data, dtype="<f4")

    def as_i8(self):
        if self.vec_type != VectorType.Int8:
            raise InvalidVectorType("Vector is not Int8 type")
        return np.frombuffer(self.data, dtype="i1")

    # ---- producers, on the GPU (K7)
    def _f32_in(self, what):
        if self.vec_type != VectorType.Float32:
            raise InvalidVectorType(what)
        return np.ascontiguousarray(np.frombuffer(self.data, dtype="<f4"))

    def normalize(self, device=0):  # :444-466
        if self.vec_type == VectorType.Int8:
            raise InvalidVectorType("Cannot normalize Int8 vectors (would lose precision)")
        if self.vec_type == VectorType.Bit:
            raise InvalidVectorType("Cannot normalize binary vectors")
        return Vector.from_f32(normalize(self._f32_in("")[None, :], device)[0])

    def quantize_int8(self, device=0):  # :514-545
        return Vector.from_i8(quantize_int8(self._f32_in("Can only quantize Float32 vectors")[None, :], device)[0])

    def quantize_int8_for_index(self, device=0):  # :554-575
        return Vector.from_i8(
            quantize_int8_for_index(self._f32_in("Can only quantize Float32 vectors")[None, :], device)[0]
        )

    def quantize_binary(self, device=0):  # :579-608
        bits = quantize_binary(self._f32_in("Can only quantize Float32 vectors to binary")[None, :], device)[0]
        return Vector(VectorType.Bit, self.dimensions, bits.tobytes())


def _producer(fn_name, x, out_dtype, out_cols, device):
    x = np.ascontiguousarray(x, dtype="<f4")
    if x.ndim != 2:
        raise InvalidParameter("expected a 2-D array [n, dims]")
    n, d = x.shape
    out = np.empty((n, out_cols(d)), dtype=out_dtype)
    _check(getattr(_lib.load(), fn_name)(_ptr(x), n, d, device, _ptr(out)))
    return out


def normalize(x, device=0):
    return _producer("vecgpu_normalize_f32", x, "<f4", lambda d: d, device)


def quantize_int8(x, device=0):
    return _producer("vecgpu_quantize_int8", x, "i1", lambda d: d, device)


def quantize_int8_for_index(x, device=0):
    return _producer("vecgpu_quantize_int8_for_index", x, "i1", lambda d: d, device)


def quantize_binary(x, device=0):
    return _producer("vecgpu_quantize_binary", x, "u1", lambda d: (d + 7) // 8, device)


# ---------------------------------------------------------------- distance() (src/distance/mod.rs:52-84)
def distance(a, b, metric, device=0):
    """One pair.  Check order as in the reference: dimensions, types, then the
    (type, metric) match."""
    metric = DistanceMetric(metric)
    if a.dimensions != b.dimensions:
        raise DimensionMismatch(
            f"Dimension mismatch: expected {a.dimensions}, got {b.dimensions}", a.dimensions, b.dimensions
        )
    if a.vec_type != b.vec_type:
        raise InvalidVectorType("Vector types must match for distance calculation")
    rb = a.vec_type.row_bytes(a.dimensions)
    if len(a.data) != rb or len(b.data) != rb:
        # simsimd returns None on mismatched slices -> InvalidParameter (src/distance/scalar.rs:18)
        raise InvalidParameter("distance calculation failed: blob length does not match dimensions")
    out = np.empty(1, dtype="<f4")
    abuf = np.frombuffer(a.data, dtype="u1")
    bbuf = np.frombuffer(b.data, dtype="u1")
    _check(
        _lib.load().vecgpu_distance_pairs(
            int(a.vec_type), a.dimensions, b.dimensions, _ptr(abuf), _ptr(bbuf), 1, int(metric), device, _ptr(out)
        )
    )
    return float(out[0])


def distance_pairs(vec_type, a, b, metric, device=0):
    """n pairs at once: a[i] vs b[i]; a, b are [n, dims] (f32 / i8) or [n, nbytes] (bit, with dims given by 8*nbytes)."""
    vec_type, metric = VectorType(vec_type), DistanceMetric(metric)
    a = np.ascontiguousarray(a, dtype=_NP[vec_type])
    b = np.ascontiguousarray(b, dtype=_NP[vec_type])
    da = a.shape[1] * (8 if vec_type == VectorType.Bit else 1)
    db = b.shape[1] * (8 if vec_type == VectorType.Bit else 1)
    out = np.empty(a.shape[0], dtype="<f4")
    _check(
        _lib.load().vecgpu_distance_pairs(
            int(vec_type), da, db, _ptr(a), _ptr(b), a.shape[0], int(metric), device, _ptr(out)
        )
    )
    return out


# ---------------------------------------------------------------- HNSW metric rule (src/hnsw/mod.rs:129-146)
def internal_distance_metric(metric, normalize_vectors):
    metric = DistanceMetric(metric)
    return DistanceMetric.L2 if (metric == DistanceMetric.Cosine and normalize_vectors) else metric


def convert_distance_for_output(metric, normalize_vectors, internal_dist):
    if DistanceMetric(metric) == DistanceMetric.Cosine and normalize_vectors:
        d = np.float32(internal_dist)
        return float(np.float32(np.float32(d * d) / np.float32(2.0)))
    return float(internal_dist)


# ---------------------------------------------------------------- Slab
class Slab:
    """HBM-resident, rowid-indexed copy of one vector column ({table}_data.vecNN)."""

    def __init__(self, vec_type, dims, capacity_hint=0, device=0):
        self.vec_type = VectorType(vec_type)
        self.dims = int(dims)
        self.device = device
        self.row_bytes = self.vec_type.row_bytes(self.dims)
        self._h = C.c_void_p()
        self._lib = _lib.load()
        _check(self._lib.vecgpu_slab_create(int(self.vec_type), self.dims, capacity_hint, device, C.byref(self._h)))

    @classmethod
    def _adopt(cls, handle, vec_type, dims, device):
        """Wrap a slab handle the library created (e.g. vecgpu_hnsw_stored_slab); the wrapper owns it."""
        self = cls.__new__(cls)
        self.vec_type = VectorType(vec_type)
        self.dims = int(dims)
        self.device = device
        self.row_bytes = self.vec_type.row_bytes(self.dims)
        self._lib = _lib.load()
        self._h = handle
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vecgpu_slab_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _rows(self, vectors, n=None):
        v = _as_raw(vectors, self.vec_type).reshape(-1)
        if v.size * v.itemsize % self.row_bytes:
            raise InvalidParameter("vector buffer is not a whole number of rows")
        return v

    def load(self, vectors, rowids=None):
        v = self._rows(vectors)
        n = v.size * v.itemsize // self.row_bytes
        r = None if rowids is None else np.ascontiguousarray(rowids, dtype="<i8")
        if r is not None and r.size != n:
            raise InvalidParameter("rowids and vectors disagree on the number of rows")
        _check(self._lib.vecgpu_slab_load(self._h, None if r is None else _ptr(r), _ptr(v), n))
        return n

    def append(self, vectors, rowids=None):
        v = self._rows(vectors)
        n = v.size * v.itemsize // self.row_bytes
        r = None if rowids is None else np.ascontiguousarray(rowids, dtype="<i8")
        _check(self._lib.vecgpu_slab_append(self._h, None if r is None else _ptr(r), _ptr(v), n))
        return n

    def upsert(self, rowid, blob):
        """blob: bytes of one row; a wrong-length / empty blob marks the row as skipped."""
        b = bytes(blob) if blob is not None else b""
        buf = np.frombuffer(b, dtype="u1") if b else np.zeros(1, dtype="u1")
        _check(self._lib.vecgpu_slab_upsert(self._h, int(rowid), _ptr(buf), len(b)))

    def delete(self, rowid):
        _check(self._lib.vecgpu_slab_delete(self._h, int(rowid)))

    def compact(self):
        """Physically drop deleted / unreadable rows; -> number of rows removed.  HNSW indexes over the slab must be rebuilt."""
        removed = C.c_uint64()
        _check(self._lib.vecgpu_slab_compact(self._h, C.byref(removed)))
        return removed.value

    def count(self):
        rows, live = C.c_uint64(), C.c_uint64()
        _check(self._lib.vecgpu_slab_count(self._h, C.byref(rows), C.byref(live)))
        return rows.value, live.value

    def get(self, rowid):
        out = np.empty(self.row_bytes, dtype="u1")
        found = C.c_int()
        _check(self._lib.vecgpu_slab_get(self._h, int(rowid), _ptr(out), C.byref(found)))
        return out.tobytes() if found.value else None

    def fill_synthetic(self, seed, n, first_rowid=1, kind=_lib.SYNTH_UNIFORM):
        _check(self._lib.vecgpu_slab_fill_synthetic(self._h, seed, first_rowid, n, kind))

    def device_view(self):
        p, stride, rows = C.c_void_p(), C.c_uint32(), C.c_uint64()
        _check(self._lib.vecgpu_slab_device_view(self._h, C.byref(p), C.byref(stride), C.byref(rows)))
        return p.value, stride.value, rows.value

    def _queries(self, queries):
        q = _as_raw(queries, self.vec_type)
        nbytes = q.size * q.itemsize
        if nbytes == 0 or nbytes % self.row_bytes:
            raise DimensionMismatch(
                f"Dimension mismatch: expected {self.dims}, got a query of {nbytes} bytes", self.dims, None
            )
        return q, nbytes // self.row_bytes

    def knn(self, queries, k, metric):
        """-> (rowids [nq,k] i64, dists [nq,k] f32, counts [nq] u32)."""
        q, nq = self._queries(queries)
        k = int(k)
        rowids = np.empty((nq, k), dtype="<i8")  # every slot is written by the library (padding: -1 / +inf)
        dists = np.empty((nq, k), dtype="<f4")
        counts = np.empty(nq, dtype="<u4")
        _check(self._lib.vecgpu_knn(self._h, q.ctypes.data, nq, k, int(metric), rowids.ctypes.data, dists.ctypes.data, counts.ctypes.data))
        return rowids, dists, counts

    def score(self, queries, cand_rowids, cand_offsets, metric):
        """CSR candidate lists -> distances (NaN for absent / skipped rowids)."""
        q, nq = self._queries(queries)
        cr = np.ascontiguousarray(cand_rowids, dtype="<i8")
        co = np.ascontiguousarray(cand_offsets, dtype="<u4")
        if co.size != nq + 1:
            raise InvalidParameter("cand_offsets must have nq+1 entries")
        out = np.empty(int(co[-1]), dtype="<f4")
        if cr.size != out.size:
            raise InvalidParameter("cand_rowids length must equal cand_offsets[-1]")
        _check(self._lib.vecgpu_score(self._h, _ptr(q), nq, _ptr(cr), _ptr(co), int(DistanceMetric(metric)), _ptr(out)))
        return out

    # ---- device-resident variants: torch tensors in, torch tensors out (no host copies)
    def knn_device(self, d_queries, k, metric, stream=None):
        import torch

        nq = d_queries.numel() * d_queries.element_size() // self.row_bytes
        rowids = torch.empty((nq, k), dtype=torch.int64, device=d_queries.device)
        dists = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
        st = torch.cuda.current_stream(d_queries.device).cuda_stream if stream is None else stream
        _check(
            self._lib.vecgpu_knn_device(
                self._h, d_queries.data_ptr(), nq, int(k), int(DistanceMetric(metric)), rowids.data_ptr(),
                dists.data_ptr(), C.c_void_p(st),
            )
        )
        return rowids, dists


def merge_device(d_rowids, d_dists, stream=None):
    """[nlists, nq, k] per-shard results (torch, on one device) -> global top-k [nq, k]."""
    import torch

    nlists, nq, k = d_rowids.shape
    out_r = torch.empty((nq, k), dtype=torch.int64, device=d_rowids.device)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=d_rowids.device)
    st = torch.cuda.current_stream(d_rowids.device).cuda_stream if stream is None else stream
    _check(
        _lib.load().vecgpu_merge_device(
            d_rowids.device.index or 0, d_rowids.contiguous().data_ptr(), d_dists.contiguous().data_ptr(), nlists, nq, k,
            out_r.data_ptr(), out_d.data_ptr(), C.c_void_p(st),
        )
    )
    return out_r, out_d


# ---------------------------------------------------------------- sharded slabs, C entry points of csrc/xchg_host.inl
class Exchange:
    """One rank's exchange endpoint (vecgpu_xchg): gather buffer in this GPU's HBM that the peers write their local top-k
    into over NVLink.  One per (rank, GPU); introduce the endpoints with attach_local (same process) or
    export_handle + attach_ipc (one process per GPU)."""

    def __init__(self, device, rank, world, max_queries=1024, max_k=128):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device, self.rank, self.world = device, rank, world
        _check(self._lib.vecgpu_xchg_create(device, rank, world, max_queries, max_k, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vecgpu_xchg_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def export_handle(self):
        buf = np.zeros(_lib.XCHG_HANDLE_BYTES, dtype="u1")
        _check(self._lib.vecgpu_xchg_export(self._h, _ptr(buf)))
        return buf

    def attach_ipc(self, handles):
        """handles: [world, 128] u8, row p = rank p's export_handle()."""
        h = np.ascontiguousarray(handles, dtype="u1").reshape(self.world, _lib.XCHG_HANDLE_BYTES)
        _check(self._lib.vecgpu_xchg_attach_ipc(self._h, _ptr(h), self.world))

    @staticmethod
    def attach_local(endpoints):
        arr = (C.c_void_p * len(endpoints))(*[e._h.value for e in endpoints])
        _check(_lib.load().vecgpu_xchg_attach_local(arr, len(endpoints)))

    def shard_knn(self, slab, queries, k, metric):
        """vecgpu_shard_knn: this rank's part of a sharded query, host buffers; -> global (rowids, dists, counts)."""
        q, nq = slab._queries(queries)
        k = int(k)
        rowids = np.empty((nq, k), dtype="<i8")  # every slot is written by the library (padding: -1 / +inf)
        dists = np.empty((nq, k), dtype="<f4")
        counts = np.empty(nq, dtype="<u4")
        _check(self._lib.vecgpu_shard_knn(slab._h, self._h, q.ctypes.data, nq, k, int(metric), rowids.ctypes.data, dists.ctypes.data,
                                          counts.ctypes.data))
        return rowids, dists, counts

    def shard_knn_device(self, slab, d_queries, k, metric, stream=None):
        import torch

        nq = d_queries.numel() * d_queries.element_size() // slab.row_bytes
        rowids = torch.empty((nq, k), dtype=torch.int64, device=d_queries.device)
        dists = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
        st = torch.cuda.current_stream(d_queries.device).cuda_stream if stream is None else stream
        _check(self._lib.vecgpu_shard_knn_device(slab._h, self._h, d_queries.data_ptr(), nq, int(k), int(DistanceMetric(metric)),
                                                 rowids.data_ptr(), dists.data_ptr(), C.c_void_p(st)))
        return rowids, dists

    def merge_device(self, d_rowids, d_dists, stream=None):
        """Exchange + merge of this rank's [nq, k] device results with every peer's (collective)."""
        import torch

        nq, k = d_rowids.shape
        out_r = torch.empty((nq, k), dtype=torch.int64, device=d_rowids.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=d_rowids.device)
        st = torch.cuda.current_stream(d_rowids.device).cuda_stream if stream is None else stream
        _check(self._lib.vecgpu_xchg_merge_device(self._h, d_rowids.contiguous().data_ptr(), d_dists.contiguous().data_ptr(), None, nq, k,
                                                  out_r.data_ptr(), out_d.data_ptr(), C.c_void_p(st)))
        return out_r, out_d

    def check(self, stream=None):
        _check(self._lib.vecgpu_xchg_check(self._h, C.c_void_p(stream or 0)))


class ShardedSlab:
    """vecgpu_sharded: one slab cut by rowid range over several GPUs of THIS process (the form the Rust extension links)."""

    def __init__(self, vec_type, dims, devices=None, capacity_hint=0, max_queries=1024, max_k=128):
        self.vec_type = VectorType(vec_type)
        self.dims = int(dims)
        self.row_bytes = self.vec_type.row_bytes(self.dims)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if devices is None:
            arr, n = None, 0
        else:
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        _check(self._lib.vecgpu_sharded_create(int(self.vec_type), self.dims, capacity_hint, arr, n, max_queries, max_k, C.byref(self._h)))
        self.n_shards = int(self._lib.vecgpu_sharded_num_shards(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vecgpu_sharded_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def load(self, vectors, rowids=None):
        v = _as_raw(vectors, self.vec_type).reshape(-1)
        n = v.size * v.itemsize // self.row_bytes
        r = None if rowids is None else np.ascontiguousarray(rowids, dtype="<i8")
        _check(self._lib.vecgpu_sharded_load(self._h, None if r is None else _ptr(r), _ptr(v), n))
        return n

    def fill_synthetic(self, seed, n, first_rowid=1, kind=_lib.SYNTH_UNIFORM):
        _check(self._lib.vecgpu_sharded_fill_synthetic(self._h, seed, first_rowid, n, kind))

    def upsert(self, rowid, blob):
        b = bytes(blob) if blob is not None else b""
        buf = np.frombuffer(b, dtype="u1") if b else np.zeros(1, dtype="u1")
        _check(self._lib.vecgpu_sharded_upsert(self._h, int(rowid), _ptr(buf), len(b)))

    def delete(self, rowid):
        _check(self._lib.vecgpu_sharded_delete(self._h, int(rowid)))

    def count(self):
        rows, live = C.c_uint64(), C.c_uint64()
        _check(self._lib.vecgpu_sharded_count(self._h, C.byref(rows), C.byref(live)))
        return rows.value, live.value

    _queries = Slab._queries

    def knn(self, queries, k, metric):
        q, nq = self._queries(queries)
        k = int(k)
        rowids = np.full((nq, k), -1, dtype="<i8")
        dists = np.full((nq, k), np.inf, dtype="<f4")
        counts = np.zeros(nq, dtype="<u4")
        _check(self._lib.vecgpu_sharded_knn(self._h, _ptr(q), nq, k, int(DistanceMetric(metric)), _ptr(rowids), _ptr(dists), _ptr(counts)))
        return rowids, dists, counts


# ---------------------------------------------------------------- brute_force_search (src/vtab.rs:2573-2623)
def brute_force_search(slab, query_vector, k, distance_metric):
    """Exact k-NN: every live row scored, stable order by (distance, rowid), first k.

    `query_vector` is the raw query blob (bytes), as `filter()` hands it over
    (src/vtab.rs:2119-2143).  `k` follows `k as usize` (src/vtab.rs:2292): 0 ->
    empty, larger than the table -> every live row, negative -> wraps to "all".
    Returns [(rowid, distance_f32), ...].
    Deviation from the reference noted in DESIGN.md: the reference parses both
    sides as Float32 whatever the column type (F3); here the slab's element
    type is used, which is what north_star asks for.
    """
    q = np.frombuffer(bytes(query_vector), dtype="u1")
    if q.size != slab.row_bytes:
        # distance() would fail DimensionMismatch on every row and each row would be skipped (vtab.rs:2610-2613)
        return []
    k = int(k)
    rows, live = slab.count()
    if k < 0 or k > live:
        k = live
    if k == 0:
        return []
    rowids, dists, counts = slab.knn(q, k, distance_metric)
    n = int(counts[0])
    return [(int(rowids[0, i]), float(dists[0, i])) for i in range(n)]


# ---------------------------------------------------------------- HNSW (src/hnsw/{insert,search,rebuild}.rs)
class HnswIndex:
    """Graph over the rows of a slab that holds the STORED node vectors; built and searched with GPU-batched
    candidate scoring.  `metric` is the column's metric; cosine columns must load normalised vectors and are
    searched with L2 internally, distances converted on output (src/hnsw/mod.rs:129-146).
    HnswIndex.for_column() derives the stored slab from a float32 column slab the way the reference does
    (normalise for cosine, quantize_int8_for_index with index_quantization=int8: src/hnsw/insert.rs:300-322)."""

    def __init__(self, slab, metric, M=32, ef_construction=400, seed=42, normalize_vectors=True, index_quantization="none",
                 _column_type=None, _owns_slab=False):
        # defaults of HnswParams (src/hnsw/mod.rs:35-47)
        self.slab = slab
        self.metric = DistanceMetric(metric)
        self.normalize_vectors = bool(normalize_vectors)
        self.index_quantization = str(index_quantization).lower()
        if self.index_quantization not in ("none", "int8"):
            raise InvalidParameter(f"unknown index_quantization {index_quantization!r}")
        self.column_type = VectorType(slab.vec_type if _column_type is None else _column_type)
        self._owns_slab = _owns_slab
        self.internal = internal_distance_metric(self.metric, self.normalize_vectors)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _check(self._lib.vecgpu_hnsw_create(slab._h, int(self.internal), M, ef_construction, seed, C.byref(self._h)))

    @classmethod
    def for_column(cls, column_slab, metric, M=32, ef_construction=400, seed=42, index_quantization="none"):
        """Index over a vec0 column: builds the stored representation on the device (vecgpu_hnsw_stored_slab) when the
        column is float32 and cosine and/or index_quantization=int8, otherwise indexes the column slab itself."""
        metric = DistanceMetric(metric)
        f32 = column_slab.vec_type == VectorType.Float32
        norm = f32 and metric == DistanceMetric.Cosine          # HnswMetadata.normalize_vectors (src/hnsw/mod.rs:120-123)
        q8 = f32 and str(index_quantization).lower() == "int8"   # insert.rs:303-313: float32 columns only
        h = C.c_void_p()
        _check(_lib.load().vecgpu_hnsw_stored_slab(column_slab._h, 1 if norm else 0, 1 if q8 else 0, C.byref(h)))
        if h.value:
            stored = Slab._adopt(h, VectorType.Int8 if q8 else VectorType.Float32, column_slab.dims, column_slab.device)
            return cls(stored, metric, M, ef_construction, seed, normalize_vectors=True, index_quantization=index_quantization,
                       _column_type=VectorType.Float32, _owns_slab=True)
        return cls(column_slab, metric, M, ef_construction, seed, normalize_vectors=True, index_quantization="none")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vecgpu_hnsw_destroy(self._h)
            self._h = C.c_void_p()
        if getattr(self, "_owns_slab", False) and self.slab is not None:
            self.slab.close()
            self.slab = None

    __del__ = close

    def rebuild(self, batch=0):
        """vec_rebuild_hnsw (src/sql_functions.rs:436-534): index every live row of the slab."""
        _check(self._lib.vecgpu_hnsw_build(self._h, batch))
        return self.stats()["nodes"]

    def insert_appended(self, batch=0, new_vectors=None, new_rowids=None):
        """insert_hnsw (src/hnsw/insert.rs:279-532) for rows that arrive in rowid order: index the rows appended to the slab
        since the last rebuild / insert_appended.  An index made by for_column() keeps its own slab of STORED vectors
        (normalised and / or int8): pass the raw column vectors of the new rows (and their rowids, if the column's are not
        1..n) and they are converted like the rest (insert.rs:300-322) and appended there first.  -> nodes added."""
        if new_vectors is not None:
            if not getattr(self, "_owns_slab", False):
                raise InvalidParameter("new_vectors is for indexes made by for_column(); append to the column slab instead")
            raw = np.ascontiguousarray(new_vectors, dtype="<f4").reshape(-1, self.slab.dims)
            bad = np.zeros(len(raw), dtype=bool)
            if self.normalize_vectors and self.metric == DistanceMetric.Cosine:
                bad = ~(np.einsum("ij,ij->i", raw.astype(np.float64), raw.astype(np.float64)) > 0)  # "Cannot normalize zero vector"
                safe = raw.copy()
                safe[bad] = 1.0
                raw = normalize(safe)
            stored = quantize_int8_for_index(raw) if self.slab.vec_type == VectorType.Int8 else raw
            first = self.slab.count()[0]
            self.slab.append(stored, new_rowids)
            if bad.any():
                ids = np.arange(first + 1, first + 1 + len(raw)) if new_rowids is None else np.asarray(new_rowids)
                for i in np.flatnonzero(bad):
                    self.slab.upsert(int(ids[i]), b"")  # cannot be stored: a skipped row, like in the rebuild
        n = C.c_uint64()
        _check(self._lib.vecgpu_hnsw_insert_appended(self._h, batch, C.byref(n)))
        return n.value

    def reinsert(self, rowid, new_vector=None):
        """Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): the row's node and its edges in both directions are
        deleted and the row is inserted again with the vector it has in the slab now.  An index made by for_column() takes
        the new raw column vector (None / empty: the row leaves the index) and stores its converted form first."""
        rowid = int(rowid)
        if getattr(self, "_owns_slab", False):
            if new_vector is None or len(new_vector) == 0:
                self.slab.upsert(rowid, b"")
            else:
                raw = np.ascontiguousarray(new_vector, dtype="<f4").reshape(1, self.slab.dims)
                ok = True
                if self.normalize_vectors and self.metric == DistanceMetric.Cosine:
                    ok = float(np.dot(raw[0].astype(np.float64), raw[0].astype(np.float64))) > 0
                    raw = normalize(raw) if ok else raw
                stored = quantize_int8_for_index(raw) if self.slab.vec_type == VectorType.Int8 else raw
                self.slab.upsert(rowid, stored.tobytes() if ok else b"")
        _check(self._lib.vecgpu_hnsw_reinsert(self._h, rowid))

    def insert_at(self, rowid, new_vector=None):
        """insert_hnsw for a row inserted OUT of rowid order (Vec0Tab::insert with an explicit rowid, src/vtab.rs:1409-1682): the
        slab has just received `rowid` between existing rows (every later row moved one position up); the resident graph is
        renumbered on the device and the row is inserted like any other.  One such row per call.  An index made by
        for_column() takes the raw column vector and stores its converted form first."""
        rowid = int(rowid)
        if getattr(self, "_owns_slab", False):
            if new_vector is None or len(new_vector) == 0:
                self.slab.upsert(rowid, b"")
            else:
                raw = np.ascontiguousarray(new_vector, dtype="<f4").reshape(1, self.slab.dims)
                ok = True
                if self.normalize_vectors and self.metric == DistanceMetric.Cosine:
                    ok = float(np.dot(raw[0].astype(np.float64), raw[0].astype(np.float64))) > 0
                    raw = normalize(raw) if ok else raw
                stored = quantize_int8_for_index(raw) if self.slab.vec_type == VectorType.Int8 else raw
                self.slab.upsert(rowid, stored.tobytes() if ok else b"")
        _check(self._lib.vecgpu_hnsw_insert_at(self._h, rowid))

    def stats(self):
        n, e, r, sc = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        lvl = C.c_int32()
        _check(self._lib.vecgpu_hnsw_stats(self._h, C.byref(n), C.byref(e), C.byref(lvl), C.byref(sc), C.byref(r)))
        return dict(nodes=n.value, edges=e.value, entry_level=lvl.value, distances_scored=sc.value, rounds=r.value)

    def entry_point(self):
        """-> (rowid, level) of the graph's entry point, (-1, -1) when empty."""
        r, l = C.c_int64(), C.c_int32()
        _check(self._lib.vecgpu_hnsw_entry_point(self._h, C.byref(r), C.byref(l)))
        return r.value, l.value

    def device_stats(self):
        """Counters of the on-device search kernel: queries/inserts it answered, capacity fallbacks, launches."""
        q, f, l = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._lib.vecgpu_hnsw_device_stats(self._h, C.byref(q), C.byref(f), C.byref(l)))
        return dict(queries=q.value, fallbacks=f.value, launches=l.value)

    def batch_histogram(self):
        """Expansions of the device walks by number of unvisited neighbours scored (the reference's BATCH_SIZE_* buckets,
        src/hnsw/search.rs:443-455)."""
        h = (C.c_uint64 * 5)()
        _check(self._lib.vecgpu_hnsw_batch_histogram(self._h, h))
        return {"1-4": h[0], "5-16": h[1], "17-32": h[2], "33-64": h[3], "65+": h[4]}

    def search(self, queries, k, ef_search=200):
        """search_hnsw (src/hnsw/search.rs:267-335).  Queries are raw column vectors; cosine queries are
        normalised here (search.rs:291-293).  -> (rowids [nq,k], distances in the column's metric, counts)."""
        q = _as_raw(queries, self.column_type)
        if self.column_type == VectorType.Float32:
            # the query gets the stored representation's treatment (search.rs:285-302): normalise, then quantise
            if self.metric == DistanceMetric.Cosine and self.normalize_vectors:
                q = normalize(np.frombuffer(q.tobytes(), dtype="<f4").reshape(-1, self.slab.dims), self.slab.device)
            if self.index_quantization == "int8" and self.slab.vec_type == VectorType.Int8:
                q = quantize_int8_for_index(np.frombuffer(np.ascontiguousarray(q).tobytes(), dtype="<f4").reshape(-1, self.slab.dims),
                                            self.slab.device)
        nbytes = q.size * q.itemsize
        if nbytes == 0 or nbytes % self.slab.row_bytes:
            raise DimensionMismatch(f"Dimension mismatch: expected {self.slab.dims}", self.slab.dims, None)
        nq = nbytes // self.slab.row_bytes
        rowids = np.full((nq, k), -1, dtype="<i8")
        dists = np.full((nq, k), np.inf, dtype="<f4")
        counts = np.zeros(nq, dtype="<u4")
        _check(self._lib.vecgpu_hnsw_search(self._h, _ptr(np.ascontiguousarray(q)), nq, k, ef_search, _ptr(rowids), _ptr(dists), _ptr(counts)))
        if self.internal != self.metric:  # cosine: d_out = d_L2^2 / 2
            with np.errstate(invalid="ignore"):
                dists = ((dists * dists) / np.float32(2.0)).astype("<f4")
        return rowids, dists, counts

    def export_nodes(self):
        """-> (rowid, level) arrays in the shape of the {t}_{c}_hnsw_nodes table (the vector column is the slab row)."""
        n = C.c_uint64()
        _check(self._lib.vecgpu_hnsw_export_nodes(self._h, 0, None, None, C.byref(n)))
        rid, lv = np.empty(n.value, dtype="<i8"), np.empty(n.value, dtype="<i4")
        _check(self._lib.vecgpu_hnsw_export_nodes(self._h, n.value, _ptr(rid), _ptr(lv), C.byref(n)))
        return rid, lv

    def export_edges(self):
        """-> (from_rowid, to_rowid, level, distance) arrays in the shape of the {t}_{c}_hnsw_edges table."""
        n = C.c_uint64()
        _check(self._lib.vecgpu_hnsw_export_edges(self._h, 0, None, None, None, None, C.byref(n)))
        fr, to = np.empty(n.value, dtype="<i8"), np.empty(n.value, dtype="<i8")
        lv, ds = np.empty(n.value, dtype="<i4"), np.empty(n.value, dtype="<f4")
        _check(self._lib.vecgpu_hnsw_export_edges(self._h, n.value, _ptr(fr), _ptr(to), _ptr(lv), _ptr(ds), C.byref(n)))
        return fr, to, lv, ds
