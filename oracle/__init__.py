"""ctypes wrapper of the CPU oracle (oracle/vecgpu_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product (sqlite-vec-hnsw_b200/, libvecgpu.so) never does.
See the header of vecgpu_oracle.c for the parity status of each metric.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvecgpu_oracle.so")
LIB_PATH_AVX512 = os.path.join(_HERE, "libvecgpu_oracle_avx512.so")  # same sources, -march=skylake-avx512 (bit-identical results)
_SOURCES = ("vecgpu_oracle.c", "simsimd_shapes.c", "hnsw_seq.c")

F32, I8, BIT = 0, 1, 2
L2, L1, COSINE, HAMMING = 0, 1, 2, 3
_NP = {F32: np.dtype("<f4"), I8: np.dtype("i1"), BIT: np.dtype("u1")}

_lib = None


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc only)."""
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in _SOURCES)
    if force or any(not os.path.exists(x) or os.path.getmtime(x) < newest for x in (LIB_PATH, LIB_PATH_AVX512)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def _host_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd", "fma", "avx2"} <= fl
    except OSError:
        pass
    return False


LIB_PATH_NATIVE = os.path.join(_HERE, "libvecgpu_oracle_native.so")


def build_native():
    """-march=native build made ON THE MACHINE THAT RUNS IT (bench.py's CPU arm calls this on the GPU box, so the CPU
    number is not held back by a portable build).  Same sources, same flags otherwise: bit-identical results.
    Returns the path, or None when it cannot be built here."""
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _bind(LIB_PATH_NATIVE)  # must load and resolve every symbol on this CPU
        return LIB_PATH_NATIVE
    except Exception:
        return None


def use_library(path):
    """Switch the bound library (e.g. to the native build)."""
    global _lib
    _lib = _bind(path)
    return _lib


def lib_path():
    """The build this host can run fastest; VECGPU_ORACLE_PORTABLE=1 pins the portable one."""
    if os.environ.get("VECGPU_ORACLE_PORTABLE", "0") != "1" and os.path.exists(LIB_PATH_AVX512) and _host_has_avx512():
        return LIB_PATH_AVX512
    return LIB_PATH


def _bind(path):
        L = C.CDLL(path)
        p = C.c_void_p
        L.orc_row_bytes.restype = C.c_uint32
        L.orc_row_bytes.argtypes = [C.c_int, C.c_uint32]
        L.orc_metric_supported.argtypes = [C.c_int, C.c_int]
        L.orc_distance.argtypes = [C.c_int, C.c_uint32, C.c_uint32, p, p, C.c_int, p]
        L.orc_knn.argtypes = [C.c_int, C.c_uint32, p, p, p, C.c_uint64, p, C.c_uint32, C.c_uint32, C.c_int, p, p, p]
        L.orc_distances.argtypes = [C.c_int, C.c_uint32, p, C.c_uint64, p, C.c_int, p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_normalize_f32.argtypes = [p, C.c_uint32, p]
        L.orc_quantize_int8.argtypes = [p, C.c_uint32, p]
        L.orc_quantize_int8_for_index.argtypes = [p, C.c_uint32, p]
        L.orc_quantize_binary.argtypes = [p, C.c_uint32, p]
        L.orc_convert_cosine_output.restype = C.c_float
        L.orc_convert_cosine_output.argtypes = [C.c_float]
        L.orc_synth_rows.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, p]
        L.orc_knn_select.argtypes = L.orc_knn.argtypes
        L.orc_knn_synth.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_int64, C.c_uint64, C.c_int, p, C.c_uint32, C.c_uint32,
                                    C.c_int, p, p, p]
        L.orc_force_plain.argtypes = [C.c_int]
        L.orc_round_haz.restype = C.c_float
        L.orc_round_haz.argtypes = [C.c_float]
        L.orc_shape_supported.argtypes = [C.c_int, C.c_int]
        L.orc_shape_distances_f32.argtypes = [C.c_int, C.c_int, C.c_uint32, p, C.c_uint64, p, C.c_int, p]
        L.orc_shape_distances_i8cos.argtypes = [C.c_int, C.c_uint32, p, C.c_uint64, p, p]
        L.orc_hnsw_new.restype = C.c_void_p
        L.orc_hnsw_new.argtypes = [C.c_int, C.c_uint32, C.c_int, p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_hnsw_free.argtypes = [p]
        L.orc_hnsw_insert.argtypes = [p, C.c_uint32, C.c_int]
        L.orc_hnsw_reinsert.argtypes = [p, C.c_uint32, C.c_int, C.c_int]
        L.orc_hnsw_build.argtypes = [p, p, p]
        L.orc_hnsw_search.restype = C.c_uint32
        L.orc_hnsw_search.argtypes = [p, p, C.c_uint32, C.c_uint32, p, p]
        L.orc_hnsw_export.restype = C.c_uint64
        L.orc_hnsw_export.argtypes = [p, p, p, p, p]
        L.orc_hnsw_info.argtypes = [p, p, p, p, p, p, p]
        L.orc_hnsw_levels.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, p]
        return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = _bind(lib_path())
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def row_bytes(elem, dims):
    return int(lib().orc_row_bytes(elem, dims))


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def distance(elem, a, b, metric, dims_a=None, dims_b=None):
    a = np.ascontiguousarray(a, dtype=_NP[elem])
    b = np.ascontiguousarray(b, dtype=_NP[elem])
    mul = 8 if elem == BIT else 1
    da = a.size * mul if dims_a is None else dims_a
    db = b.size * mul if dims_b is None else dims_b
    out = C.c_float()
    rc = lib().orc_distance(elem, da, db, _ptr(a), _ptr(b), metric, C.byref(out))
    if rc:
        raise OracleError(rc)
    return np.float32(out.value)


def knn(elem, dims, vectors, queries, k, metric, rowids=None, skip=None):
    """-> (rowids [nq,k], dists [nq,k], counts [nq]) with the reference's scan + stable sort + truncate."""
    rb = row_bytes(elem, dims)
    v = np.ascontiguousarray(vectors).view(np.uint8).reshape(-1)
    q = np.ascontiguousarray(queries).view(np.uint8).reshape(-1)
    n, nq = v.size // rb, q.size // rb
    r = None if rowids is None else np.ascontiguousarray(rowids, dtype="<i8")
    s = None if skip is None else np.ascontiguousarray(skip, dtype="u1")
    out_r = np.empty((nq, k), dtype="<i8")
    out_d = np.empty((nq, k), dtype="<f4")
    out_c = np.empty(nq, dtype="<u4")
    rc = lib().orc_knn(
        elem, dims, None if r is None else _ptr(r), _ptr(v), None if s is None else _ptr(s), n, _ptr(q), nq, k, metric,
        _ptr(out_r), _ptr(out_d), _ptr(out_c),
    )
    if rc:
        raise OracleError(rc)
    return out_r, out_d, out_c


def knn_select(elem, dims, vectors, queries, k, metric, rowids=None, skip=None):
    """Same result as knn() (stable sort + truncate == the k smallest of the (distance, position) order), computed by
    per-thread selection on all host threads: the form used for big inputs and for bench.py's CPU arm."""
    rb = row_bytes(elem, dims)
    v = np.ascontiguousarray(vectors).view(np.uint8).reshape(-1)
    q = np.ascontiguousarray(queries).view(np.uint8).reshape(-1)
    n, nq = v.size // rb, q.size // rb
    r = None if rowids is None else np.ascontiguousarray(rowids, dtype="<i8")
    s = None if skip is None else np.ascontiguousarray(skip, dtype="u1")
    out_r = np.empty((nq, k), dtype="<i8")
    out_d = np.empty((nq, k), dtype="<f4")
    out_c = np.empty(nq, dtype="<u4")
    rc = lib().orc_knn_select(
        elem, dims, None if r is None else _ptr(r), _ptr(v), None if s is None else _ptr(s), n, _ptr(q), nq, k, metric,
        _ptr(out_r), _ptr(out_d), _ptr(out_c),
    )
    if rc:
        raise OracleError(rc)
    return out_r, out_d, out_c


def knn_synth(elem, dims, seed, first_rowid, n, kind, queries, k, metric):
    """Exact scan over rows first_rowid..first_rowid+n-1 of the synthetic corpus WITHOUT materialising it (each host
    thread regenerates a row, scores it against every query and keeps its k best): the CPU side of the parity checks
    at BASELINE.json's full sizes."""
    rb = row_bytes(elem, dims)
    q = np.ascontiguousarray(queries).view(np.uint8).reshape(-1)
    nq = q.size // rb
    out_r = np.empty((nq, k), dtype="<i8")
    out_d = np.empty((nq, k), dtype="<f4")
    out_c = np.empty(nq, dtype="<u4")
    rc = lib().orc_knn_synth(elem, dims, seed, first_rowid, n, kind, _ptr(q), nq, k, metric, _ptr(out_r), _ptr(out_d), _ptr(out_c))
    if rc:
        raise OracleError(rc)
    return out_r, out_d, out_c


ACC_SHAPES = ["canonical", "serial", "serial_fma", "lanes8_f64red", "lanes16_hadd", "lanes4", "f64"]
FIN_SHAPES = ["ieee_f64", "ieee_f32", "rsqrt12_nr_f64", "rsqrt14_nr_f64", "rsqrt12_nr_f32"]


def shape_supported(acc, fin):
    return bool(lib().orc_shape_supported(acc, fin))


def shape_distances_f32(acc, fin, vectors, query, metric):
    """f32 L2 / cosine of one query against every row under one of the SimSIMD-shaped accumulation / finish variants
    (oracle/simsimd_shapes.c)."""
    v = np.ascontiguousarray(vectors, dtype="<f4")
    q = np.ascontiguousarray(query, dtype="<f4")
    out = np.empty(v.shape[0], dtype="<f4")
    rc = lib().orc_shape_distances_f32(acc, fin, v.shape[1], _ptr(v), v.shape[0], _ptr(q), metric, _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


def shape_distances_i8cos(fin, vectors, query):
    v = np.ascontiguousarray(vectors, dtype="i1")
    q = np.ascontiguousarray(query, dtype="i1")
    out = np.empty(v.shape[0], dtype="<f4")
    rc = lib().orc_shape_distances_i8cos(fin, v.shape[1], _ptr(v), v.shape[0], _ptr(q), _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


def distances(elem, dims, vectors, query, metric):
    rb = row_bytes(elem, dims)
    v = np.ascontiguousarray(vectors).view(np.uint8).reshape(-1)
    q = np.ascontiguousarray(query).view(np.uint8).reshape(-1)
    n = v.size // rb
    out = np.empty(n, dtype="<f4")
    rc = lib().orc_distances(elem, dims, _ptr(v), n, _ptr(q), metric, _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


def num_threads():
    return int(lib().orc_num_threads())


def set_threads(n):
    lib().orc_set_threads(int(n))


def normalize(x):
    x = np.ascontiguousarray(x, dtype="<f4")
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        rc = lib().orc_normalize_f32(_ptr(x[i]), x.shape[1], _ptr(out[i]))
        if rc:
            raise OracleError(rc)
    return out


def _rowwise(fn, x, out):
    for i in range(x.shape[0]):
        fn(_ptr(x[i]), x.shape[1], _ptr(out[i]))
    return out


def quantize_int8(x):
    x = np.ascontiguousarray(x, dtype="<f4")
    return _rowwise(lib().orc_quantize_int8, x, np.empty(x.shape, dtype="i1"))


def quantize_int8_for_index(x):
    x = np.ascontiguousarray(x, dtype="<f4")
    return _rowwise(lib().orc_quantize_int8_for_index, x, np.empty(x.shape, dtype="i1"))


def quantize_binary(x):
    x = np.ascontiguousarray(x, dtype="<f4")
    return _rowwise(lib().orc_quantize_binary, x, np.empty((x.shape[0], (x.shape[1] + 7) // 8), dtype="u1"))


def convert_cosine_output(d):
    return np.float32(lib().orc_convert_cosine_output(float(d)))


def synth_rows(elem, seed, first_rowid, n, dims, kind=0):
    """Regenerate rows first_rowid..first_rowid+n-1 of the synthetic corpus (same bytes as the device generator)."""
    rb = row_bytes(elem, dims)
    out = np.empty(n * rb, dtype="u1")
    lib().orc_synth_rows(elem, seed, first_rowid, n, dims, kind, _ptr(out))
    if elem == F32:
        return out.view("<f4").reshape(n, dims)
    if elem == I8:
        return out.view("i1").reshape(n, dims)
    return out.reshape(n, rb)


class HnswSeq:
    """Strictly sequential HNSW (oracle/hnsw_seq.c): the reference's insert_hnsw / search_hnsw, one insert at a time.
    `vectors` are the STORED node vectors (normalised / quantised by the caller as the column demands), `metric` the
    internal metric.  Node ids are row positions (0-based)."""

    def __init__(self, elem, dims, metric, vectors, M=16, ef_construction=200, quirk=False):
        self.elem, self.dims = elem, dims
        self._v = np.ascontiguousarray(vectors)  # keep alive: the C side borrows it
        self.n = self._v.view(np.uint8).size // row_bytes(elem, dims)
        self._h = lib().orc_hnsw_new(elem, dims, metric, _ptr(self._v), self.n, M, ef_construction, 1 if quirk else 0)
        self.M = M

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_hnsw_free(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def levels(seed, n, M):
        """The product's reproducible level sequence (same hash as csrc/hnsw.inl)."""
        out = np.empty(n, dtype="i1")
        lib().orc_hnsw_levels(seed, n, M, _ptr(out))
        return out

    def insert(self, node, level):
        """insert_hnsw (src/hnsw/insert.rs:279-532) of one row of the array this object was made from."""
        lib().orc_hnsw_insert(self._h, int(node), int(level))

    def reinsert(self, node, level, insert_again=True):
        """Vec0Tab::update (src/vtab.rs:1860-1895): delete the node and its edges in both directions, then insert it again (the
        caller has replaced its vector in the array this object was made from)."""
        lib().orc_hnsw_reinsert(self._h, int(node), int(level), 1 if insert_again else 0)

    def build(self, levels, skip=None):
        lv = np.ascontiguousarray(levels, dtype="i1")
        sk = None if skip is None else np.ascontiguousarray(skip, dtype="u1")
        lib().orc_hnsw_build(self._h, _ptr(lv), None if sk is None else _ptr(sk))

    def search(self, queries, k, ef_search):
        q = np.ascontiguousarray(queries)
        rb = row_bytes(self.elem, self.dims)
        qb = q.view(np.uint8).reshape(-1, rb)
        nodes = np.full((qb.shape[0], k), -1, dtype="<i8")
        dists = np.full((qb.shape[0], k), np.inf, dtype="<f4")
        on = np.empty(k, dtype="<u4")
        od = np.empty(k, dtype="<f4")
        for i in range(qb.shape[0]):
            c = lib().orc_hnsw_search(self._h, _ptr(qb[i]), k, ef_search, _ptr(on), _ptr(od))
            nodes[i, :c] = on[:c]
            dists[i, :c] = od[:c]
        return nodes, dists

    def export(self):
        """-> (from, to, level, distance) of every edge, grouped by (from, level), neighbours ascending."""
        e = int(lib().orc_hnsw_export(self._h, None, None, None, None))
        fr, to = np.empty(e, dtype="<u4"), np.empty(e, dtype="<u4")
        lv, ds = np.empty(e, dtype="<i4"), np.empty(e, dtype="<f4")
        lib().orc_hnsw_export(self._h, _ptr(fr), _ptr(to), _ptr(lv), _ptr(ds))
        return fr, to, lv, ds

    def info(self):
        entry, nodes, dist, fetches = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        lvl = C.c_int32()
        hist = (C.c_uint64 * 5)()
        lib().orc_hnsw_info(self._h, C.byref(entry), C.byref(lvl), C.byref(nodes), C.byref(dist), hist, C.byref(fetches))
        return dict(entry=entry.value, entry_level=lvl.value, nodes=nodes.value, distances=dist.value,
                    batch_hist={"1-4": hist[0], "5-16": hist[1], "17-32": hist[2], "33-64": hist[3], "65+": hist[4]}, fetches=fetches.value)
