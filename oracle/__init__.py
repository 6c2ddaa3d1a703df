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

F32, I8, BIT = 0, 1, 2
L2, L1, COSINE, HAMMING = 0, 1, 2, 3
_NP = {F32: np.dtype("<f4"), I8: np.dtype("i1"), BIT: np.dtype("u1")}

_lib = None


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc only)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "vecgpu_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
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
        _lib = L
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
