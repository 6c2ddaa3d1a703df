"""sqlite-vec-hnsw_b200 — B200-native distance scoring for sqlite-vec-hnsw.

Only the hot path of SURVEY.md §8: exact KNN scan + HNSW candidate scoring,
behind the C ABI of include/vecgpu.h (libvecgpu.so, built from csrc/).
Import as ``sqlite_vec_hnsw_b200`` (alias module at the repo root).
"""
from . import _lib
from .vec0 import (  # noqa: F401
    DimensionMismatch,
    DistanceMetric,
    Exchange,
    HnswIndex,
    InvalidDistanceMetric,
    InvalidParameter,
    InvalidState,
    InvalidVectorFormat,
    InvalidVectorType,
    NotImplementedVec,
    ShardedSlab,
    Slab,
    VecError,
    Vector,
    VectorType,
    brute_force_search,
    convert_distance_for_output,
    distance,
    distance_pairs,
    internal_distance_metric,
    merge_device,
    normalize,
    quantize_binary,
    quantize_int8,
    quantize_int8_for_index,
)

LIB_PATH = _lib.LIB_PATH


def load_library():
    """Load libvecgpu.so now (raises if it has not been built)."""
    return _lib.load()


def version():
    return _lib.load().vecgpu_version().decode()


def launch_count():
    return int(_lib.load().vecgpu_launch_count())


def tc_stats():
    """(queries served by the tensor-core batched path, of which fell back to the exact scan)."""
    import ctypes as C

    q, f = C.c_uint64(), C.c_uint64()
    _lib.load().vecgpu_tc_stats(C.byref(q), C.byref(f))
    return int(q.value), int(f.value)
