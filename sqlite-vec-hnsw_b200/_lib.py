"""ctypes binding of libvecgpu.so (include/vecgpu.h).  Loading fails loudly:
there is no Python / CPU fallback for any compute entry point."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvecgpu.so")

F32, I8, BIT = 0, 1, 2
L2, L1, COSINE, HAMMING = 0, 1, 2, 3
OK, ERR_INVALID_PARAM, ERR_DIM_MISMATCH, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3, 4
SYNTH_UNIFORM, SYNTH_GAUSS4 = 0, 1
XCHG_HANDLE_BYTES = 128

_c_slab = C.c_void_p
_p = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/vecgpu.h declares
SIGNATURES = {
    "vecgpu_last_error": (C.c_char_p, []),
    "vecgpu_version": (C.c_char_p, []),
    "vecgpu_device_count": (C.c_int, []),
    "vecgpu_row_bytes": (C.c_uint32, [C.c_int, C.c_uint32]),
    "vecgpu_metric_supported": (C.c_int, [C.c_int, C.c_int]),
    "vecgpu_slab_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_c_slab)]),
    "vecgpu_slab_destroy": (None, [_c_slab]),
    "vecgpu_slab_load": (C.c_int, [_c_slab, _p, _p, C.c_uint64]),
    "vecgpu_slab_append": (C.c_int, [_c_slab, _p, _p, C.c_uint64]),
    "vecgpu_slab_upsert": (C.c_int, [_c_slab, C.c_int64, _p, C.c_uint32]),
    "vecgpu_slab_delete": (C.c_int, [_c_slab, C.c_int64]),
    "vecgpu_slab_compact": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "vecgpu_slab_count": (C.c_int, [_c_slab, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vecgpu_slab_get": (C.c_int, [_c_slab, C.c_int64, _p, C.POINTER(C.c_int)]),
    "vecgpu_knn": (C.c_int, [_c_slab, _p, C.c_uint32, C.c_uint32, C.c_int, _p, _p, _p]),
    "vecgpu_score": (C.c_int, [_c_slab, _p, C.c_uint32, _p, _p, C.c_int, _p]),
    "vecgpu_distance_pairs": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, _p, _p, C.c_uint64, C.c_int, C.c_int, _p]),
    "vecgpu_normalize_f32": (C.c_int, [_p, C.c_uint64, C.c_uint32, C.c_int, _p]),
    "vecgpu_quantize_int8": (C.c_int, [_p, C.c_uint64, C.c_uint32, C.c_int, _p]),
    "vecgpu_quantize_int8_for_index": (C.c_int, [_p, C.c_uint64, C.c_uint32, C.c_int, _p]),
    "vecgpu_quantize_binary": (C.c_int, [_p, C.c_uint64, C.c_uint32, C.c_int, _p]),
    "vecgpu_slab_fill_synthetic": (C.c_int, [_c_slab, C.c_uint64, C.c_int64, C.c_uint64, C.c_int]),
    "vecgpu_slab_device_view": (C.c_int, [_c_slab, C.POINTER(_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "vecgpu_knn_device": (C.c_int, [_c_slab, _p, C.c_uint32, C.c_uint32, C.c_int, _p, _p, _p]),
    "vecgpu_merge_device": (C.c_int, [C.c_int, _p, _p, C.c_uint32, C.c_uint32, C.c_uint32, _p, _p, _p]),
    "vecgpu_hnsw_stored_slab": (C.c_int, [_c_slab, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vecgpu_hnsw_create": (C.c_int, [_c_slab, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "vecgpu_hnsw_destroy": (None, [C.c_void_p]),
    "vecgpu_hnsw_build": (C.c_int, [C.c_void_p, C.c_uint32]),
    "vecgpu_hnsw_insert_appended": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "vecgpu_hnsw_reinsert": (C.c_int, [C.c_void_p, C.c_int64]),
    "vecgpu_hnsw_insert_at": (C.c_int, [C.c_void_p, C.c_int64]),
    "vecgpu_hnsw_search": (C.c_int, [C.c_void_p, _p, C.c_uint32, C.c_uint32, C.c_uint32, _p, _p, _p]),
    "vecgpu_hnsw_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vecgpu_hnsw_entry_point": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "vecgpu_hnsw_device_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vecgpu_hnsw_batch_histogram": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "vecgpu_hnsw_export_nodes": (C.c_int, [C.c_void_p, C.c_uint64, _p, _p, C.POINTER(C.c_uint64)]),
    "vecgpu_hnsw_export_edges": (C.c_int, [C.c_void_p, C.c_uint64, _p, _p, _p, _p, C.POINTER(C.c_uint64)]),
    "vecgpu_sharded_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, _p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "vecgpu_sharded_destroy": (None, [C.c_void_p]),
    "vecgpu_sharded_load": (C.c_int, [C.c_void_p, _p, _p, C.c_uint64]),
    "vecgpu_sharded_fill_synthetic": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_int]),
    "vecgpu_sharded_upsert": (C.c_int, [C.c_void_p, C.c_int64, _p, C.c_uint32]),
    "vecgpu_sharded_delete": (C.c_int, [C.c_void_p, C.c_int64]),
    "vecgpu_sharded_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vecgpu_sharded_num_shards": (C.c_uint32, [C.c_void_p]),
    "vecgpu_sharded_shard": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "vecgpu_sharded_knn": (C.c_int, [C.c_void_p, _p, C.c_uint32, C.c_uint32, C.c_int, _p, _p, _p]),
    "vecgpu_xchg_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "vecgpu_xchg_destroy": (None, [C.c_void_p]),
    "vecgpu_xchg_export": (C.c_int, [C.c_void_p, _p]),
    "vecgpu_xchg_attach_ipc": (C.c_int, [C.c_void_p, _p, C.c_uint32]),
    "vecgpu_xchg_attach_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "vecgpu_shard_knn": (C.c_int, [_c_slab, C.c_void_p, _p, C.c_uint32, C.c_uint32, C.c_int, _p, _p, _p]),
    "vecgpu_shard_knn_device": (C.c_int, [_c_slab, C.c_void_p, _p, C.c_uint32, C.c_uint32, C.c_int, _p, _p, _p]),
    "vecgpu_xchg_merge_device": (C.c_int, [C.c_void_p, _p, _p, _p, C.c_uint32, C.c_uint32, _p, _p, _p]),
    "vecgpu_xchg_check": (C.c_int, [C.c_void_p, _p]),
    "vecgpu_debug_scan_timeline": (None, [_p]),
    "vecgpu_launch_count": (C.c_uint64, []),
    "vecgpu_tc_stats": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load libvecgpu.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C sqlite-vec-hnsw_b200/csrc).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().vecgpu_last_error().decode("utf-8", "replace")
