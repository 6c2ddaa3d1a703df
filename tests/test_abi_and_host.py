"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/vecgpu.h declares; host-side mirror logic (enums, parsing, error
mapping); and the product fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vecgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vecgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(vg):
    lib = ctypes.CDLL(vg.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libvecgpu.so does not export {n}"


def test_binding_covers_header(vg):
    from sqlite_vec_hnsw_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_product_never_references_oracle():
    # the product path must not import / link / call anything under oracle/
    pkg = os.path.join(ROOT, "sqlite-vec-hnsw_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libvecgpu_oracle" not in text and "orc_" not in text, f
    out = os.popen(f"ldd '{os.path.join(pkg, 'libvecgpu.so')}'").read()
    assert "oracle" not in out


def test_row_bytes_and_supported_pairs(vg):
    lib = vg.load_library()
    assert lib.vecgpu_row_bytes(0, 768) == 3072  # src/vector.rs:223-233
    assert lib.vecgpu_row_bytes(1, 1024) == 1024  # :236-242
    assert lib.vecgpu_row_bytes(2, 1024) == 128 and lib.vecgpu_row_bytes(2, 13) == 2  # :592-600
    ok = {(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (2, 3)}  # src/distance/mod.rs:70-77
    for e in range(3):
        for m in range(4):
            assert bool(lib.vecgpu_metric_supported(e, m)) == ((e, m) in ok)


def test_enum_parsing(vg):
    M, T = vg.DistanceMetric, vg.VectorType
    # src/distance/mod.rs:26-44
    assert M.from_str("L2") == M.L2 == M.from_str("euclidean")
    assert M.from_str("manhattan") == M.L1 == M.from_str("l1")
    assert M.from_str("Cosine") == M.Cosine and M.from_str("hamming") == M.Hamming
    assert [m.as_str() for m in M] == ["l2", "l1", "cosine", "hamming"]
    with pytest.raises(vg.InvalidDistanceMetric):
        M.from_str("dot")
    # src/vector.rs:30-46
    assert T.from_str("float") == T.Float32 == T.from_str("FLOAT32")
    assert T.from_str("binary") == T.Bit == T.from_str("bit") and T.from_str("int8") == T.Int8
    with pytest.raises(vg.InvalidVectorType):
        T.from_str("float16")
    assert int(T.Float32) == 0 and int(T.Int8) == 1 and int(T.Bit) == 2  # order of src/vector.rs:9-16


def test_vector_constructors(vg):
    v = vg.Vector.from_json("[1.0, 2.5, -3]", vg.VectorType.Float32)  # src/vector.rs:239-256
    assert v.dimensions == 3 and np.array_equal(v.as_f32(), np.array([1, 2.5, -3], dtype="<f4"))
    w = vg.Vector.from_json("[1.9, -1.9, 300, -300]", vg.VectorType.Int8)  # Rust `as i8`: truncate + saturate
    assert list(w.as_i8()) == [1, -1, 127, -128]
    with pytest.raises(vg.NotImplementedVec):
        vg.Vector.from_json("[1,0]", vg.VectorType.Bit)
    with pytest.raises(vg.InvalidVectorFormat):
        vg.Vector.from_json("[1,", vg.VectorType.Float32)
    assert vg.Vector.from_f32([1, 2]).as_bytes() == np.array([1, 2], dtype="<f4").tobytes()
    with pytest.raises(vg.InvalidVectorType):
        vg.Vector.from_i8([1, 2]).as_f32()


def test_distance_checks_before_device(vg):
    # order of checks as in src/distance/mod.rs:57-68: dimensions, then types
    a, b = vg.Vector.from_f32([1, 2, 3]), vg.Vector.from_f32([1, 2])
    with pytest.raises(vg.DimensionMismatch) as e:
        vg.distance(a, b, vg.DistanceMetric.L2)
    assert e.value.expected == 3 and e.value.actual == 2
    with pytest.raises(vg.InvalidVectorType):
        vg.distance(vg.Vector.from_f32([1, 2]), vg.Vector.from_i8([1, 2]), vg.DistanceMetric.L2)


def test_abi_argument_errors_without_device(vg):
    lib = vg.load_library()
    out = np.zeros(1, dtype="<f4")
    a = np.zeros(4, dtype="<f4")
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert lib.vecgpu_distance_pairs(0, 4, 3, p(a), p(a), 1, 0, 0, p(out)) == 2  # DimensionMismatch
    assert b"Dimension mismatch: expected 4, got 3" in lib.vecgpu_last_error()
    assert lib.vecgpu_distance_pairs(0, 4, 4, p(a), p(a), 1, 3, 0, p(out)) == 3  # f32 x hamming
    assert b"Distance metric Hamming not supported for vector type Float32" in lib.vecgpu_last_error()
    assert lib.vecgpu_distance_pairs(2, 8, 8, p(a), p(a), 1, 0, 0, p(out)) == 3  # bit x l2
    assert lib.vecgpu_distance_pairs(7, 8, 8, p(a), p(a), 1, 0, 0, p(out)) == 3  # bad type
    h = ctypes.c_void_p()
    assert lib.vecgpu_slab_create(0, 0, 0, 0, ctypes.byref(h)) == 1  # dims == 0
    assert lib.vecgpu_slab_create(9, 4, 0, 0, ctypes.byref(h)) == 3


def test_no_cpu_fallback(vg):
    """Without a device every compute entry point fails loudly (VECGPU_ERR_CUDA -> InvalidState)."""
    if vg.load_library().vecgpu_device_count() > 0:
        pytest.skip("a GPU is visible; the loud-failure path is for device-less hosts")
    with pytest.raises(vg.InvalidState):
        vg.Slab(vg.VectorType.Float32, 8)
    with pytest.raises(vg.InvalidState):
        vg.distance(vg.Vector.from_f32([1, 2]), vg.Vector.from_f32([3, 4]), vg.DistanceMetric.L2)
    with pytest.raises(vg.InvalidState):
        vg.quantize_int8(np.zeros((1, 4), dtype="<f4"))


def test_hnsw_metric_rule(vg):
    # src/hnsw/mod.rs:129-146
    M = vg.DistanceMetric
    assert vg.internal_distance_metric(M.Cosine, True) == M.L2
    assert vg.internal_distance_metric(M.Cosine, False) == M.Cosine
    assert vg.internal_distance_metric(M.L1, True) == M.L1
    assert vg.convert_distance_for_output(M.Cosine, True, 1.0) == 0.5
    assert vg.convert_distance_for_output(M.L2, True, 1.5) == 1.5
