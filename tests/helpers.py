"""Shared case generators for the parity tests (seeded, small)."""
import numpy as np

F32, I8, BIT = 0, 1, 2
L2, L1, COSINE, HAMMING = 0, 1, 2, 3
PAIRS = [(F32, L2), (F32, L1), (F32, COSINE), (I8, L2), (I8, L1), (I8, COSINE), (BIT, HAMMING)]
PAIR_IDS = ["f32-l2", "f32-l1", "f32-cos", "i8-l2", "i8-l1", "i8-cos", "bit-hamming"]
NP = {F32: np.dtype("<f4"), I8: np.dtype("i1"), BIT: np.dtype("u1")}


def row_bytes(elem, dims):
    return (dims * 4, dims, (dims + 7) // 8)[elem]


def random_rows(elem, n, dims, seed, ties=False):
    """n rows of the given type.  ties=True draws from a tiny alphabet so equal distances are frequent."""
    rng = np.random.default_rng(seed)
    if elem == F32:
        if ties:
            return rng.integers(-2, 3, size=(n, dims)).astype("<f4")
        return rng.standard_normal((n, dims)).astype("<f4")
    if elem == I8:
        if ties:
            return rng.integers(-2, 3, size=(n, dims)).astype("i1")
        return rng.integers(-128, 128, size=(n, dims)).astype("i1")
    nb = (dims + 7) // 8
    rows = rng.integers(0, 256, size=(n, nb)).astype("u1")
    if dims % 8:
        rows[:, -1] &= (1 << (dims % 8)) - 1
    if ties:
        rows[:, : max(0, nb - 1)] = 0
    return rows


def rel_close(a, b, tol=1e-5):
    """|a-b| <= tol*max(|a|,|b|) elementwise, inf==inf."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b))
    return bool(np.all(ok | both_inf | both_nan))


def same_bits(a, b):
    """float32 arrays equal bit for bit; NaNs only need to coincide (payloads may differ)."""
    a = np.ascontiguousarray(a, dtype="<f4")
    b = np.ascontiguousarray(b, dtype="<f4")
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view("<u4")[~na], b.view("<u4")[~nb]))
