import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The product routes small batches (fewer than 64 queries and fewer than ~2 M (query, row) pairs) to the exact CUDA-core multi-query scan because
# the tensor-core launch has a fixed cost.  The parity tests use small tables on purpose, so they lower the threshold:
# every batch of >= 16 queries over >= 8192 rows then exercises the tensor-core kernels (both routes are compared with
# the oracle anyway; test_small_batches_take_the_cuda_core_route checks the default routing).
os.environ.setdefault("VECGPU_TC_MIN_WORK", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle

    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def vg():
    """The product package; loads libvecgpu.so or fails loudly."""
    import sqlite_vec_hnsw_b200 as v

    v.load_library()
    return v


@pytest.fixture(scope="session")
def gpu(vg):
    n = vg.load_library().vecgpu_device_count()
    if n <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return n
