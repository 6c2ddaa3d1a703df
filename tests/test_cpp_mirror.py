"""The C++ host mirror (include/vecgpu.hpp) restates the reference's Rust unit tests in tests/cpp/test_mirror.cpp.
The binary is built by __graft_entry__.build() (make -C tests/cpp) and linked against the in-tree libvecgpu.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_mirror")


def _run():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    return subprocess.run([BIN], capture_output=True, text=True, timeout=120)


def test_cpp_mirror_host_logic_and_loud_failure(vg):
    """Without a device: enum parsing, check order of distance(), and InvalidState from every compute call."""
    if vg.load_library().vecgpu_device_count() > 0:
        pytest.skip("a GPU is visible; covered by the gpu-marked run")
    r = _run()
    assert r.returncode == 0 and "NO_DEVICE_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_reference_unit_tests(gpu):
    r = _run()
    assert r.returncode == 0 and "ALL_PASSED" in r.stdout, r.stdout + r.stderr
