"""Differential fuzz (tools/fuzz_paths.py) as part of the GPU suite: ~10 s of random shapes, every fast path against the
plain exact scan / the lockstep HNSW driver."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_fast_paths_agree_with_the_exact_scan(gpu, seed):
    env = {k: v for k, v in os.environ.items() if not k.startswith("VECGPU_")}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_paths.py"), "10", str(seed)], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
