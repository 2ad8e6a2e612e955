"""compute-sanitizer is not available on the GPU pool, so the library has a self-checking build
(libad4cl_cuda_dbg.so, -DAD4CL_DEBUG_BOUNDS=1): the recorders check every push against the thread's own
column, and both sweeps validate every id word / cell offset they are about to follow (operand precedes its
use, independent exists, offset names the row's own cell or a unit cell, far rows have a plane); a violation is
reported instead of executed.  This test runs
examples/sanity_small.py (every device code path) against that build and requires identical output."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_bounds_checked_build_runs_every_path_clean():
    dbg = os.path.join(REPO, "ad4cl_b200", "libad4cl_cuda_dbg.so")
    if not os.path.exists(dbg):
        pytest.skip("libad4cl_cuda_dbg.so not built (make -C ad4cl_b200/csrc debug)")
    prog = os.path.join(REPO, "examples", "sanity_small.py")
    plain = subprocess.run([sys.executable, prog], capture_output=True, text=True, timeout=300)
    assert plain.returncode == 0, plain.stderr[-2000:]
    env = dict(os.environ, AD4CL_CUDA_LIB=dbg)
    checked = subprocess.run([sys.executable, prog], capture_output=True, text=True, timeout=300, env=env)
    assert checked.returncode == 0, checked.stderr[-2000:]      # a failed bounds check raises Ad4clError
    import re
    num = re.compile(r"-?\d+\.\d+(?:e[-+]?\d+)?")
    a, b = plain.stdout.splitlines(), checked.stdout.splitlines()
    assert len(a) == len(b) > 10
    for la, lb in zip(a, b):
        assert num.sub("#", la) == num.sub("#", lb)
        for x, y in zip(num.findall(la), num.findall(lb)):      # fp64 atomics (acc = global) are not order-deterministic
            assert abs(float(x) - float(y)) <= 1e-12 * max(1.0, abs(float(x)))
