"""A/B of the two tape stream formats (ad4cl_tape.cuh).  libad4cl_cuda.so records a row ONCE per warp
when all 32 lanes push the same id word (8-byte header + cells only for non-unit partials) and sweeps such
work-items warp-synchronously; libad4cl_cuda_general.so (-DAD4CL_FORCE_GENERAL=1) records every row in
the per-lane format and sweeps with the general loop.  Both perform the same fp64 operations in the same
order per lane, so objective and gradient must agree BIT FOR BIT, and so must the operator / slot counts."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def run(lib):
    env = dict(os.environ, AD4CL_CUDA_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "format_ab_worker.py")], capture_output=True,
                       text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout.splitlines()


def test_uniform_and_general_stream_formats_agree_bit_for_bit():
    std = os.path.join(REPO, "ad4cl_b200", "libad4cl_cuda.so")
    gen = os.path.join(REPO, "ad4cl_b200", "libad4cl_cuda_general.so")
    if not os.path.exists(gen):
        pytest.skip("libad4cl_cuda_general.so not built (make -C ad4cl_b200/csrc general)")
    a, b = run(std), run(gen)
    assert len(a) == len(b) >= 8
    for la, lb in zip(a, b):
        assert la == lb, (la[:200], lb[:200])
