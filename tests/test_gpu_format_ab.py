"""A/B of the two tape stream formats (ad4cl_tape.cuh).  libad4cl_cuda.so first tries the uniform
recorder (ONE 8-byte header per warp and row, cells only for non-unit partials, warp-synchronous sweep);
libad4cl_cuda_general.so (-DAD4CL_FORCE_GENERAL=1) records every row in the per-lane format and sweeps
with the general loop.  Per lane the tape arithmetic is identical, so with per-thread accumulators
(P <= 16) objective and gradient must agree BIT FOR BIT; with per-warp accumulators only the order in
which an independent's contributions are summed over the warp differs (pairs in the uniform sweep): 1e-13.
Operator and slot counts must be equal."""
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
        fa, fb = la.split(), lb.split()
        assert fa[:2] == fb[:2] and fa[-2:] == fb[-2:], (la[:120], lb[:120])      # workload, n, ops, slots
        if fa[2] == "exact":
            assert la == lb, (la[:200], lb[:200])
        else:
            xa = [float.fromhex(v) for v in fa[3:-2]]
            xb = [float.fromhex(v) for v in fb[3:-2]]
            scale = max(abs(v) for v in xb[1:]) if len(xb) > 1 else 1.0
            assert abs(xa[0] - xb[0]) <= 1e-13 * abs(xb[0])
            assert max(abs(u - v) for u, v in zip(xa[1:], xb[1:])) <= 1e-13 * scale
