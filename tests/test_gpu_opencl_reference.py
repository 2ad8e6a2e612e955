"""The reference's UNMODIFIED OpenCL path, run on the GPU box's own OpenCL platform
(oracle/opencl_ref.py: ad.cl + kernel text built with clBuildProgram, the simple.cpp launch /
read-back protocol, the reference's ad4cl.h host sweep), against

  (1) the committed golden vectors -- which were produced by the same reference sources compiled
      as C on the CPU (oracle/_ref): the two transports of the reference must tell the same story,
      which pins the CPU oracle to a real OpenCL run of the reference;
  (2) the CUDA path through the C ABI on the same inputs, to the north star's 1e-12.

Skipped when the box has no OpenCL GPU platform or oracle/_ref/cl was not built.
"""
import json
import os
import sys

import numpy as np
import pytest

from ad4cl_b200 import synth
from ad4cl_b200 import workloads as W

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
RTOL = 1e-12


def rel_err(x, ref):
    x, ref = np.atleast_1d(x), np.atleast_1d(ref)
    scale = np.maximum(np.abs(ref), np.abs(ref).max())
    return float(np.max(np.abs(x - ref) / np.maximum(scale, 1e-300)))


@pytest.fixture(scope="module")
def opencl():
    import opencl_ref
    if not opencl_ref.available():
        pytest.skip("no OpenCL GPU platform (or oracle/_ref/cl not built)")
    return {True: opencl_ref.OpenCLReference(True), False: opencl_ref.OpenCLReference(False)}


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "ref_outputs.json")))


def unhex(v):
    return np.array([float.fromhex(s) for s in (v if isinstance(v, list) else [v])])


@pytest.mark.parametrize("quirks", [True, False], ids=["quirks", "fixed"])
@pytest.mark.parametrize("variant,nm", [(0, "kernel_cl"), (1, "simple_cl")])
def test_reference_example_kernels_on_opencl_match_golden_and_cuda(ctx, opencl, golden, variant, nm, quirks):
    """kernel.cl / test/admb/simple.cl exactly as the reference ships and launches them (config 1)."""
    x, y, (a, b) = synth.main_cpp_recipe(1000)
    r = opencl[quirks].eval_simple(variant, a, b, x, y)
    g = golden[f"{'quirks' if quirks else 'fixed'}/ref_{nm}_main_1000"]
    assert r["entries"] == g["entries"]
    assert rel_err(r["f"], unhex(g["f"])) <= RTOL
    assert rel_err(r["grad"], unhex(g["grad"])) <= RTOL
    w = W.describe("linreg2_main", 1000)
    k = W.bind(ctx, w, quirks=quirks)
    try:
        f, grad = k.eval(w.theta)
    finally:
        k.free()
    assert rel_err(f, r["f"]) <= RTOL
    assert rel_err(grad, r["grad"]) <= RTOL


CASES = [
    ("linreg2_main", 4099, {}, "linreg2_p_main_4099"),
    ("linreg2_simple", 5003, {}, "linreg2_p_simple_5003"),
    ("regression_linear", 3001, {}, "regression_linear_3001"),
    ("regression_logistic", 3001, {}, "regression_logistic_3001"),
    ("catch_at_age", 80000, {}, "catch_at_age_80000"),
    ("tape_stress", 777, {"steps": 33}, "tape_stress_33"),
    ("tape_stress", 777, {"steps": 200}, "tape_stress_200"),
    ("op_zoo", 513, {}, "op_zoo_513"),
]


@pytest.mark.parametrize("quirks", [True, False], ids=["quirks", "fixed"])
@pytest.mark.parametrize("name,n,kw,gold", CASES, ids=[c[3] for c in CASES])
def test_objectives_on_the_reference_opencl_runtime(ctx, opencl, golden, name, n, kw, gold, quirks):
    """The same kernel TEXT (ad4cl_b200/kernels/objectives.cl) on the reference's OpenCL runtime and on
    the CUDA runtime: same f, same gradient, same number of recorded operators."""
    w = W.describe(name, n, **kw)
    epi = 1 if w.epilogue == "half_n_log" else 0
    r = opencl[quirks].eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1, epilogue=epi,
                                      ops_per_item=w.max_ops)
    g = golden[f"{'quirks' if quirks else 'fixed'}/{gold}"]
    assert r["entries"] == g["entries"]
    assert rel_err(r["f"], unhex(g["f"])) <= RTOL, (r["f"], unhex(g["f"]))
    assert rel_err(r["grad"], unhex(g["grad"])) <= RTOL
    k = W.bind(ctx, w, quirks=quirks)
    try:
        f, grad = k.eval(w.theta)
        st = k.stats()
    finally:
        k.free()
    assert rel_err(f, r["f"]) <= RTOL, (f, r["f"])
    assert rel_err(grad, r["grad"]) <= RTOL, (grad, r["grad"])
    assert int(st["ops"]) == r["counter"]


def test_matrixmul_cl_forward_matches_the_reference_golden_file(ctx, opencl):
    """test/matrix/matrixmul.cl on OpenCL at the reference's own 256^2 case: C against the reference's
    committed device.txt/host.txt (6 significant digits as printed) and against the CUDA dense path;
    the device counter is the n^3 times + n^3 plus_eq entries of matrix_mul.cpp:272's printout minus
    the host sum."""
    n = 256
    A, B = synth.msvcrt_rand_matrices(n)
    r = opencl[True].matrix_forward(n, A, B)
    gold = np.load(os.path.join(HERE, "golden", "matrix256_host_txt.npz"))
    assert r["counter"] == int(gold["entries"]) - n * n
    printed = np.array([float("%g" % v) for v in r["C"]]).reshape(n, n)
    assert np.mean(printed == gold["C"]) > 0.999          # a last-digit rounding tie may fall either way (FMA contraction)
    assert np.max(np.abs(r["C"].reshape(n, n) - gold["C"]) / gold["C"]) < 1e-5
    bA, bB, bC = ctx.buffer(A), ctx.buffer(B), ctx.buffer(nbytes=8 * n * n)
    ws = ctx.buffer(nbytes=8 * (n * n + (n // 64) ** 2))
    ctx.matrix_product_dense(n, bA, bB, C=bC, workspace=ws)
    C = bC.download()
    for b in (bA, bB, bC, ws):
        b.free()
    assert rel_err(C, r["C"]) <= RTOL
