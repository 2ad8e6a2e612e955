"""Operator-level parity on the GPU: every operator of every family of ad.cuh (31 x 3)
against what the reference's ad.cl records for the same operands (golden op probes,
generated from the reference itself), plus the chains through the pad_ and private
families and the edge cases of the launch shape."""
import json
import os

import numpy as np
import pytest

from ad4cl_b200 import synth, workloads as W

pytestmark = pytest.mark.gpu
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(REPO, "tests", "golden")
OPS = ["plus", "plus_vd", "plus_dv", "minus", "minus_vd", "minus_dv", "times", "times_vd", "times_dv",
       "divide", "divide_vd", "divide_dv", "plus_eq", "plus_eq_d", "cos", "sin", "tan", "acos", "asin", "atan",
       "cosh", "sinh", "tanh", "exp", "log", "log10", "sqrt", "pow", "pow_vd", "pow_dv"]
POINTS = [(0.3, 1.7), (0.81, 2.25), (0.05, 1.01)]
RTOL_LIBM = 1e-14       # CUDA libm vs glibc: <= a few ulp per transcendental (north_star tolerance is 1e-12)


def close(x, ref):
    return abs(x - ref) <= RTOL_LIBM * max(abs(ref), 1e-300)


@pytest.mark.parametrize("quirks", [True, False], ids=["quirks", "fixed"])
@pytest.mark.parametrize("family", ["global", "pad", "private"])
def test_every_operator_records_what_the_reference_records(ctx, family, quirks):
    from ad4cl_b200 import host
    gold = json.load(open(os.path.join(GOLD, "op_probes.json")))
    mode = "quirks" if quirks else "fixed"
    n = len(OPS) * len(POINTS)
    d0 = ctx.buffer(np.array(POINTS, dtype=np.float64).ravel())
    d1 = ctx.buffer(nbytes=8 * 8 * n)
    k = ctx.builtin("op_probe", quirks=quirks)
    k.set_args(None, None, host.params(0, 2), d0, d1, None, n, {"global": 0, "pad": 1, "private": 2}[family], 0)
    k.configure(n, max_ops=8, max_slots=8, local_size=32)
    k.eval([0.0, 0.0])
    rec = d1.download().reshape(n, 8)
    k.free()
    checked = 0
    for pi, (a, b) in enumerate(POINTS):
        for oi, op in enumerate(OPS):
            r = rec[pi * len(OPS) + oi]
            g = gold[f"{mode}/{family}/{op}/{a}/{b}"]
            if g is None:      # pad_ math: absent from the reference (Q12) -> same table row as the global family
                assert family == "pad"
                g = gold[f"{mode}/global/{op}/{a}/{b}"]
            assert close(r[0], float.fromhex(g["value"])), (family, op, r[0])
            assert int(r[1]) == g["size"], (family, op)
            for s in range(g["size"]):
                assert close(r[2 + 2 * s], float.fromhex(g["dx"][s])), (family, op, s, r[2 + 2 * s], g["dx"][s])
                assert int(r[3 + 2 * s]) == {7: 0, 9: 1}[g["ids"][s]], (family, op, s)
            assert r[6] == 1.0                      # the result is a temporary of this thread's tape
            checked += 1
    assert checked == 90


@pytest.mark.parametrize("kernel,quirks", [("op_zoo_pad", False), ("op_zoo_private", False), ("op_zoo_private", True)])
def test_chains_through_pad_and_private_families(ctx, oracle_api, kernel, quirks):
    """Same operator chain as op_zoo; the oracle runs it through the reference's global family
    (identical table rows; in quirks mode the pad family differs only by times_dv, ad.cl:840)."""
    from ad4cl_b200 import host
    w = W.describe("op_zoo", 513)
    ref = oracle_api.best(quirks).eval_objective("op_zoo", w.theta, w.d0, None, w.size, ops_per_item=48)
    k = ctx.builtin(kernel, quirks=quirks)
    k.set_args(None, None, host.params(0, 3), ctx.buffer(w.d0), None, None, w.size, 0, 0)
    k.configure(w.size, max_ops=48, max_slots=96)
    f, g = k.eval(w.theta)
    st = k.stats()
    k.free()
    assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))
    if kernel == "op_zoo_private":
        assert st["slots"] == 5 * w.size        # pre-accumulated: 1 + 2 + 2 slots instead of 64


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 1000])
@pytest.mark.parametrize("local", [32, 64, 256])
def test_ragged_sizes(ctx, oracle_api, n, local):
    w = W.describe("regression_logistic", n)
    ref = oracle_api.best(False).eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1, ops_per_item=w.max_ops)
    k = W.bind(ctx, w, local_size=local)
    f, g = k.eval(w.theta)
    k.free()
    assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))


def test_long_tape_small_grid(ctx, oracle_api):
    """10^4 operators per work-item (config 5's upper end) on a handful of work-items."""
    w = W.describe("tape_stress", 97, steps=3333)
    ref = oracle_api.best(False).eval_objective(w.kernel, w.theta, w.d0, None, w.size, w.i0, w.i1, ops_per_item=w.max_ops)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    st = k.stats()
    k.free()
    assert int(st["ops"]) == 97 * 10000 and int(st["slots"]) == 97 * 13333
    assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))


def test_matrix_ragged_and_invalid_shapes(ctx, oracle_api):
    from ad4cl_b200 import host
    n = 20                                               # not a multiple of the 16 x 16 work-group
    w = W.describe("matrix_product", n)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    k.free()
    A, B = w.theta[: n * n].reshape(n, n), w.theta[n * n:].reshape(n, n)
    assert abs(f - (A @ B).sum()) <= 1e-12 * abs(f)
    gA = np.repeat(B.sum(axis=1)[None, :], n, axis=0).ravel()
    gB = np.repeat(A.sum(axis=0)[:, None], n, axis=1).ravel()
    assert np.max(np.abs(g - np.concatenate([gA, gB]))) <= 1e-12 * np.abs(g).max()
    k = ctx.builtin("regression_linear")
    with pytest.raises(host.Ad4clError) as e:
        k.configure(0, max_ops=4)                         # empty NDRange: invalid, as in OpenCL
    assert e.value.code == host.ERR_INVALID
    with pytest.raises(host.Ad4clError) as e:
        k.configure(100, max_ops=4, local_size=48)        # work-group must be whole warps
        k.set_args(None, None, host.params(0, 1), None, None, None, 100, 1, 0)
        k.eval([1.0])
    assert e.value.code == host.ERR_INVALID
    with pytest.raises(host.Ad4clError) as e:             # unset argument is reported, not undefined behaviour
        k2 = ctx.builtin("regression_linear")
        k2.configure(100, max_ops=4)
        k2.eval([1.0])
    assert e.value.code == host.ERR_INVALID and "not set" in str(e.value)
    k.free()


def test_var_operators_are_the_function_calls(ctx):
    """ad4cl::var (device-side Variable.hpp): objectives written with operators record the same tape
    and return the same bits as their ad_* function-call forms."""
    from ad4cl_b200 import host
    w = W.describe("linreg2_main", 10007)
    res = {}
    for name in ("linreg2_g", "linreg2_var"):
        k = ctx.builtin(name)
        k.set_args(None, None, host.params(0), host.params(1), ctx.buffer(w.d0), ctx.buffer(w.d1), None, w.size)
        k.configure(w.size, max_ops=4, max_slots=6, epilogue="half_n_log")
        res[name] = k.eval(w.theta) + (k.stats()["slots"],)
        k.free()
    assert res["linreg2_g"][0] == res["linreg2_var"][0] and np.array_equal(res["linreg2_g"][1], res["linreg2_var"][1])
    assert res["linreg2_g"][2] == res["linreg2_var"][2]
    w = W.describe("regression_logistic", 3001)
    out = []
    for name in ("regression_logistic", "regression_logistic_var"):
        w.kernel = name
        k = W.bind(ctx, w, tape="hbm")      # the function-call form also has a register-tier instance: compare like with like
        out.append(k.eval(w.theta))
        k.free()
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


def test_var_kernel_through_the_jit(ctx):
    """#include "ad4cl_var.cuh" is available to JIT-compiled kernels as well."""
    from ad4cl_b200 import host
    src = '#include "ad4cl_var.cuh"\n' + open(os.path.join(REPO, "ad4cl_b200", "kernels", "var_kernels.cl")).read()
    k = ctx.from_source(src, "linreg2_var", "GSVVDDOi")
    w = W.describe("linreg2_main", 1000)
    k.set_args(None, None, host.params(0), host.params(1), ctx.buffer(w.d0), ctx.buffer(w.d1), None, w.size)
    k.configure(w.size, max_ops=4, max_slots=6, epilogue="half_n_log")
    f, g = k.eval(w.theta)
    k.free()
    kb = W.bind(ctx, w)
    fb, gb = kb.eval(w.theta)
    kb.free()
    assert f == fb and np.array_equal(g, gb)


def test_config1_on_a_private_stack(ctx, oracle_api):
    """linreg2_private: simple.cl's four operators through the *_p family + ad_commit_p; same f and
    gradient as the reference's evaluation of the pad-family kernel, 3 slots per work-item on the tape."""
    w = W.describe("linreg2_private", 100003)
    ref = oracle_api.best(False).eval_objective("linreg2_p", w.theta, w.d0, w.d1, w.size, epilogue=1, ops_per_item=4)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    st = k.stats()
    k.free()
    assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))
    assert int(st["slots"]) == 3 * w.size and int(st["ops"]) == 2 * w.size
