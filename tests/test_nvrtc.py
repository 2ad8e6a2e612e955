"""The JIT path (NVRTC): kernel text written against ad.cl is compiled at run time, the
way the reference builds ad.cl + user file (test/admb/simple.cpp:138-175).

CPU part: compilation only (NVRTC needs no GPU) -- including the reference's OWN
simple.cl and matrixmul.cl, byte-for-byte unmodified, read from /root/reference.
GPU part: the JIT-built kernels run and match the oracle; the reference's kernels are
pre-compiled by __graft_entry__.build() into oracle/_ref/*.cubin (the sources cannot
travel to the GPU box) and loaded with ad4cl_cuda_kernel_from_cubin.
"""
import os

import numpy as np
import pytest

from ad4cl_b200 import synth, workloads as W

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"
CUBINS = os.path.join(REPO, "oracle", "_ref")
RTOL = 1e-12


def rel_err(x, ref):
    x, ref = np.atleast_1d(x), np.atleast_1d(ref)
    return float(np.max(np.abs(x - ref) / np.maximum(np.maximum(np.abs(ref), np.abs(ref).max()), 1e-300)))


def test_reference_kernels_compile_unmodified():
    if not os.path.exists(os.path.join(REF, "ad.cl")):
        pytest.skip("reference tree not mounted")
    from ad4cl_b200 import host
    for path, entry, sig in (("test/admb/simple.cl", "AD", "GSVVDDOi"),
                             ("test/matrix/matrixmul.cl", "matrixMult", "GSVVOii")):
        text = open(os.path.join(REF, path)).read()
        for quirks in (True, False):
            cubin = host.compile_source(text, entry, sig, quirks=quirks)
            assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000


def test_objectives_compile_from_source_and_errors_carry_the_log():
    from ad4cl_b200 import host
    text = open(os.path.join(REPO, "ad4cl_b200", "kernels", "objectives.cl")).read()
    assert host.compile_source(text, "catch_at_age", "GSVDDOiii")[:4] == b"\x7fELF"
    with pytest.raises(host.Ad4clError) as e:
        host.compile_source("__kernel void K(__global struct ad_gradient_structure* gs) { undeclared_fn(gs); }", "K", "GO")
    assert e.value.code == host.ERR_COMPILE and "undeclared_fn" in str(e.value)
    with pytest.raises(host.Ad4clError) as e:
        host.compile_source(text, "catch_at_age", "GSVDDiii")      # no O in the signature
    assert e.value.code == host.ERR_INVALID


@pytest.mark.gpu
def test_jit_kernel_matches_builtin_and_oracle(ctx, oracle_api):
    from ad4cl_b200 import host
    text = open(os.path.join(REPO, "ad4cl_b200", "kernels", "objectives.cl")).read()
    w = W.describe("regression_logistic", 3001)
    ref = oracle_api.best(False).eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1, ops_per_item=w.max_ops)
    kb = W.bind(ctx, w, tape="hbm")     # the built-in also has a register-tier instance; compare the same tier
    fb, gb = kb.eval(w.theta)
    kb.free()
    kj = ctx.from_source(text, "regression_logistic", "GSVDDOiii")
    d0, d1 = ctx.buffer(w.d0), ctx.buffer(w.d1)
    kj.set_args(None, None, host.params(0, w.nparams), d0, d1, None, w.size, w.i0, w.i1)
    kj.configure(w.size, max_ops=w.max_ops, max_slots=w.max_slots, tape="hbm")
    fj, gj = kj.eval(w.theta)
    kj.free()
    assert rel_err(fj, ref["f"]) <= RTOL and rel_err(gj, ref["grad"]) <= RTOL
    assert fj == fb and np.array_equal(gj, gb), "JIT and ahead-of-time builds of the same text disagree"


@pytest.mark.gpu
@pytest.mark.parametrize("quirks", [True, False], ids=["quirks", "fixed"])
def test_reference_simple_cl_runs_on_the_cuda_runtime(ctx, oracle_api, quirks):
    """test/admb/simple.cl, unmodified, on the main.cpp recipe; checked against the
    reference's own evaluation of the same file on the CPU."""
    from ad4cl_b200 import host
    path = os.path.join(CUBINS, f"ref_simple_cl.{'q' if quirks else 'f'}.{host.device_abi()}.cubin")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_simple_cl.*.cubin not built for this library (run __graft_entry__.build() where the reference is mounted)")
    n = 100003
    x, y, (a, b) = synth.main_cpp_recipe(n)
    k = ctx.from_cubin(open(path, "rb").read(), "AD", "GSVVDDOi", quirks=quirks)
    k.set_args(None, None, host.params(0), host.params(1), ctx.buffer(x), ctx.buffer(y), None, n)
    k.configure(n, max_ops=4, max_slots=6, local_size=64, epilogue="half_n_log")     # local 64: simple.cpp:235
    f, g = k.eval([a, b])
    ops = k.stats()["ops"]
    k.free()
    if oracle_api.have_reference():
        ref = oracle_api.Reference(quirks).eval_simple(1, a, b, x, y, global_size=((n + 63) // 64 + 1) * 64)
    else:
        ref = oracle_api.Port(quirks).eval_objective("linreg2_p", [a, b], x, y, n, epilogue=1, ops_per_item=4)
    assert rel_err(f, ref["f"]) <= RTOL and rel_err(g, ref["grad"]) <= RTOL
    assert ops == 4 * n == ref["entries"] - n - 2


def _matrix_case(ctx, oracle_api, kernel, n, quirks, acc):
    from ad4cl_b200 import host
    A, B = synth.msvcrt_rand_matrices(n)
    kernel.set_args(None, None, host.params(0, n * n), host.params(n * n, n * n), None, n, n)
    kernel.configure((n, n), max_ops=2 * n + 1, max_slots=4 * n + 1, local_size=(16, 16), acc=acc)
    f, g = kernel.eval(np.concatenate([A, B]))
    st = kernel.stats()
    kernel.free()
    orc = oracle_api.Reference(quirks) if oracle_api.have_reference() else oracle_api.Port(quirks)
    ref = orc.matrix(n, A, B, want_grad=True) if orc.kind == "port" else orc.matrix(n, A, B, variant=-1, want_grad=True)
    assert rel_err(f, ref["C"].sum()) <= RTOL
    assert rel_err(g, ref["grad"]) <= RTOL
    # n^2 (leaf + 2n operators); the reference's count has n^2 host-sum entries in place of the leaves
    assert int(st["ops"]) == n * n * (2 * n + 1) == ref["entries"]
    return f


@pytest.mark.gpu
@pytest.mark.parametrize("acc", ["warp", "global"])
@pytest.mark.parametrize("quirks", [True, False], ids=["quirks", "fixed"])
def test_matrix_product_gradient(ctx, oracle_api, quirks, acc):
    """config 2 (tape path): d(sum C)/d(A,B), 2 n^2 independents, 2-D NDRange (n, n) local (16, 16)."""
    n = 16 if acc == "warp" else 48     # per-warp accumulator rows: P = 2 n^2 <= 512
    _matrix_case(ctx, oracle_api, ctx.builtin("matrix_product", quirks=quirks), n, quirks, acc)


@pytest.mark.gpu
def test_reference_matrixmul_cl_runs_on_the_cuda_runtime(ctx, oracle_api):
    from ad4cl_b200 import host
    path = os.path.join(CUBINS, f"ref_matrixmul_cl.q.{host.device_abi()}.cubin")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_matrixmul_cl.q.cubin not built")
    k = ctx.from_cubin(open(path, "rb").read(), "matrixMult", "GSVVOii", quirks=True)
    _matrix_case(ctx, oracle_api, k, 32, True, "global")


@pytest.mark.gpu
def test_reference_tiled_matrixmul_runs_on_the_cuda_runtime(ctx, oracle_api):
    """matrixmul.cl:63-172: __local 16x16 tiles, get_group_id/get_local_id, two work-group barriers
    per tile -- a CTA is one OpenCL work-group, barrier() is __syncthreads()."""
    from ad4cl_b200 import host
    path = os.path.join(CUBINS, f"ref_matrixmul_tiled_cl.q.{host.device_abi()}.cubin")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_matrixmul_tiled_cl.q.*.cubin not built")
    k = ctx.from_cubin(open(path, "rb").read(), "matrixMult", "GSVVOii", quirks=True)
    f = _matrix_case(ctx, oracle_api, k, 48, True, "global")
    # a kernel with barriers cannot be re-run by a single warp: its "uniform" entry records with the lane recorder, same result
    A, B = synth.msvcrt_rand_matrices(48)
    k = ctx.from_cubin(open(path, "rb").read(), "matrixMult", "GSVVOii", quirks=True)
    k.set_args(None, None, host.params(0, 48 * 48), host.params(48 * 48, 48 * 48), None, 48, 48)
    k.configure((48, 48), max_ops=2 * 48 + 1, max_slots=4 * 48 + 1, local_size=(16, 16), acc="global", recorder="uniform")
    f2, _ = k.eval(np.concatenate([A, B]))
    k.free()
    assert abs(f2 - f) <= 1e-12 * abs(f)


@pytest.mark.gpu
def test_matrix_golden_256(ctx):
    """The reference's golden file test/matrix/host.txt: sum of the printed C vs our sum C
    (6 printed digits per element), and the tape length on its last line."""
    from ad4cl_b200 import host
    z = np.load(os.path.join(REPO, "tests", "golden", "matrix256_host_txt.npz"))
    n = 256
    A, B = synth.msvcrt_rand_matrices(n)
    k = ctx.builtin("matrix_product")
    k.set_args(None, None, host.params(0, n * n), host.params(n * n, n * n), None, n, n)
    k.configure((n, n), max_ops=2 * n + 1, max_slots=4 * n + 1, local_size=(16, 16), acc="global")
    f, g = k.eval(np.concatenate([A, B]))
    st = k.stats()
    k.free()
    assert abs(f - z["C"].sum()) / z["C"].sum() < 1e-6
    # matrix_mul.cpp:272 prints stack_current after the host sum: 2n entries per element + n^2 sum
    # entries; ours records the n^2 leaf accumulators of ad_init_var_g instead of the host sum
    assert int(st["ops"]) == int(z["entries"]) == 33619968
    gA = np.repeat(B.reshape(n, n).sum(axis=1)[None, :], n, axis=0).ravel()
    gB = np.repeat(A.reshape(n, n).sum(axis=0)[:, None], n, axis=1).ravel()
    assert rel_err(g, np.concatenate([gA, gB])) <= RTOL


PRIVATE_OVERFLOW_SRC = r"""
__kernel void private_chain(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __constant struct ad_variable* theta,
        __global double* x,
        __global struct ad_variable* out, int size, int links) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_private_gradient_structure p;
        ad_init_p(&p);
        p.current_ad_variable_id = gs->current_ad_variable_id;
        struct ad_variable v = ad_times_vd_p(&p, theta[0], x[id]);
        for (int j = 0; j < links; j++) v = ad_plus_p(&p, ad_times_vd_p(&p, v, 0.5), theta[1]);
        out[id] = ad_commit_p(gs, &p, v);
    }
}
"""


@pytest.mark.gpu
def test_private_stack_overflow_is_reported_not_clipped(ctx):
    """ad_private_gradient_structure (ad.cl:91-97) holds PRIVATE_STACK_SIZE = 100 two-operand entries (200 operand
    slots here) and the reference never checks: a private stack within its limits gives the closed-form gradient,
    one that outgrows them fails the evaluation with AD4CL_ERR_TAPE_OVERFLOW instead of pairing slots with the
    wrong ids or sweeping past the private gradient."""
    from ad4cl_b200 import host
    n = 4099
    x = np.linspace(0.5, 1.5, n)
    xb = ctx.buffer(x)
    theta = np.array([1.25, -0.75])

    def run(links):
        k = ctx.from_source(PRIVATE_OVERFLOW_SRC, "private_chain", "GSVDOii")
        k.set_args(None, None, host.params(0, 2), xb, None, n, links)
        k.configure(n, max_ops=8, max_slots=8)
        try:
            return k.eval(theta)
        finally:
            k.free()

    links = 40                                   # 1 + 2 * 40 = 81 private operators: fits
    f, g = run(links)
    r = 0.5 ** links
    v = theta[0] * x * r + theta[1] * 2.0 * (1.0 - r)
    assert abs(f - v.sum()) <= 1e-12 * abs(v.sum())
    assert abs(g[0] - (x * r).sum()) <= 1e-12 * abs((x * r).sum()) and abs(g[1] - n * 2.0 * (1.0 - r)) <= 1e-12 * n
    with pytest.raises(host.Ad4clError) as e:
        run(80)                                  # 161 operators / 241 operand slots: more than the stack's 100 two-operand entries
    assert e.value.code == host.ERR_TAPE_OVERFLOW


@pytest.mark.gpu
def test_kernels_from_source_get_a_register_tier_instance(ctx, oracle_api):
    """AD4CL_TAPE_AUTO for a kernel created from SOURCE: when it is bound (independents in argument order, integer
    arguments set) the library rebuilds it as a register-tier instance -- local copies of the independents with
    literal ids, small integer arguments as literals so the loops they bound unroll -- and uses it if the compiler
    scalarised the tape (no local memory).  Same bits as the ahead-of-time register-tier build of the same text;
    a tape that cannot be a compile-time shape stays in the arena, and asking for the register tier then fails."""
    from ad4cl_b200 import host
    text = open(os.path.join(REPO, "ad4cl_b200", "kernels", "objectives.cl")).read()
    for name, n in (("regression_logistic", 30011), ("linreg2_simple", 50021)):
        w = W.describe(name, n)
        kb = W.bind(ctx, w)
        fb, gb = kb.eval(w.theta)
        assert kb.stats()["tape_tier"] == "register"
        kb.free()
        d0, d1 = ctx.buffer(w.d0), ctx.buffer(w.d1)
        if w.linreg2:
            kj = ctx.from_source(text, w.kernel, "GSVVDDOi")
            kj.set_args(None, None, host.params(0), host.params(1), d0, d1, None, w.size)
        else:
            kj = ctx.from_source(text, w.kernel, "GSVDDOiii")
            kj.set_args(None, None, host.params(0, w.nparams), d0, d1, None, w.size, w.i0, w.i1)
        kj.configure(w.size, max_ops=w.max_ops, max_slots=w.max_slots, epilogue=w.epilogue, epilogue_n=w.epilogue_n)
        fj, gj = kj.eval(w.theta)
        st = kj.stats()
        fj2, gj2 = kj.eval(w.theta)
        launches = kj.launches_per_eval
        kj.free()
        assert st["tape_tier"] == "register" and launches == 1
        assert fj == fb and np.array_equal(gj, gb), "JIT and ahead-of-time register-tier builds of the same text disagree"
        assert fj2 == fj and np.array_equal(gj2, gj)
        ref = oracle_api.best(False).eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1,
                                                   epilogue=1 if w.epilogue == "half_n_log" else 0, ops_per_item=w.max_ops)
        assert rel_err(fj, ref["f"]) <= RTOL and rel_err(gj, ref["grad"]) <= RTOL
    # 200 steps = 801 slots: not a register-tier tape
    w = W.describe("tape_stress", 5003, steps=200)
    kj = ctx.from_source(text, "tape_stress", "GSVDDOiii")
    d0 = ctx.buffer(w.d0)
    kj.set_args(None, None, host.params(0, w.nparams), d0, None, None, w.size, w.i0, w.i1)
    kj.configure(w.size, max_ops=w.max_ops, max_slots=w.max_slots)
    f, g = kj.eval(w.theta)
    assert kj.stats()["tape_tier"] == "hbm"
    kj.configure(w.size, max_ops=w.max_ops, max_slots=w.max_slots, tape="register")
    with pytest.raises(host.Ad4clError) as e:
        kj.eval(w.theta)
    assert e.value.code == host.ERR_INVALID
    kj.free()
