"""Kernels that do NOT record a warp-uniform tape: lane-dependent parameter ids, data-dependent control
flow, in-place accumulation into the virtual `out`, ragged 2-D NDRanges.  The uniform fast path of
ad4cl_tape.cuh must hand such warps to the per-lane format / general sweep without changing a result.
Checked against closed forms computed with numpy (the objectives are simple on purpose)."""
import numpy as np
import pytest

from ad4cl_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12

SRC = r'''
/* out = theta[g(id)] * x + theta[P-1]^2 : the parameter id depends on the lane (g = id % (P-1)) */
__kernel void lane_params(__global struct ad_gradient_structure* gs, __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta, __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_variable t = ad_times_vd(gs, theta[id % (i0 - 1)], d0[id]);
        out[id] = ad_plus(gs, t, ad_times(gs, theta[i0 - 1], theta[i0 - 1]));
    }
}

/* data-dependent control flow: work-items with x > 0.5 record 3 extra operators, every 7th none at all */
__kernel void branchy(__global struct ad_gradient_structure* gs, __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta, __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size && id % 7 != 3) {
        const double x = d0[id];
        struct ad_variable v = ad_times_vd(gs, theta[0], x);
        if (x > 0.5) {
            v = ad_exp(gs, ad_times_vd(gs, v, 0.25));
            v = ad_plus(gs, v, theta[1]);
        }
        v = ad_times(gs, v, v);
        out[id] = ad_plus_vd(gs, v, 1.0);
    }
}

/* accumulates into out[id] IN PLACE (ad.cl:263-283, the __global-lhs += of the reference) */
__kernel void acc_out(__global struct ad_gradient_structure* gs, __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta, __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        ad_plus_eq_g(gs, &out[id], ad_times_vd(gs, theta[0], d0[id]));
        ad_plus_eq_g(gs, &out[id], ad_times(gs, theta[1], theta[1]));
        ad_plus_eq_d(gs, &out[id], 3.0);
    }
}
'''


def rel(x, ref):
    x, ref = np.atleast_1d(x), np.atleast_1d(ref)
    return float(np.max(np.abs(x - ref)) / max(np.abs(ref).max(), 1e-300))


def make(ctx, entry, theta, x, n, i0=0, **cfg):
    from ad4cl_b200 import host
    k = ctx.from_source(SRC, entry, "GSVDDOiii")
    k.set_args(None, None, host.params(0, len(theta)), ctx.buffer(x), None, None, n, i0, 0)
    k.configure(n, max_ops=16, **cfg)
    return k


@pytest.mark.parametrize("n", [5, 1000, 4099])
@pytest.mark.parametrize("acc", ["thread", "warp", "global"])
def test_lane_dependent_parameter_ids(ctx, n, acc):
    P = 6
    rng = np.random.default_rng(7)
    theta, x = rng.uniform(0.5, 1.5, P), rng.uniform(0.1, 0.9, n)
    k = make(ctx, "lane_params", theta, x, n, i0=P, acc=acc)
    f, g = k.eval(theta)
    k.free()
    idx = np.arange(n) % (P - 1)
    f_ref = float(np.sum(theta[idx] * x) + n * theta[-1] ** 2)
    g_ref = np.zeros(P)
    np.add.at(g_ref, idx, x)
    g_ref[-1] = 2 * n * theta[-1]
    assert rel(f, f_ref) <= RTOL and rel(g, g_ref) <= RTOL


@pytest.mark.parametrize("n", [31, 1000, 70001])
def test_data_dependent_control_flow(ctx, n):
    rng = np.random.default_rng(11)
    theta, x = np.array([1.3, 0.7]), rng.uniform(0.0, 1.0, n)
    k = make(ctx, "branchy", theta, x, n)
    f, g = k.eval(theta)
    f2, g2 = k.eval(theta)
    k.free()
    assert f == f2 and np.array_equal(g, g2)           # bit-reproducible
    live = (np.arange(n) % 7) != 3
    hi = live & (x > 0.5)
    v = theta[0] * x
    e = np.exp(0.25 * v)
    u = np.where(hi, e + theta[1], v)
    out = np.where(live, u * u + 1.0, 0.0)
    du0 = np.where(hi, e * 0.25 * x, x)
    du1 = np.where(hi, 1.0, 0.0)
    g_ref = np.array([np.sum(np.where(live, 2 * u * du0, 0.0)), np.sum(np.where(live, 2 * u * du1, 0.0))])
    assert rel(f, out.sum()) <= RTOL and rel(g, g_ref) <= RTOL


def test_in_place_accumulation_into_out(ctx):
    """The virtual out starts as 'not a variable': using it as the left-hand side of += must record a
    constant operand, not an id the sweep follows out of the arena (round-1 advisor finding)."""
    n = 1003
    rng = np.random.default_rng(3)
    theta, x = np.array([0.9, 1.1]), rng.uniform(0.1, 0.9, n)
    k = make(ctx, "acc_out", theta, x, n)
    f, g = k.eval(theta)
    k.free()
    assert rel(f, float(np.sum(theta[0] * x) + n * (theta[1] ** 2 + 3.0))) <= RTOL
    assert rel(g, np.array([x.sum(), 2 * n * theta[1]])) <= RTOL


@pytest.mark.parametrize("n", [20, 33])
def test_ragged_2d_ndrange(ctx, n):
    """n is not a multiple of the 16 x 16 work-group: the NDRange is padded, the kernel indexes
    C[i + widthA * j] with the CONFIGURED width (matrixmul.cl:27), and the gradient is that of sum(A B)."""
    from ad4cl_b200 import host
    A, B = synth.msvcrt_rand_matrices(n)
    k = ctx.builtin("matrix_product")
    k.set_args(None, None, host.params(0, n * n), host.params(n * n, n * n), None, n, n)
    k.configure((n, n), max_ops=2 * n + 1, max_slots=4 * n + 1, local_size=(16, 16), acc="global")
    f, g = k.eval(np.concatenate([A, B]))
    k.free()
    Am, Bm = A.reshape(n, n), B.reshape(n, n)
    gA = np.repeat(Bm.sum(axis=1)[None, :], n, axis=0)       # dL/dA[j,k] = sum_i B[k,i]
    gB = np.repeat(Am.sum(axis=0)[:, None], n, axis=1)       # dL/dB[k,i] = sum_j A[j,k]
    assert rel(f, (Am @ Bm).sum()) <= RTOL
    assert rel(g, np.concatenate([gA.ravel(), gB.ravel()])) <= RTOL
