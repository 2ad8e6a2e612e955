"""BASELINE.json's full sizes, where the CPU oracle would take minutes: parity through properties
that do not need it --
  * idempotence: the same evaluation twice gives the same bits (fixed-order reduction);
  * additivity over shards: f and grad of the whole = sum over replicate / observation shards
    evaluated separately (the objective is a sum over data; this is also what the multi-GPU path
    relies on);
  * the gradient is the derivative of the objective the same kernel computes: central
    differences of f along random directions against grad . d;
  * the device's own operator / slot counters equal the analytic totals (every work-item recorded
    its whole tape; nothing overflowed silently).
plus the oracle itself on a strided subsample where the workload allows it."""
import numpy as np
import pytest

from ad4cl_b200 import synth, workloads as W

pytestmark = pytest.mark.gpu


def evaluate(ctx, w, **cfg):
    k = W.bind(ctx, w, **cfg)
    f, g = k.eval(w.theta)
    f2, g2 = k.eval(w.theta)
    st = k.stats()
    assert f == f2 and np.array_equal(g, g2), "not idempotent"
    return k, f, g, st


@pytest.mark.parametrize("name,n,kw", [("catch_at_age", 10_000_000, {}), ("regression_linear", 1_000_000, {}),
                                       ("regression_logistic", 1_000_000, {}),
                                       ("tape_stress", 10_000_000, {"steps": 33})],
                         ids=["config4-1e7x100", "config3-linear-1e6x10", "config3-logistic-1e6x10", "config5-1e7xL100"])
def test_full_size_properties(ctx, name, n, kw):
    w = W.describe(name, n, **kw)
    k, f, g, st = evaluate(ctx, w)
    assert int(st["ops"]) == w.ops_total and int(st["slots"]) == w.slots_total
    # directional derivative of the SAME device objective
    rng = np.random.default_rng(7)
    for _ in range(2):
        d = rng.standard_normal(w.nparams)
        d /= np.linalg.norm(d)
        eps = 1e-6
        fp, _ = k.eval(w.theta + eps * d, want_grad=False)
        fm, _ = k.eval(w.theta - eps * d, want_grad=False)
        fd = (fp - fm) / (2 * eps)
        assert abs(fd - g @ d) <= 2e-7 * max(abs(g @ d), np.linalg.norm(g) * 1e-3), (fd, g @ d)
    k.free()
    # additivity over 3 shards (raw sums; the (n/2) log epilogue is applied after the sum)
    if w.epilogue == "sum":
        parts = [W.describe(name, n, rank=r, world=3, **kw) for r in range(3)]
        fs, gs = 0.0, np.zeros(w.nparams)
        for p in parts:
            kp = W.bind(ctx, p)
            fp_, gp_ = kp.eval(p.theta)
            kp.free()
            fs += fp_
            gs += gp_
        assert abs(fs - f) <= 1e-12 * abs(f)
        assert np.max(np.abs(gs - g)) <= 1e-12 * np.max(np.abs(g))


def test_catch_at_age_full_size_against_oracle_on_a_replicate_slice(ctx, oracle_api):
    """The oracle cannot run 10^7 observations in seconds, but it can run one replicate slice of the
    10^7-observation problem (every (year, age) cell, 50 replicates): same inputs, same kernel."""
    n = 10_000_000
    obs, theta0, R = synth.catch_at_age(n, 12345, 50)
    ref = oracle_api.best(False).eval_objective("catch_at_age", theta0, obs, np.array([float(R)]), len(obs), 50, 12345, ops_per_item=130)
    w = W.describe("catch_at_age", n)
    w.d0, w.size, w.i0, w.i1 = obs, len(obs), 50, 12345
    k = W.bind(ctx, w)
    f, g = k.eval(theta0)
    k.free()
    assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))


def test_catch_at_age_full_size_against_oracle_summed_over_shards(ctx, oracle_api):
    """Config 4 at its full size, 10^7 observations x 100 parameters, against the CPU oracle itself: the
    objective is a plain sum over observations, so the oracle evaluates 10 replicate shards of 10^6
    observations (the reference's 40-byte-entry stack of one shard is ~5 GB) and the shard results are added
    in fp64; objective and all 100 gradient components at 1e-12."""
    n, shards = 10_000_000, 10
    orc = oracle_api.best(False)
    f_ref, g_ref = 0.0, None
    for r in range(shards):
        p = W.describe("catch_at_age", n, rank=r, world=shards)
        ref = orc.eval_objective("catch_at_age", p.theta, p.d0, p.d1, p.size, p.i0, p.i1, ops_per_item=130)
        f_ref += ref["f"]
        g_ref = np.array(ref["grad"]) if g_ref is None else g_ref + np.array(ref["grad"])
    w = W.describe("catch_at_age", n)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    k.free()
    assert abs(f - f_ref) <= 1e-12 * abs(f_ref)
    assert np.max(np.abs(g - g_ref)) <= 1e-12 * np.max(np.abs(g_ref))


def tape_stress_closed_form(sum_x, steps, P):
    """v_0 = theta_0 x, v_{j+1} = v_j / 2 + theta_{j mod P} x at theta = 1 (objectives.cl, tape_stress)."""
    g = np.zeros(P)
    for j in range(max(0, steps - 1100), steps):
        g[j % P] += 2.0 ** -(steps - 1 - j)
    g[0] += 2.0 ** -steps
    return sum_x * (2.0 - 2.0 ** -steps), sum_x * g


@pytest.mark.parametrize("n,steps", [(100_003, 3333), (1_000_000, 333)], ids=["1e5x1e4ops", "1e6x1e3ops"])
def test_long_tapes_on_many_work_items(ctx, n, steps):
    """Config 5's upper end: 10^4 AD operators per work-item on 10^5 work-items (13 333 slots per tape, the
    arena streams through HBM), against the closed form of the recurrence; counters = analytic totals."""
    w = W.describe("tape_stress", n, steps=steps)
    k, f, g, st = evaluate(ctx, w, blocks_per_sm=1 if steps > 1000 else 0)
    k.free()
    assert int(st["ops"]) == w.ops_total and int(st["slots"]) == w.slots_total
    f_ref, g_ref = tape_stress_closed_form(float(np.sum(w.d0)), steps, w.nparams)
    assert abs(f - f_ref) <= 1e-12 * abs(f_ref)
    assert np.max(np.abs(g - g_ref)) <= 1e-12 * np.max(np.abs(g_ref))


def test_matrix_1024_tape_and_dense_paths(ctx):
    """Config 2 at the BASELINE size n = 1024 (2.1e9 AD operators through the tape; the reference's own
    stack for it would be 86 GB, so the check is the closed form of d(sum A B): dL/dA[j,k] = sum_i B[k,i],
    dL/dB[k,i] = sum_j A[j,k]) -- for the tape path and for the dense-contraction path."""
    n = 1024
    A, B = synth.msvcrt_rand_matrices(n)
    Am, Bm = A.reshape(n, n), B.reshape(n, n)
    f_ref = float((Am.sum(axis=0) * Bm.sum(axis=1)).sum())
    gA = np.repeat(Bm.sum(axis=1)[None, :], n, axis=0).ravel()
    gB = np.repeat(Am.sum(axis=0)[:, None], n, axis=1).ravel()
    g_ref = np.concatenate([gA, gB])
    w = W.describe("matrix_product", n)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    st = k.stats()
    k.free()
    assert int(st["ops"]) == n * n * (2 * n + 1)
    assert abs(f - f_ref) <= 1e-12 * abs(f_ref) and np.max(np.abs(g - g_ref)) <= 1e-12 * np.abs(g_ref).max()
    bA, bB = ctx.buffer(A), ctx.buffer(B)
    bC, bdA, bdB = (ctx.buffer(nbytes=8 * n * n) for _ in range(3))
    tot, ws = ctx.buffer(nbytes=8), ctx.buffer(nbytes=8 * (2 * n * n + 4096))
    for variant in (0, 1, 2):
        ctx.matrix_product_dense(n, bA, bB, C=bC, dA=bdA, dB=bdB, total=tot, workspace=ws, variant=variant)
        C = bC.download().reshape(n, n)
        assert np.max(np.abs(C - Am @ Bm)) <= 1e-12 * np.abs(Am @ Bm).max()
        assert abs(tot.download()[0] - f_ref) <= 1e-12 * abs(f_ref)
        assert np.max(np.abs(bdA.download() - gA)) <= 1e-12 * np.abs(gA).max()
        assert np.max(np.abs(bdB.download() - gB)) <= 1e-12 * np.abs(gB).max()
    for b in (bA, bB, bC, bdA, bdB, tot, ws):
        b.free()
