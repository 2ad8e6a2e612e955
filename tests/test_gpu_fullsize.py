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
