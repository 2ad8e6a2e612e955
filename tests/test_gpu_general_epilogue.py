"""Non-uniform seeds: an objective that is NOT of the form h(sum_i out_i).
f = sum_i w_i log(out_i + 3): pass 1 stores out_i, the seeds s_i = w_i / (out_i + 3) are formed, pass 2 sweeps every
work-item from its own seed (ad4cl_cuda_eval_general) -- against the reference's own runtime, which records
the same epilogue with ad_plus_vd / ad_log / ad_times_dv / ad_plus_eq_v on the host tape and sweeps the whole stack with
compute_gradient (ad4cl.h:775-802; oracle/ref_host.cpp ref_eval_weighted_log), at 1e-12.  No tape export."""
import numpy as np
import pytest

from ad4cl_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n,cfg", [("regression_linear", 5003, {}), ("regression_linear", 5003, {"tape": "hbm"}),
                                          ("linreg2_main", 3000, {}), ("catch_at_age", 8000, {})],
                         ids=["regression-register-tier", "regression-arena", "linreg2", "catch-at-age"])
def test_weighted_log_objective_matches_the_reference_runtime(ctx, oracle_api, name, n, cfg):
    w = W.describe(name, n)
    rng = np.random.default_rng(5)
    wts = rng.uniform(0.5, 1.5, w.size)
    k = W.bind(ctx, w, **cfg)
    s, out_buf = k.eval_outputs(w.theta, w.size)
    out = out_buf.download()
    SHIFT = 3.0
    assert np.all(out + SHIFT > 0) and abs(out.sum() - s) <= 1e-12 * abs(s)
    f = float(np.sum(wts * np.log(out + SHIFT)))
    seeds = ctx.buffer(wts / (out + SHIFT))
    g = k.eval_seeded(w.theta, seeds)
    g2 = k.eval_seeded(w.theta, seeds)
    k.free()
    assert np.array_equal(g, g2)
    if oracle_api.have_reference():
        ref = oracle_api.Reference(False).eval_weighted_log(w.kernel, w.theta, w.d0, w.d1, w.size, wts, w.i0, w.i1,
                                                            ops_per_item=w.max_ops, shift=SHIFT)
        assert np.max(np.abs(out - ref["out"])) <= 1e-12 * np.abs(ref["out"]).max()
        assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
        assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.abs(ref["grad"]).max()
    if name == "regression_linear":     # closed form: out_i = r_i^2, df/dtheta_j = sum_i w_i 2 r_i X_ji / (r_i^2 + 3)
        X = w.d0.reshape(w.nparams, -1)
        r = X.T @ w.theta - w.d1
        g_cf = (X * (2.0 * wts * r / (r * r + SHIFT))).sum(axis=1)
        assert np.max(np.abs(g - g_cf)) <= 1e-12 * np.abs(g_cf).max()
    out_buf.free()
    seeds.free()
