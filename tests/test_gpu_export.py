"""Tape export (ad4cl_cuda_export_tape): the device tape rewritten as the reference's 40-byte
entries, consumed by the reference's own host-side procedure -- host sum of out[] with in-place
entries, scalar epilogue, the serial sweep of ad4cl.h:775-802 (restated below in numpy) -- must give the
objective and gradient of the fused device path and of the oracle, and as many entries as the
reference records."""
import numpy as np
import pytest

from ad4cl_b200 import workloads as W

pytestmark = pytest.mark.gpu


def reference_host_finish(entries, out, nparams, epilogue_n=None):
    """simple.cpp:357-369 on an exported stack: sum (fresh id) += out[i]; optional (n/2) log; sweep."""
    stack = [(int(e["id"]), int(e["size"]), float(e["dx0"]), int(e["id0"]), float(e["dx1"]), int(e["id1"])) for e in entries]
    next_id = max([nparams] + [s[0] + 1 for s in stack])
    sid, S = next_id, 0.0
    next_id += 1
    for o in out:
        S += float(o["value"])
        stack.append((sid, 2, 1.0, sid, 1.0, int(o["id"])))            # ad_plus_eq_v, ad4cl.h:186-199
    f, last = S, sid
    if epilogue_n is not None:
        lid, fid = next_id, next_id + 1
        next_id += 2
        stack.append((lid, 1, 1.0 / S, sid, 0.0, 0))                     # ad_log
        stack.append((fid, 1, epilogue_n / 2.0, lid, 0.0, 0))            # ad_times_dv
        f, last = epilogue_n / 2.0 * np.log(S), fid
    g = np.zeros(next_id + 1)
    g[last] = 1.0
    for (rid, size, dx0, id0, dx1, id1) in reversed(stack):              # ad4cl.h:786-797
        w = g[rid]
        g[rid] = 0.0
        if size >= 1:
            g[id0] += w * dx0
        if size == 2:
            g[id1] += w * dx1
    return f, g[:nparams]


@pytest.mark.parametrize("name,n,kw", [("linreg2_main", 1000, {}), ("regression_logistic", 517, {}),
                                       ("catch_at_age", 2000, {}), ("op_zoo", 257, {}), ("matrix_product", 16, {})])
def test_exported_tape_reproduces_the_gradient(ctx, oracle_api, name, n, kw):
    w = W.describe(name, n, **kw)
    k = W.bind(ctx, w)
    f_dev, g_dev = k.eval(w.theta)
    ops = int(k.stats()["ops"])
    items = n * n if w.matrix else w.size
    entries, out = k.export_tape(w.theta, items, capacity=ops + 16)
    k.free()
    assert len(entries) == ops                                 # one reference entry per recorded operator
    f, g = reference_host_finish(entries, out, w.nparams, float(w.epilogue_n) if w.epilogue == "half_n_log" else None)
    assert abs(f - f_dev) <= 1e-12 * abs(f_dev)
    assert np.max(np.abs(g - g_dev)) <= 1e-12 * np.max(np.abs(g_dev))
    if not w.matrix:
        ref = oracle_api.best(False).eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1,
                                                    epilogue=1 if w.epilogue == "half_n_log" else 0, ops_per_item=w.max_ops)
        extra = w.size + (2 if w.epilogue == "half_n_log" else 0)
        assert len(entries) == ref["entries"] - extra          # the reference's own tape length
        assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
        assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))


def test_export_capacity_is_checked(ctx):
    from ad4cl_b200 import host
    w = W.describe("linreg2_main", 1000)
    k = W.bind(ctx, w)
    with pytest.raises(host.Ad4clError) as e:
        k.export_tape(w.theta, 1000, capacity=100)
    assert e.value.code == host.ERR_TAPE_OVERFLOW
    k.free()
