"""The sharded evaluation in ONE call (ad4cl_cuda_eval_sharded): every rank's kernel reduces its shard and
its LAST CTA exchanges the raw sums with the peers (one-shot all-reduce over peer memory), applies the
epilogue and writes the result -- one launch per evaluation, replayed as one CUDA graph.

Runs on a 1-GPU box: the ranks are contexts of the same process (ad4cl_cuda_comm_create_local), all on
device 0 when there is only one (the protocol -- slots, epochs, parities, rank-ordered sum -- is the same;
with several GPUs the ranks are spread over them and the stores cross NVLink)."""
import threading

import numpy as np
import pytest

from ad4cl_b200 import workloads as W

pytestmark = pytest.mark.gpu


def run_ranks(world, name, n, epochs=5, **kw):
    import torch
    from ad4cl_b200 import host
    ndev = torch.cuda.device_count()
    ctxs = [host.Context(r % ndev) for r in range(world)]
    ws = [W.describe(name, n, rank=r, world=world, **kw) for r in range(world)]
    P = ws[0].nparams
    comms = host.Comm.create_local(ctxs, P + 3)
    for c in comms:
        c.set_timeout(60.0)
    ks = [W.bind(ctxs[r], ws[r]) for r in range(world)]
    out = [[] for _ in range(world)]
    err = []

    def work(r):
        try:
            for _ in range(epochs):                 # both parities of the inbox, graph capture + replays
                out[r].append(ks[r].eval_sharded(comms[r], ws[r].theta))
            out[r].append(ks[r].eval_sharded(comms[r], ws[r].theta, want_grad=False))
        except Exception as e:  # noqa: BLE001
            err.append((r, e))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=180)
    assert not err, err
    launches = [k.launches_per_eval for k in ks]
    for k in ks:
        k.free()
    for c in comms:
        c.destroy()
    for c in ctxs:
        c.close()
    return out, launches


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name,n,kw", [("catch_at_age", 240000, {}), ("regression_linear", 100003, {}),
                                        ("tape_stress", 30001, {"steps": 100})])
def test_sharded_evaluation_in_one_call(ctx, world, name, n, kw):
    out, launches = run_ranks(world, name, n, **kw)
    full = W.describe(name, n, **kw)
    k = W.bind(ctx, full)
    f1, g1 = k.eval(full.theta)
    f1_only, _ = k.eval(full.theta, want_grad=False)
    k.free()
    assert launches == [1] * world                              # all-reduce and epilogue inside the evaluation kernel
    scale = np.abs(g1).max()
    first = out[0][0]
    for r in range(world):
        for f, g in out[r][:-1]:
            assert abs(f - f1) <= 1e-12 * abs(f1) and np.abs(g - g1).max() <= 1e-12 * scale
            # rank-ordered sum: every rank and every epoch holds the same bits
            assert f == first[0] and np.array_equal(g, first[1])
        f_only, g_only = out[r][-1]
        assert g_only is None and abs(f_only - f1_only) <= 1e-12 * abs(f1_only)


def test_a_missing_peer_is_an_error_not_a_wrong_number(ctx):
    """One rank evaluates, its peer never does: after the time-out the call fails (AD4CL_ERR_CUDA) instead of
    returning the un-reduced local sums (round-1 advisor finding)."""
    from ad4cl_b200 import host
    ctxs = [host.Context(0), host.Context(0)]
    w = W.describe("regression_linear", 20000, rank=0, world=2)
    comms = host.Comm.create_local(ctxs, w.nparams + 3)
    comms[0].set_timeout(1.0)
    k = W.bind(ctxs[0], w)
    with pytest.raises(host.Ad4clError) as e:
        k.eval_sharded(comms[0], w.theta)
    assert e.value.code == host.ERR_CUDA and "timed out" in str(e.value)
    k.free()
    for c in comms:
        c.destroy()
    for c in ctxs:
        c.close()
