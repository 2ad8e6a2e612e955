"""The N > 1 path on CPU: world_size 2 (and 3) over gloo.  The per-shard evaluator is
the CPU oracle here (test infrastructure standing in for the device kernel); what is
under test is the host logic of ad4cl_b200/dist.py: contiguous sharding, the single
all-reduce of the raw sums, and the epilogue applied after the reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, n, out_q):
    import sys
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_api as O
    from ad4cl_b200 import dist as D, workloads as W

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = W.describe(name, n, rank=rank, world=world)
    orc = O.Port(False)

    def local_eval(theta, raw):
        # raw sums of this shard: epilogue 0 gives S and dS/dtheta
        r = orc.eval_objective(w.kernel, theta.numpy(), w.d0, w.d1, w.size, w.i0, w.i1, epilogue=0,
                               ops_per_item=w.max_ops)
        raw[: w.nparams] = torch.from_numpy(r["grad"])
        raw[w.nparams] = r["f"]
        raw[w.nparams + 1] = r["entries"] - w.size
        raw[w.nparams + 2] = 0

    obj = D.ShardedObjective(local_eval, w.nparams, w.epilogue, n, torch.device("cpu"))
    f, g = obj.eval(torch.from_numpy(w.theta.copy()))
    out_q.put((rank, f, g.numpy(), float(obj.raw[w.nparams + 1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,n,world", [("regression_linear", 3001, 2), ("catch_at_age", 40000, 2),
                                          ("regression_logistic", 1000, 3)])
def test_sharded_evaluation_matches_single_process(oracle_api, name, n, world):
    from ad4cl_b200 import workloads as W
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = W.describe(name, n)
    ref = oracle_api.Port(False).eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1,
                                                epilogue=1 if w.epilogue == "half_n_log" else 0,
                                                ops_per_item=w.max_ops)
    for rank, f, g, ops in results:
        assert abs(f - ref["f"]) <= 1e-12 * abs(ref["f"])
        assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.max(np.abs(ref["grad"]))
        assert ops == w.ops_total
    # every rank holds the same bits after the all-reduce
    assert all(r[1] == results[0][1] and np.array_equal(r[2], results[0][2]) for r in results)


def test_epilogue_math():
    from ad4cl_b200 import dist as D
    raw = torch.tensor([2.0, -4.0, 8.0, 11.0, 12.0], dtype=torch.float64)   # P = 2, S = 8
    out = D.apply_epilogue(raw, 2, "half_n_log", 10.0)
    assert out[2].item() == 5.0 * np.log(8.0)
    assert out[0].item() == 2.0 * (5.0 * (1.0 / 8.0)) and out[1].item() == -4.0 * (5.0 * (1.0 / 8.0))
    assert out[3].item() == 11.0 and out[4].item() == 12.0
    assert D.apply_epilogue(raw, 2, "sum", 10.0) is raw


@pytest.mark.parametrize("name,n", [("catch_at_age", 8000), ("regression_linear", 1001), ("tape_stress", 333)])
def test_shards_partition_the_problem_for_every_world_size(oracle_api, name, n):
    """What `bench.py --gpus N` relies on at N = 1, 2, 4, 8 (and any other N): the shards workloads.describe hands
    to the ranks partition the observations -- the raw sums S and dS/dtheta of the shards add up to the
    single-shard values, and so do the operator / slot / input-byte totals the roofline is computed from.
    catch-at-age switches layout with N: whole (year, age) cells dealt round-robin when N divides 400 (2, 4, 5, 8),
    a slice of the replicates of every cell otherwise (3, 6, 7); both are the same kernel text, told apart by i1
    (kernels/objectives.cl).  One process, the CPU oracle standing in for the device kernel."""
    from ad4cl_b200 import workloads as W
    orc = oracle_api.Port(False)

    def raw(w):
        r = orc.eval_objective(w.kernel, w.theta, w.d0, w.d1, w.size, w.i0, w.i1, epilogue=0, ops_per_item=w.max_ops)
        return r["f"], np.asarray(r["grad"]), r["entries"] - w.size

    whole = W.describe(name, n)
    S, dS, ops = raw(whole)
    assert ops == whole.ops_total
    for world in range(2, 9):
        shards = [W.describe(name, n, rank=r, world=world) for r in range(world)]
        assert sum(w.size for w in shards) == whole.size
        assert sum(w.ops_total for w in shards) == whole.ops_total
        assert sum(w.slots_total for w in shards) == whole.slots_total
        assert sum(w.input_bytes for w in shards) == whole.input_bytes
        if name == "catch_at_age":   # the layout switch, as documented
            assert all((w.i1 < 0) == (400 % world == 0) for w in shards)
        parts = [raw(w) for w in shards]
        assert [p[2] for p in parts] == [w.ops_total for w in shards]
        Ssum = sum(p[0] for p in parts)
        dSsum = np.sum([p[1] for p in parts], axis=0)
        assert abs(Ssum - S) <= 1e-12 * abs(S), (world, Ssum, S)
        assert np.max(np.abs(dSsum - dS)) <= 1e-12 * np.max(np.abs(dS)), world
