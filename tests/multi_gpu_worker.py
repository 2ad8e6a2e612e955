"""torchrun worker for tests/test_multi_gpu.py: every rank evaluates its shard on its own GPU,
the raw sums are combined (a) by NCCL through torch.distributed and (b) by the library's one-shot
NVLink all-reduce, and rank 0 additionally evaluates the whole problem on one GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
from ad4cl_b200 import dist as D, host, workloads as W  # noqa: E402


def main():
    name, n = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = host.Context(local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.use_stream(stream.cuda_stream)
    w = W.describe(name, n, rank=rank, world=world)
    k = W.bind(ctx, w)
    theta = torch.from_numpy(w.theta.copy()).to(dev)
    out = {}
    obj = D.from_kernel(k, w, dev)
    f, g = obj.eval(theta)
    out["nccl"] = (f, g.numpy().tolist())
    comm = D.connect_oneshot(ctx, w.nparams)
    obj2 = D.from_kernel(k, w, dev, comm=comm)
    res = []
    for _ in range(5):                      # several epochs: both parities of the inbox
        f2, g2 = obj2.eval(theta)
        res.append((f2, g2.numpy().tolist()))
    comm.check()
    out["oneshot"] = res
    k.check()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        full = W.describe(name, n)
        kf = W.bind(ctx, full)
        ff, gf = kf.eval(full.theta)
        kf.free()
        print("RESULT " + json.dumps({"single": (ff, gf.tolist()), "ranks": gathered}))
    k.free()
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
