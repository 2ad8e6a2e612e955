"""Evaluate rank 0's shard of an 8-way sharded catch-at-age 10^7 on ONE GPU (no communicator) a few times:
the per-launch cost of a small shard, for ncu / event timing."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from ad4cl_b200 import host, workloads as W

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = {}
if len(sys.argv) > 2 and sys.argv[2] != "auto":
    rec, ls, block = sys.argv[2].split(",")
    cfg = dict(recorder=rec, lockstep=ls, local_size=int(block), prefetch="off")
ctx = host.Context(0)
w = W.describe("catch_at_age", 10_000_000, rank=0, world=world)
k = W.bind(ctx, w, **cfg)
k.profile(True)
for _ in range(5):
    k.eval(w.theta)
k.profile_read(reset=True)
for _ in range(int(os.environ.get("PROBE_EVALS", "20"))):
    k.eval(w.theta)
ms, n = k.profile_read()
print("world", world, "items", w.size, "kernel_ms", round(ms, 4), "lib", os.path.basename(host.LIB_PATH), k.stats())
k.free()
