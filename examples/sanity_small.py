#!/usr/bin/env python
"""A handful of small evaluations touching every device code path (three accumulator modes, the tape
tiers, both recorders with and without prefetch / lock step, far plane, private stack + pre-accumulation,
2-D NDRange, dense GEMMs, JIT incl. its register-tier instance) -- the program run
under compute-sanitizer (`compute-sanitizer --tool memcheck python examples/sanity_small.py`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host, synth, workloads as W  # noqa: E402


def main():
    ctx = host.Context(0)
    for name, n, cfg in (("catch_at_age", 4000, {}), ("catch_at_age", 4000, {"window": 4}),
                         ("regression_logistic", 777, {"acc": "thread"}), ("regression_logistic", 777, {"acc": "warp"}),
                         ("regression_logistic", 777, {"acc": "global"}),
                         ("regression_logistic", 333, {"tape": "shared", "local_size": 64}),
                         ("tape_stress", 97, {}), ("matrix_product", 32, {}), ("linreg2_simple", 1003, {}),
                         # both recorders, each with and without the sweep prefetch / the lock-step barrier, in the arena tier
                         ("catch_at_age", 4000, {"recorder": "uniform", "prefetch": "on", "lockstep": "on", "tape": "hbm"}),
                         ("catch_at_age", 4000, {"recorder": "uniform", "prefetch": "off", "lockstep": "off", "local_size": 768}),
                         ("catch_at_age", 4000, {"recorder": "lane", "prefetch": "on", "lockstep": "on", "local_size": 96}),
                         ("tape_stress", 3001, {"recorder": "uniform", "prefetch": "on", "lockstep": "off"}),
                         ("tape_stress", 3001, {"recorder": "lane", "prefetch": "off", "lockstep": "off"}),
                         ("regression_linear", 2003, {"tape": "hbm", "recorder": "uniform"}),
                         ("linreg2_simple", 1003, {"tape": "hbm", "recorder": "uniform", "acc": "warp"})):
        w = W.describe(name, n)
        k = W.bind(ctx, w, **cfg)
        f, g = k.eval(w.theta)
        f0, _ = k.eval(w.theta, want_grad=False)
        k.free()
        print(name, n, cfg, f, float(np.abs(g).max()))
    for kern in ("op_zoo_pad", "op_zoo_private"):
        w = W.describe("op_zoo", 100)
        k = ctx.builtin(kern)
        k.set_args(None, None, host.params(0, 3), ctx.buffer(w.d0), None, None, w.size, 0, 0)
        k.configure(w.size, max_ops=48, max_slots=96)
        print(kern, k.eval(w.theta)[0])
        k.free()
    n = 64
    A, B = synth.msvcrt_rand_matrices(n)
    bA, bB, bC, bdA, bdB = ctx.buffer(A), ctx.buffer(B), ctx.buffer(nbytes=8 * n * n), ctx.buffer(nbytes=8 * n * n), ctx.buffer(nbytes=8 * n * n)
    tot, ws = ctx.buffer(nbytes=8), ctx.buffer(nbytes=8 * (n * n + 1))
    for variant in (0, 1, 2):
        ctx.matrix_product_dense(n, bA, bB, C=bC, dA=bdA, dB=bdB, total=tot, workspace=ws, variant=variant)
        print("dense", variant, tot.download()[0])
    text = open(os.path.join(os.path.dirname(__file__), "..", "ad4cl_b200", "kernels", "objectives.cl")).read()
    kj = ctx.from_source(text, "linreg2_g", "GSVVDDOi")
    x, y, (a, b) = synth.main_cpp_recipe(500)
    kj.set_args(None, None, host.params(0), host.params(1), ctx.buffer(x), ctx.buffer(y), None, 500)
    kj.configure(500, max_ops=4, epilogue="half_n_log")
    print("jit", kj.eval([a, b])[0])
    kj.free()
    ctx.close()


if __name__ == "__main__":
    main()
