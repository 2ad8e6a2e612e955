#!/usr/bin/env python
"""A caller of the hot path: fit the 100 catch-at-age parameters by L-BFGS, every
objective+gradient evaluation on the GPU through the C ABI -- the role ADMB's
function_minimizer plays around userfunction() in the reference
(test/admb/simple.cpp:481-492).  The optimiser itself is scipy's; nothing of the
evaluation happens on the host.

    python examples/fit_catch_at_age.py [n_observations]
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host, synth, workloads as W  # noqa: E402


def fit(n: int = 400_000, maxiter: int = 200, verbose: bool = True):
    ctx = host.Context(0)
    w = W.describe("catch_at_age", n)
    k = W.bind(ctx, w)
    evals = 0

    def fun(theta):
        nonlocal evals
        evals += 1
        f, g = k.eval(theta)
        return f, g

    t0 = time.perf_counter()
    res = minimize(fun, w.theta, jac=True, method="L-BFGS-B",
                   options={"maxiter": maxiter, "maxfun": 4 * maxiter, "ftol": 1e-15, "gtol": 1e-10 * n})
    dt = time.perf_counter() - t0
    k.free()
    ctx.close()
    truth = synth.caa_theta_true()
    if verbose:
        print(f"n={n} evaluations={evals} iterations={res.nit} wall={dt:.2f}s ({evals / dt:.0f} evals/s) "
              f"f={res.fun:.6f} |grad|inf={np.abs(res.jac).max():.3e}")
        print(f"max |theta_hat - theta_true| = {np.abs(res.x - truth).max():.4f} "
              f"(start: {np.abs(w.theta - truth).max():.4f})")
    return res, truth, evals


if __name__ == "__main__":
    fit(int(sys.argv[1]) if len(sys.argv) > 1 else 400_000)
