#!/usr/bin/env python
"""The reference's test/admb example end to end on the CUDA path, with its own file formats:
reads a `simple.dat`-format configuration (test/admb/simple.dat:1-12: number of observations,
gradient method, stack size, path to ad.cl, path to the kernel file, gpu index), generates the
simple.cpp:58-69 data (x ~ 150 U, y = 2 x + 4 + 7 N(0,1)), JIT-builds the kernel file it names
(exactly what simple.cpp:138-175 does with OpenCL), fits a and b with the library's L-BFGS (the role
of ADMB's function_minimizer, simple.cpp:481-492) and writes `simple.par` in ADMB's format
(test/admb/bin/simple.par:1-5).

    python examples/simple_admb.py [simple.dat] [out_dir]

With no arguments: 1 000 003 observations, the built-in linreg2_p kernel (the text of simple.cl's
AD), device 0.  Only "gradient method 1 = ad4cl_device" exists here: there is no host method.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host, synth  # noqa: E402


def read_dat(path):
    """ADMB .dat: '#' comment lines, whitespace-separated values in declaration order."""
    vals = [l.strip() for l in open(path) if l.strip() and not l.lstrip().startswith("#")]
    keys = ["nobs", "method", "ad4cl_stack_size", "ad4cl_api", "kernel_code", "gpu_index"]
    cfg = dict(zip(keys, vals))
    for k in ("nobs", "method", "ad4cl_stack_size", "gpu_index"):
        cfg[k] = int(cfg[k])
    return cfg


def write_par(path, names, theta, f, gmax):
    with open(path, "w") as fh:
        fh.write(f"# Number of parameters = {len(theta)}  Objective function value = {f:.6g}  "
                 f"Maximum gradient component = {gmax:.5e}\n")
        for n, v in zip(names, theta):
            fh.write(f"# {n}:\n{v:.11f}\n")


def main(dat=None, out_dir="."):
    cfg = read_dat(dat) if dat else {"nobs": 1_000_003, "method": 1, "kernel_code": None, "gpu_index": 0}
    if cfg["method"] != 1:
        raise SystemExit("only gradient method 1 (ad4cl_device) exists in the CUDA build: there is no host path")
    n = cfg["nobs"]
    ctx = host.Context(cfg["gpu_index"])
    x, y, start = synth.simple_cpp_recipe(n)
    kernel_file = cfg.get("kernel_code")
    if kernel_file and os.path.exists(os.path.join(os.path.dirname(dat), kernel_file)):
        text = open(os.path.join(os.path.dirname(dat), kernel_file)).read()
        k = ctx.from_source(text, "AD", "GSVVDDOi")                      # simple.cpp:138-175, 232
    else:
        k = ctx.builtin("linreg2_p")
    k.set_args(None, None, host.params(0), host.params(1), ctx.buffer(x), ctx.buffer(y), None, n)   # :240-262
    k.configure(n, max_ops=4, max_slots=6, local_size=64, epilogue="half_n_log")                    # :235-238, :365
    t0 = time.perf_counter()
    theta, res = k.minimize(np.array(start), gradient_tolerance=1e-6)
    dt = time.perf_counter() - t0
    k.free()
    ctx.close()
    par = os.path.join(out_dir, "simple.par")
    write_par(par, ["a", "b"], theta, res["f"], res["gradient_max"])
    print(f"AD4CL_Device(CUDA)  {dt:.3f} s  {res['f']:.4e}  {res['gradient_max']:.4e}  "
          f"({res['evaluations']} evaluations, {res['iterations']} iterations) -> {par}")
    return theta, res


if __name__ == "__main__":
    main(*(sys.argv[1:3]))
