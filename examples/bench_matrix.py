#!/usr/bin/env python
"""Config 2 (test/matrix, n x n fp64): the tape path against the dense-contraction path, DFMA
against DMMA, each timed with CUDA events.  `python examples/bench_matrix.py [n]`"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host, synth, workloads as W  # noqa: E402


def main(n=1024, reps=10):
    torch.cuda.set_device(0)
    ctx = host.Context(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx.use_stream(st.cuda_stream)
    A, B = synth.msvcrt_rand_matrices(n)
    dA_t, dB_t, C_t = (torch.empty(n * n, dtype=torch.float64, device="cuda") for _ in range(3))
    A_t, B_t = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    tot = torch.zeros(1, dtype=torch.float64, device="cuda")
    ws = torch.empty(n * n + (n // 64) ** 2, dtype=torch.float64, device="cuda")
    wrap = lambda t: ctx.wrap(t.data_ptr(), t.numel() * 8)  # noqa: E731
    out = {"n": n, "fp64_fma_peak_tflops_measured": ctx.measure_fp64_peak()}
    for variant, name in ((0, "dense_dfma"), (1, "dense_dmma")):
        def step():
            ctx.matrix_product_dense(n, wrap(A_t), wrap(B_t), C=wrap(C_t), dA=wrap(dA_t), dB=wrap(dB_t),
                                     total=wrap(tot), workspace=wrap(ws), variant=variant)
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            step()
        e1.record(st)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms": ms, "tflops": 3 * 2 * n ** 3 / (ms * 1e-3) / 1e12, "sumC": float(tot.item())}
    w = W.describe("matrix_product", n)
    k = W.bind(ctx, w)
    k.profile(True)
    for _ in range(3):
        f, g = k.eval(w.theta)
    k.profile_read(reset=True)
    for _ in range(reps):
        f, g = k.eval(w.theta)
    ms, _ = k.profile_read()
    out["tape"] = {"ms": ms, "algorithmic_GBps": w.algorithmic_bytes() / (ms * 1e-3) / 1e9, "sumC": f}
    k.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
