#!/usr/bin/env python
"""Dense-contraction path of config 2 at several sizes: C = A B, sum C, dL/dA = G B^T, dL/dB = A^T G (3 GEMMs,
6 n^3 flop) per call, CUDA events, every variant.  `python examples/bench_dense.py [n ...]`"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host  # noqa: E402


def main(sizes, reps=10):
    torch.cuda.set_device(0)
    ctx = host.Context(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx.use_stream(st.cuda_stream)
    peak = ctx.measure_fp64_peak()
    rows = []
    for n in sizes:
        rng = np.random.default_rng(n)
        A_t = torch.from_numpy(rng.uniform(-1, 1, n * n)).cuda()
        B_t = torch.from_numpy(rng.uniform(-1, 1, n * n)).cuda()
        dA_t, dB_t, C_t = (torch.empty(n * n, dtype=torch.float64, device="cuda") for _ in range(3))
        tot = torch.zeros(1, dtype=torch.float64, device="cuda")
        ws = torch.empty(n * n + (n // 64) ** 2, dtype=torch.float64, device="cuda")
        wrap = lambda t: ctx.wrap(t.data_ptr(), t.numel() * 8)  # noqa: E731
        ref = None
        for variant, name in ((1, "dmma, one launch per product"), (2, "dmma, one persistent launch")):
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
            got = (C_t.clone(), dA_t.clone(), dB_t.clone(), float(tot.item()))
            if ref is None:
                ref = got
            same = all(torch.equal(a, b) for a, b in zip(ref[:3], got[:3]))
            tf = 3 * 2 * n ** 3 / (ms * 1e-3) / 1e12
            rows.append({"n": n, "variant": variant, "what": name, "ms": ms, "tflops": tf, "frac_of_measured_fp64_peak": tf / peak,
                         "bit_identical_to_variant_1": same, "sumC": got[3]})
    print(json.dumps({"fp64_fma_peak_tflops_measured": peak, "rows": rows}))


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [1024, 2048, 4096])
