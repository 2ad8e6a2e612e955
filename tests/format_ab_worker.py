"""Worker of tests/test_gpu_format_ab.py: evaluates a fixed set of workloads with the library named by
AD4CL_CUDA_LIB and prints objective and gradient as hex floats (exact comparison between builds)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ad4cl_b200 import host, workloads as W  # noqa: E402

CASES = (("catch_at_age", 120000, {}, {}), ("catch_at_age", 4000, {}, {"window": 4}),
         ("tape_stress", 5003, {"steps": 150}, {}), ("regression_linear", 30001, {}, {}),
         ("regression_logistic", 30001, {}, {"acc": "warp"}), ("regression_logistic", 777, {}, {"tape": "shared", "local_size": 64}),
         ("linreg2_simple", 100003, {}, {}), ("op_zoo", 999, {}, {}))


def main():
    ctx = host.Context(0)
    for name, n, kw, cfg in CASES:
        w = W.describe(name, n, **kw)
        if name == "op_zoo":
            k = ctx.builtin("op_zoo")
            k.set_args(None, None, host.params(0, 3), ctx.buffer(w.d0), None, None, w.size, 0, 0)
            k.configure(w.size, max_ops=48, max_slots=96)
        else:
            k = W.bind(ctx, w, **cfg)
        f, g = k.eval(w.theta)
        st = k.stats()
        k.free()
        kind = "exact" if len(w.theta) <= 16 and cfg.get("acc", "auto") in ("auto", "thread") else "close"
        print(name, n, kind, float(f).hex(), " ".join(float(x).hex() for x in np.asarray(g)), int(st["ops"]), int(st["slots"]))
    ctx.close()


if __name__ == "__main__":
    main()
