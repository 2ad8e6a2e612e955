"""Regenerates tests/golden/*.json|*.npz from the REFERENCE.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py

Produces
  * matrix256_host_txt.npz   the reference's own golden file test/matrix/host.txt
                             (== device.txt apart from the timing line), parsed:
                             C[256*256] as printed (6 significant digits) and the
                             entry count of its last line.  Only numbers are kept.
  * ref_outputs.json         f, gradient and tape length produced by the reference's
                             own ad.cl / ad4cl.h / example kernels (oracle/_ref) on the
                             seeded inputs of ad4cl_b200/synth.py, for both the
                             unmodified reference ("quirks") and the copy patched at
                             the cited lines ("fixed").  Doubles are stored as hex
                             strings so the fixtures are bit exact.
  * op_probes.json           every operator of every family on fixed operands:
                             value, result id and recorded entry, as the reference
                             produces them.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle_api as O  # noqa: E402
from ad4cl_b200 import synth  # noqa: E402

REFERENCE = "/root/reference"

OPS_BIN = ["plus", "minus", "times", "divide"]
OPS_MATH = ["cos", "sin", "tan", "acos", "asin", "atan", "cosh", "sinh", "tanh",
            "exp", "log", "log10", "sqrt"]
ALL_OPS = ([o + s for o in OPS_BIN for s in ("", "_vd", "_dv")] + ["plus_eq", "plus_eq_d"]
           + OPS_MATH + ["pow", "pow_vd", "pow_dv"])
PROBE_POINTS = [(0.3, 1.7), (0.81, 2.25), (0.05, 1.01)]


def hx(a):
    return [float(v).hex() for v in np.atleast_1d(a)]


def cases():
    """name -> (kernel, theta, d0, d1, size, i0, i1, global_size, epilogue, ops_per_item)"""
    out = {}
    for n in (1000, 4099):
        x, y, (a, b) = synth.main_cpp_recipe(n)
        gsz = ((n + 63) // 64 + 1) * 64          # simple.cpp:235-238
        for k in ("linreg2_g", "linreg2_p"):
            out[f"{k}_main_{n}"] = (k, [a, b], x, y, n, 0, 0, gsz, 1, 4)
    x, y, (a, b) = synth.simple_cpp_recipe(5003)
    out["linreg2_p_simple_5003"] = ("linreg2_p", [a, b], x, y, 5003, 0, 0, 5056, 1, 4)
    X, y, beta = synth.regression(3001, 10, "linear")
    out["regression_linear_3001"] = ("regression_linear", beta, X, y, 3001, 10, 0, 3001, 1, 21)
    X, y, beta = synth.regression(3001, 10, "logistic")
    out["regression_logistic_3001"] = ("regression_logistic", beta, X, y, 3001, 10, 0, 3001, 0, 24)
    obs, th, R = synth.catch_at_age(80000)          # R = 200 replicates per cell: all six fleets occur
    out["catch_at_age_80000"] = ("catch_at_age", th, obs, np.array([float(R)]), 80000, R, 0, 80000, 0, 130)
    obs, th, R = synth.catch_at_age(80000, 64, 72)  # a replicate slice, as a multi-GPU rank holds it
    out["catch_at_age_slice"] = ("catch_at_age", th, obs, np.array([float(R)]), 400 * 72, 72, 64, 400 * 72, 0, 130)
    for steps in (33, 200):
        x, th = synth.tape_stress(777)
        out[f"tape_stress_{steps}"] = ("tape_stress", th, x, None, 777, steps, 10, 777, 0, 3 * steps + 1)
    d0, th = synth.op_zoo(513)
    out["op_zoo_513"] = ("op_zoo", th, d0, None, 513, 0, 0, 513, 0, 48)
    return out


def main():
    # 1. the reference's own golden file
    lines = open(os.path.join(REFERENCE, "test/matrix/host.txt")).read().replace("\r", "").split("\n")
    rows = [np.array(l.split(), dtype=np.float64) for l in lines[1:257]]
    C = np.stack(rows)
    assert C.shape == (256, 256)
    entries = int(lines[257].strip())
    dev = open(os.path.join(REFERENCE, "test/matrix/device.txt")).read().replace("\r", "").split("\n")
    assert dev[1:258] == lines[1:258], "host.txt and device.txt differ beyond the timing line"
    np.savez_compressed(os.path.join(HERE, "matrix256_host_txt.npz"), C=C, entries=np.int64(entries))

    # 2. reference outputs on seeded inputs
    res = {}
    for mode, quirks in (("quirks", True), ("fixed", False)):
        ref = O.Reference(quirks)
        for name, (k, th, d0, d1, size, i0, i1, gsz, epi, opi) in cases().items():
            r = ref.eval_objective(k, th, d0, d1, size, i0, i1, global_size=gsz, epilogue=epi, ops_per_item=opi)
            res[f"{mode}/{name}"] = {"f": hx(r["f"])[0], "grad": hx(r["grad"]), "entries": r["entries"]}
        # the reference's own kernels (kernel.cl / simple.cl / host AD())
        x, y, (a, b) = synth.main_cpp_recipe(1000)
        for variant, nm in ((0, "kernel_cl"), (1, "simple_cl"), (2, "host_AD")):
            r = ref.eval_simple(variant, a, b, x, y, global_size=1088)
            res[f"{mode}/ref_{nm}_main_1000"] = {"f": hx(r["f"])[0], "grad": hx(r["grad"]), "entries": r["entries"]}
        # small matrix: host twin with gradient
        A, B = synth.msvcrt_rand_matrices(16)
        r = ref.matrix(16, A, B, variant=-1, want_grad=True)
        res[f"{mode}/matrix16_host"] = {"C": hx(r["C"]), "grad": hx(r["grad"]), "entries": r["entries"]}
    json.dump(res, open(os.path.join(HERE, "ref_outputs.json"), "w"), indent=0)

    # 3. op probes
    probes = {}
    for mode, quirks in (("quirks", True), ("fixed", False)):
        ref = O.Reference(quirks)
        for fam in ("global", "pad", "private"):
            for op in ALL_OPS:
                for (a, b) in PROBE_POINTS:
                    r = ref.op_probe(fam, op, a, b)
                    key = f"{mode}/{fam}/{op}/{a}/{b}"
                    probes[key] = None if r is None else {
                        "value": hx(r["value"])[0], "id": r["id"], "size": r["size"], "eid": r["eid"],
                        "dx": hx(r["dx"][: r["size"]]), "ids": list(r["ids"][: r["size"]])}
        for op in ALL_OPS:
            for (a, b) in PROBE_POINTS:
                r = ref.host_op_probe(op, a, b)
                key = f"{mode}/host/{op}/{a}/{b}"
                probes[key] = None if r is None else {
                    "value": hx(r["value"])[0], "id": r["id"], "size": r["size"], "eid": r["eid"],
                    "dx": hx(r["dx"][: r["size"]]), "ids": list(r["ids"][: r["size"]])}
    json.dump(probes, open(os.path.join(HERE, "op_probes.json"), "w"), indent=0)
    print("golden fixtures written:", len(res), "evaluations,", len(probes), "op probes")


if __name__ == "__main__":
    main()
