"""Pins the CPU oracles.

  * the port (oracle/ad_port.h) against the committed golden fixtures, which were
    produced by the reference's own code (tests/golden/make_golden.py);
  * the reference build itself (oracle/_ref), when present, against the same
    fixtures and bit-for-bit against the port;
  * both against the reference's own golden file test/matrix/host.txt
    (fixture matrix256_host_txt.npz) and against closed forms.
All comparisons here are CPU only.
"""
import json
import os

import numpy as np
import pytest

from ad4cl_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def unhex(v):
    return np.array([float.fromhex(x) for x in v]) if isinstance(v, list) else float.fromhex(v)


@pytest.fixture(scope="module")
def gold_eval():
    return json.load(open(os.path.join(GOLD, "ref_outputs.json")))


@pytest.fixture(scope="module")
def gold_probes():
    return json.load(open(os.path.join(GOLD, "op_probes.json")))


def golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("mode", ["quirks", "fixed"])
def test_port_reproduces_reference_outputs_bit_exactly(oracle_api, gold_eval, mode):
    mg = golden_cases()
    port = oracle_api.Port(quirks=(mode == "quirks"))
    for name, (k, th, d0, d1, size, i0, i1, gsz, epi, opi) in mg.cases().items():
        r = port.eval_objective(k, th, d0, d1, size, i0, i1, global_size=gsz, epilogue=epi, ops_per_item=opi)
        g = gold_eval[f"{mode}/{name}"]
        assert r["f"] == unhex(g["f"]), name
        assert np.array_equal(r["grad"], unhex(g["grad"])), name
        assert r["entries"] == g["entries"], name


@pytest.mark.parametrize("mode", ["quirks", "fixed"])
def test_reference_build_matches_its_own_golden_outputs(oracle_api, gold_eval, mode):
    if not oracle_api.have_reference():
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    mg = golden_cases()
    ref = oracle_api.Reference(quirks=(mode == "quirks"))
    for name, (k, th, d0, d1, size, i0, i1, gsz, epi, opi) in mg.cases().items():
        r = ref.eval_objective(k, th, d0, d1, size, i0, i1, global_size=gsz, epilogue=epi, ops_per_item=opi)
        g = gold_eval[f"{mode}/{name}"]
        assert r["f"] == unhex(g["f"]) and np.array_equal(r["grad"], unhex(g["grad"])), name
    # the reference's own three config-1 paths agree with each other (benchmarks.txt:9-11 claim)
    x, y, (a, b) = synth.main_cpp_recipe(1000)
    outs = [ref.eval_simple(v, a, b, x, y, global_size=1088) for v in (0, 1, 2)]
    for o in outs[1:]:
        assert o["f"] == outs[0]["f"] and np.array_equal(o["grad"], outs[0]["grad"])
        assert o["entries"] == 5 * 1000 + 2          # 4 per obs + host sum + log + times (simple.cpp:357-365)


def test_main_cpp_recipe_closed_form(oracle_api):
    """main.cpp:267-292, 417: f = (N/2) ln S, grad = (N/2)/S * sum 2 r (x, 1)."""
    n = 4099
    x, y, (a, b) = synth.main_cpp_recipe(n)
    r = (a * x + b) - y
    S = float(np.sum(r * r))
    f = n / 2 * np.log(S)
    g = n / 2 / S * np.array([np.sum(2 * r * x), np.sum(2 * r)])
    for quirks in (True, False):
        o = oracle_api.Port(quirks).eval_objective("linreg2_p", [a, b], x, y, n, global_size=4160, epilogue=1,
                                                   ops_per_item=4)
        assert abs(o["f"] - f) <= 1e-13 * abs(f)
        assert np.all(np.abs(o["grad"] - g) <= 1e-12 * np.abs(g))


@pytest.mark.parametrize("mode", ["quirks", "fixed"])
def test_op_probes_match_reference(oracle_api, gold_probes, mode):
    """Every operator of every family: value, result id, entry size/ids bit exact."""
    mg = golden_cases()
    port = oracle_api.Port(quirks=(mode == "quirks"))
    checked = 0
    for fam in ("global", "pad", "private"):
        for op in mg.ALL_OPS:
            for (a, b) in mg.PROBE_POINTS:
                g = gold_probes[f"{mode}/{fam}/{op}/{a}/{b}"]
                r = port.op_probe(fam, op, a, b)
                if g is None:
                    assert r is None, (fam, op)
                    continue
                assert r["value"] == unhex(g["value"]), (fam, op)
                assert (r["id"], r["size"], r["eid"]) == (g["id"], g["size"], g["eid"]), (fam, op)
                assert list(r["ids"][: r["size"]]) == g["ids"], (fam, op)
                assert np.array_equal(np.array(r["dx"][: r["size"]]), unhex(g["dx"])), (fam, op)
                checked += 1
    assert checked == 3 * (14 + 14 + 30 + 16)   # points x (global 30 + pad 14 + private 30 ops)


def test_quirks_are_what_the_survey_says(gold_probes):
    """Q1/Q2/Q3 of SURVEY.md 2.3 as recorded from the unmodified reference."""
    a, b = 0.3, 1.7
    q = gold_probes[f"quirks/global/times/{a}/{b}"]
    assert unhex(q["dx"]).tolist() == [a, b]                 # Q1: own value, not the other operand's
    f = gold_probes[f"fixed/global/times/{a}/{b}"]
    assert unhex(f["dx"]).tolist() == [b, a]
    assert unhex(gold_probes[f"quirks/global/times_dv/{a}/{b}"]["dx"])[0] == b   # Q2 (ad.cl:443)
    assert unhex(gold_probes[f"quirks/pad/times_dv/{a}/{b}"]["dx"])[0] == a      # pad family is right (ad.cl:840)
    assert unhex(gold_probes[f"quirks/global/cos/{a}/{b}"]["value"]) == np.log(a)  # Q3
    assert gold_probes[f"quirks/pad/cos/{a}/{b}"] is None                           # Q12: no pad math
    assert gold_probes[f"quirks/host/plus_vd/{a}/{b}"]["size"] == 2                 # Q4


def test_reference_golden_file_host_txt(oracle_api):
    """test/matrix/host.txt == device.txt: C = A*B for the MSVCRT-rand inputs, printed
    with 6 significant digits, and the tape length 33 619 968 on its last line."""
    z = np.load(os.path.join(GOLD, "matrix256_host_txt.npz"))
    A, B = synth.msvcrt_rand_matrices(256)
    C_np = (A.reshape(256, 256) @ B.reshape(256, 256))
    printed = z["C"]
    assert printed.shape == (256, 256)
    # printed with default ostream precision (6 significant digits)
    assert np.max(np.abs(C_np - printed) / printed) < 1e-5
    assert int(z["entries"]) == 256 * 256 * 512 + 256 * 256 == 33619968


@pytest.mark.parametrize("n", [16, 48])
def test_matrix_port_vs_reference_and_numpy(oracle_api, n, gold_eval):
    A, B = synth.msvcrt_rand_matrices(n)
    C_np = (A.reshape(n, n) @ B.reshape(n, n)).ravel()
    for quirks in (True, False):
        p = oracle_api.Port(quirks).matrix(n, A, B, want_grad=True)
        assert p["entries"] == n * n * 2 * n + n * n
        assert np.allclose(p["C"], C_np, rtol=1e-13)
        if not quirks:   # d(sum C)/dA[j,k] = sum_i B[k,i] ; d/dB[k,i] = sum_j A[j,k]
            gA = np.repeat(B.reshape(n, n).sum(axis=1)[None, :], n, axis=0).ravel()
            gB = np.repeat(A.reshape(n, n).sum(axis=0)[:, None], n, axis=1).ravel()
            assert np.allclose(p["grad"][: n * n], gA, rtol=1e-12)
            assert np.allclose(p["grad"][n * n:], gB, rtol=1e-12)
        if oracle_api.have_reference():
            r = oracle_api.Reference(quirks).matrix(n, A, B, variant=-1, want_grad=True)
            assert np.array_equal(r["C"], p["C"]) and r["entries"] == p["entries"]
            assert np.array_equal(r["grad"], p["grad"])
            for variant in (0, 1):      # the device kernels: forward values and tape length only (Q6)
                d = oracle_api.Reference(quirks).matrix(n, A, B, variant=variant)
                assert np.array_equal(d["C"], p["C"]) and d["entries"] == p["entries"]
        if n == 16:
            g = gold_eval[("quirks" if quirks else "fixed") + "/matrix16_host"]
            assert np.array_equal(p["C"], unhex(g["C"])) and np.array_equal(p["grad"], unhex(g["grad"]))


def test_fixed_gradient_is_the_true_gradient(oracle_api):
    """quirk-free port vs central finite differences on the catch-at-age likelihood."""
    n = 40000
    obs, th, R = synth.catch_at_age(n)
    port = oracle_api.Port(False)
    g = port.eval_objective("catch_at_age", th, obs, np.array([float(R)]), n, R, 0, ops_per_item=130)["grad"]
    for k in (0, 50, 89, 90, 91, 92, 93, 94, 96):
        e = np.zeros(100)
        e[k] = 1e-6
        fp = port.eval_objective("catch_at_age", th + e, obs, np.array([float(R)]), n, R, 0, ops_per_item=130)["f"]
        fm = port.eval_objective("catch_at_age", th - e, obs, np.array([float(R)]), n, R, 0, ops_per_item=130)["f"]
        assert abs((fp - fm) / 2e-6 - g[k]) <= 1e-6 * max(1.0, abs(g[k]))
