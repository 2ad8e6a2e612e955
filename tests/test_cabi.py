"""The C-ABI library: it loads, exports every symbol include/ad4cl_cuda.h declares,
fails loudly without a device, and the host-side logic around it behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "ad4cl_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ad4cl_cuda_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ad4cl_b200 import host
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(host.lib, s), f"{s} declared in include/ad4cl_cuda.h but not exported"


def test_builtin_table_and_signatures():
    from ad4cl_b200 import host
    b = host.builtin_kernels()
    for k in ("linreg2_g", "linreg2_p", "regression_linear", "regression_logistic", "catch_at_age",
              "tape_stress", "op_zoo", "matrix_product", "matrix_elementwise"):
        assert k in b
    assert b["linreg2_p"] == "GSVVDDOi"          # the parameter list of test/admb/simple.cl:106-112
    assert b["matrix_product"] == "GSVVOii"      # test/matrix/matrixmul.cl:5-11


def test_no_cpu_fallback():
    """Without a CUDA device init must fail with an error, never compute on the host."""
    import torch
    from ad4cl_b200 import host
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(host.Ad4clError) as e:
        host.Context(0)
    assert e.value.code == host.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "ad4cl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cl")):
                text = open(os.path.join(root, f)).read()
                bad = re.findall(r"^\s*(?:import|from)\s+oracle|#\s*include\s*[<\"][^>\"]*(?:oracle|ad_port)|"
                                 r"CDLL\([^)]*(?:oracle|_ref)|libad4cl_(?:ref|port)\.so", text, flags=re.M)
                assert not bad, (f, bad)


def test_struct_layouts_match_header():
    """ctypes mirrors of ad4cl_kernel_config / ad4cl_eval_stats vs a C compiler's view."""
    import subprocess
    import tempfile
    from ad4cl_b200 import host
    src = r'''
#include <stdio.h>
#include "ad4cl_cuda.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ad4cl_kernel_config), sizeof(ad4cl_eval_stats),
    __builtin_offsetof(ad4cl_kernel_config, epilogue_n), __builtin_offsetof(ad4cl_eval_stats, kernel_ms),
    sizeof(ad4cl_lbfgs_options), sizeof(ad4cl_lbfgs_result)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(REPO, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(host.KernelConfig)
    assert int(out[4]) == ctypes.sizeof(host.LbfgsOptions) and int(out[5]) == ctypes.sizeof(host.LbfgsResult)
    assert int(out[1]) == ctypes.sizeof(host.EvalStats)
    assert int(out[2]) == host.KernelConfig.epilogue_n.offset
    assert int(out[3]) == host.EvalStats.kernel_ms.offset


def test_config_defaults():
    """ad4cl_cuda_kernel_config_default is pure host code: the AUTO choices (accumulator, tape tier, sweep
    prefetch) are the -1 of their enums and the far plane is on."""
    from ad4cl_b200 import host
    cfg = host.KernelConfig()
    assert host.lib.ad4cl_cuda_kernel_config_default(ctypes.byref(cfg)) == 0
    assert (cfg.acc_mode, cfg.tape_tier, cfg.sweep_prefetch) == (host.ACC["auto"], host.TAPE["auto"], host.PREFETCH["auto"])
    assert cfg.recorder == host.RECORDER["auto"]
    assert cfg.far_plane == 1 and cfg.global_size[1] == 1 and cfg.max_ops > 0
    assert host.lib.ad4cl_cuda_kernel_config_default(None) != 0
    assert b"NULL" in host.lib.ad4cl_cuda_last_error()


def test_workload_bookkeeping():
    from ad4cl_b200 import synth, workloads as W
    n = 8000
    w = W.describe("catch_at_age", n)
    parts = [W.describe("catch_at_age", n, rank=r, world=3) for r in range(3)]
    assert sum(p.size for p in parts) == n
    assert sum(p.ops_total for p in parts) == w.ops_total
    assert sum(p.slots_total for p in parts) == w.slots_total
    assert all(np.array_equal(p.theta, w.theta) for p in parts)
    # replicate slices of every (year, age) cell: shards are balanced to within one replicate per cell
    R = n // 400
    assert [p.i1 for p in parts] == [0, 7, 14] and [p.i0 for p in parts] == [7, 7, 6] and R == 20
    full = w.d0.reshape(400, R)
    for p in parts:       # counter-based generators shard exactly
        assert np.array_equal(p.d0.reshape(400, p.i0), full[:, p.i1:p.i1 + p.i0])
    assert w.algorithmic_bytes() == 24 * w.slots_total + 8 * n + 8 * 101
    parts = [W.describe("regression_linear", 1000, rank=r, world=3) for r in range(3)]
    assert np.array_equal(np.concatenate([p.d1 for p in parts]), W.describe("regression_linear", 1000).d1)
    for n, world in ((10, 3), (7, 8), (1000, 8)):
        rs = [W.shard_range(n, r, world) for r in range(world)]
        assert rs[0][0] == 0 and sum(c for _, c in rs) == n
        assert all(rs[i][0] + rs[i][1] == rs[i + 1][0] for i in range(world - 1))
