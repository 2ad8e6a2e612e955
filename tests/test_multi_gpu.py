"""The N > 1 path on real GPUs (skipped on a 1-GPU box; the host logic is covered on the CPU by
tests/test_dist_gloo.py): sharded evaluation + NCCL all-reduce, and the library's one-shot NVLink
all-reduce, against the single-GPU evaluation of the whole problem."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name,n", [("catch_at_age", 400000), ("regression_linear", 100003)])
def test_sharded_gpu_evaluation(name, n):
    world = min(_ngpus(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(REPO, "tests", "multi_gpu_worker.py"), name, str(n)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0]
    r = json.loads(line[len("RESULT "):])
    f1, g1 = r["single"][0], np.array(r["single"][1])
    scale = np.abs(g1).max()
    first = None
    for rk in r["ranks"]:
        f, g = rk["nccl"][0], np.array(rk["nccl"][1])
        assert abs(f - f1) <= 1e-12 * abs(f1) and np.abs(g - g1).max() <= 1e-12 * scale
        for f2, g2 in rk["oneshot"]:
            g2 = np.array(g2)
            assert abs(f2 - f1) <= 1e-12 * abs(f1) and np.abs(g2 - g1).max() <= 1e-12 * scale
            if first is None:
                first = (f2, g2)
            # rank-ordered sum: every rank and every epoch holds the same bits
            assert f2 == first[0] and np.array_equal(g2, first[1])
