"""Dense-contraction path of config 2 (ad4cl_cuda_matrix_product_dense): C = A B, sum C,
dL/dA = G B^T, dL/dB = A^T G -- against numpy, against the CPU oracle's tape gradient (host twin
of the reference, quirks off) and against this library's own tape path."""
import numpy as np
import pytest

from ad4cl_b200 import synth, workloads as W

pytestmark = pytest.mark.gpu


def run_dense(ctx, n, A, B, G=None, variant=0):
    bA, bB = ctx.buffer(A), ctx.buffer(B)
    bG = ctx.buffer(G) if G is not None else None
    bC, bdA, bdB = ctx.buffer(nbytes=8 * n * n), ctx.buffer(nbytes=8 * n * n), ctx.buffer(nbytes=8 * n * n)
    tot, ws = ctx.buffer(nbytes=8), ctx.buffer(nbytes=8 * (n * n + (n // 64) ** 2))
    ctx.matrix_product_dense(n, bA, bB, G=bG, C=bC, dA=bdA, dB=bdB, total=tot, workspace=ws, variant=variant)
    out = (bC.download().reshape(n, n), bdA.download().reshape(n, n), bdB.download().reshape(n, n), tot.download()[0])
    for b in (bA, bB, bC, bdA, bdB, tot, ws) + ((bG,) if bG else ()):
        b.free()
    return out


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["dfma", "dmma", "dmma-one-launch"])
@pytest.mark.parametrize("n", [64, 256])
def test_dense_matches_numpy_and_oracle(ctx, oracle_api, n, variant):
    A, B = synth.msvcrt_rand_matrices(n)
    A2, B2 = A.reshape(n, n), B.reshape(n, n)
    C, dA, dB, total = run_dense(ctx, n, A, B, variant=variant)
    Cn = A2 @ B2
    assert np.max(np.abs(C - Cn)) <= 1e-12 * np.abs(Cn).max()
    assert abs(total - Cn.sum()) <= 1e-12 * abs(Cn.sum())
    ones = np.ones((n, n))
    assert np.max(np.abs(dA - ones @ B2.T)) <= 1e-12 * n and np.max(np.abs(dB - A2.T @ ones)) <= 1e-12 * n
    if n == 64:      # the reference's own tape (host twin MatrixMultHost + compute_gradient), product rule fixed
        orc = oracle_api.best(False)
        ref = orc.matrix(n, A, B, want_grad=True) if orc.kind == "port" else orc.matrix(n, A, B, variant=-1, want_grad=True)
        g = np.concatenate([dA.ravel(), dB.ravel()])
        assert np.max(np.abs(g - ref["grad"])) <= 1e-12 * np.abs(ref["grad"]).max()
        assert np.max(np.abs(C.ravel() - ref["C"])) <= 1e-12 * np.abs(ref["C"]).max()


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["dfma", "dmma", "dmma-one-launch"])
def test_dense_with_general_seed(ctx, variant):
    n = 128
    A, B = synth.msvcrt_rand_matrices(n)
    G = synth.uniform(77, 0, n * n) - 0.5
    A2, B2, G2 = A.reshape(n, n), B.reshape(n, n), G.reshape(n, n)
    C, dA, dB, total = run_dense(ctx, n, A, B, G=G, variant=variant)
    assert np.max(np.abs(dA - G2 @ B2.T)) <= 1e-12 * n and np.max(np.abs(dB - A2.T @ G2)) <= 1e-12 * n


def test_dense_agrees_with_tape_path(ctx):
    n = 128
    w = W.describe("matrix_product", n)
    k = W.bind(ctx, w)
    f, g = k.eval(w.theta)
    k.free()
    C, dA, dB, total = run_dense(ctx, n, w.theta[: n * n], w.theta[n * n:])
    assert abs(total - f) <= 1e-12 * abs(f)
    assert np.max(np.abs(np.concatenate([dA.ravel(), dB.ravel()]) - g)) <= 1e-12 * np.abs(g).max()


def test_dense_rejects_bad_shapes(ctx):
    from ad4cl_b200 import host
    b = ctx.buffer(nbytes=8 * 100 * 100)
    with pytest.raises(host.Ad4clError) as e:
        ctx.matrix_product_dense(100, b, b, workspace=b)
    assert e.value.code == host.ERR_INVALID
    b.free()
