"""Frozen synthetic inputs for the five BASELINE.json configurations.

Every generator is counter based: element i of stream `seed` is
splitmix64(seed * 2**32 + i), so any shard of any array can be produced
independently (multi-GPU ranks generate only their own observations) and the
same numbers come out on every machine.  Uniforms are (u >> 11) * 2**-53,
normals come from Box-Muller on two such streams.

Data recipes that the reference itself defines are restated here with their
source lines:
  * main.cpp:267-274      x[i] = i, y[i] = 4.1919 x[i] + 3.2123
  * simple.cpp:58-69      x ~ 150 U(0,1), y = 2 x + 4 + 7 N(0,1)   (ADMB RNG there;
                          splitmix here -- the stream is not reproducible, the shape is)
  * matrix_mul.cpp:84-90  A then B filled from rand()/(RAND_MAX+1), MSVCRT LCG, seed 1
"""
from __future__ import annotations

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(counter: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser over uint64 counters."""
    with np.errstate(over="ignore"):
        z = (counter.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def uniform_at(seed: int, idx: np.ndarray) -> np.ndarray:
    """U[0,1) doubles for the counters `idx` (any order, any subset) of stream `seed`."""
    with np.errstate(over="ignore"):
        ctr = (np.uint64(seed) << np.uint64(40)) + idx.astype(np.uint64)
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def normal_at(seed: int, idx: np.ndarray) -> np.ndarray:
    """N(0,1) doubles by Box-Muller from streams 2*seed+1000 and 2*seed+1001 at counters `idx`."""
    u1 = uniform_at(2 * seed + 1000, idx)
    u2 = uniform_at(2 * seed + 1001, idx)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def uniform(seed: int, start: int, count: int) -> np.ndarray:
    """U[0,1) doubles for counters start .. start+count-1 of stream `seed`."""
    return uniform_at(seed, np.arange(start, start + count, dtype=np.uint64))


def normal(seed: int, start: int, count: int) -> np.ndarray:
    return normal_at(seed, np.arange(start, start + count, dtype=np.uint64))


# --------------------------------------------------------------------------- config 1
def main_cpp_recipe(n: int, start: int = 0, count: int | None = None):
    """main.cpp:267-274, 290-291: data and the first evaluation point."""
    count = n - start if count is None else count
    x = np.arange(start, start + count, dtype=np.float64)
    y = 4.1919 * x + 3.2123
    return x, y, (4.1919 - .005, 3.2123 - .0051)


def simple_cpp_recipe(n: int, start: int = 0, count: int | None = None, seed: int = 101):
    """simple.cpp:58-69 shape: x ~ 150 U, y = 2 x + 4 + 7 N(0,1); evaluate at (1.9, 3.9)."""
    count = n - start if count is None else count
    x = 150.0 * uniform(seed, start, count)
    y = 2.0 * x + 4.0 + 7.0 * normal(seed, start, count)
    return x, y, (1.9, 3.9)


# --------------------------------------------------------------------------- config 2
def msvcrt_rand_matrices(n: int):
    """A then B, n*n each, from the MSVCRT LCG seeded with 1 (the runtime the
    reference's matrix_mul.exe was built against; matrix_mul.cpp:84-90):
    s = s*214013 + 2531011 ; r = (s >> 16) & 0x7fff ; value = r / 32768."""
    total = 2 * n * n
    r = np.empty(total, dtype=np.float64)
    s = 1
    for i in range(total):          # 2 n^2 <= 2.1 M steps: ~1 s at n = 1024
        s = (s * 214013 + 2531011) & 0xFFFFFFFF
        r[i] = (s >> 16) & 0x7FFF
    r /= 32768.0
    return r[: n * n].copy(), r[n * n:].copy()


# --------------------------------------------------------------------------- config 3
def regression(n: int, p: int = 10, kind: str = "linear", start: int = 0, count: int | None = None):
    """X stored [P][count] (observation-contiguous); X[0] = 1, X[1:] ~ U(0,1) seed 3;
    beta*_j = (-1)^j (j+1)/10; linear y = X beta* + 0.5 N(0,1); logistic y ~ Bernoulli(sigmoid).
    Evaluation point beta* - 0.05."""
    count = n - start if count is None else count
    X = np.empty((p, count), dtype=np.float64)
    X[0] = 1.0
    for j in range(1, p):
        X[j] = uniform(3 * 64 + j, start, count)
    beta = np.array([(-1.0) ** j * (j + 1) / 10.0 for j in range(p)])
    eta = np.zeros(count)
    for j in range(p):            # fixed order, elementwise: a shard gets the same bits as the whole
        eta += beta[j] * X[j]
    if kind == "linear":
        y = eta + 0.5 * normal(3, start, count)
    else:
        y = (uniform(3 * 64 + 63, start, count) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return X, y, beta - 0.05


# --------------------------------------------------------------------------- config 4
CAA_YEARS, CAA_AGES, CAA_FLEETS, CAA_P = 40, 10, 6, 100


def caa_theta_true() -> np.ndarray:
    th = np.zeros(CAA_P)
    c = np.arange(49)
    th[0:49] = np.log(1000.0) + 0.3 * np.sin(0.7 * c)          # log recruitment by cohort
    yv = np.arange(CAA_YEARS)
    th[49:89] = np.log(0.25 + 0.15 * np.sin(0.3 * yv) ** 2)      # log F by year
    th[89] = 3.5                                                 # a50
    th[90] = 1.2                                                 # slope
    th[91] = np.log(0.2)                                         # log M
    th[92] = np.log(0.05)                                        # log q
    th[93] = np.log(0.3)                                         # log sigma
    th[94:100] = np.array([-0.10, 0.05, 0.00, 0.08, -0.04, 0.01])  # fleet deviations
    return th


def caa_log_chat_table(theta: np.ndarray) -> np.ndarray:
    """ln Chat for every (fleet, age, year): the model of catch_at_age in
    ad4cl_b200/kernels/objectives.cl evaluated in plain numpy."""
    M = np.exp(theta[91])
    F = np.exp(theta[49:89])
    k = np.arange(CAA_AGES, dtype=np.float64)
    sel = 1.0 / (np.exp(-(theta[90] * (k - theta[89]))) + 1.0)
    tab = np.empty((CAA_FLEETS, CAA_AGES, CAA_YEARS))
    for age in range(CAA_AGES):
        for year in range(CAA_YEARS):
            cum = 0.0
            for kk in range(age):
                yy = year - age + kk
                cum += (M + F[yy] * sel[kk]) if yy >= 0 else M
            fs = F[year] * sel[age]
            Z = M + fs
            N = np.exp(theta[year - age + 9] - cum)
            base = fs / Z * (1.0 - np.exp(-Z)) * N
            for fl in range(CAA_FLEETS):
                tab[fl, age, year] = np.log(base * np.exp(theta[92] + theta[94 + fl]))
    return tab


def catch_at_age(n: int, rep0: int = 0, nrep: int | None = None, cells=None):
    """Observed log catch for the replicate slice [rep0, rep0 + nrep) of the cells `cells` (default: all 400)
    of an n-observation problem (n = 400 R): observation (cell, rep) with cell = year + 40 age has the global
    index rep + R cell and sits at rep - rep0 + nrep * j in the returned array, j the position of its cell in
    `cells` (replicate fastest, so the operator sequence is uniform across neighbouring work-items).
    obs = ln Chat(theta*) + 0.3 N(0,1) seed 4; evaluation point theta* + 0.02 N(0,1) seed 5.
    Returns (obs, theta0, R)."""
    assert n % (CAA_YEARS * CAA_AGES) == 0, "catch-at-age needs n = 400 * replicates"
    R = n // (CAA_YEARS * CAA_AGES)
    nrep = R - rep0 if nrep is None else nrep
    th_true = caa_theta_true()
    tab = caa_log_chat_table(th_true)
    cells = np.arange(CAA_YEARS * CAA_AGES, dtype=np.int64) if cells is None else np.asarray(cells, dtype=np.int64)
    cell = np.repeat(cells, nrep)
    rep = rep0 + np.tile(np.arange(nrep, dtype=np.int64), len(cells))
    year = cell % CAA_YEARS
    age = cell // CAA_YEARS
    fleet = (rep * CAA_FLEETS) // R      # 6 contiguous blocks of replicates
    obs = tab[fleet, age, year] + 0.3 * normal_at(4, rep + R * cell)
    theta0 = th_true + 0.02 * normal(5, 0, CAA_P)
    return obs, theta0, R


# --------------------------------------------------------------------------- config 5
def tape_stress(n: int, start: int = 0, count: int | None = None):
    """x ~ U(0.5, 1.5) seed 6; theta = 1 (P = 10)."""
    count = n - start if count is None else count
    return 0.5 + uniform(6, start, count), np.ones(10)


# --------------------------------------------------------------------------- op_zoo
def op_zoo(n: int):
    """d0 in (0.1, 0.9); theta chosen so every operator of op_zoo is in its domain."""
    return 0.1 + 0.8 * uniform(9, 0, n), np.array([0.9, 1.3, 0.4])
