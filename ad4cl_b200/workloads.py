"""The BASELINE.json configurations as (data, kernel arguments, tape budget) recipes.

`describe(name, n, ...)` produces host-side inputs only (numpy; usable on a CPU
box, by the oracle and by the tests); `bind(ctx, desc)` uploads them and returns
a configured `host.Kernel`.  Nothing here computes an objective.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import synth


@dataclass
class Workload:
    name: str                 # workload key
    kernel: str               # kernel name (objectives.cl / matrix.cl)
    theta: np.ndarray         # evaluation point, P doubles
    d0: np.ndarray | None
    d1: np.ndarray | None
    size: int                 # data work-items (this shard)
    i0: int = 0
    i1: int = 0
    epilogue: str = "sum"
    epilogue_n: float = 0.0   # n of the epilogue when it is not `size` (sharded runs)
    max_ops: int = 64         # per work-item
    max_slots: int = 0
    ops_total: int = 0        # per evaluation, for the roofline arithmetic
    slots_total: int = 0
    cells_total: int = 0      # operand slots whose partial is not a compile-time +1 / -1 (they own a tape cell)
    input_bytes: int = 0      # data bytes read once per evaluation
    linreg2: bool = False     # (a, b) passed as two separate V arguments (simple.cl signature)
    matrix: int = 0           # n of an n x n matrix kernel (A, B as 2 n^2 independents, 2-D NDRange)
    extra: dict = field(default_factory=dict)

    @property
    def nparams(self) -> int:
        return len(self.theta)

    def algorithmic_bytes(self) -> int:
        """24 B per operand slot (12 written forward + 12 read in
        reverse) + inputs read once + 8 (P+1) out."""
        return 24 * self.slots_total + self.input_bytes + 8 * (self.nparams + 1)

    def minimal_record_bytes(self) -> int:
        """What the round-2 tape stream moves when a work-item is warp-uniform (ad4cl_tape.cuh): an 8-byte
        header per WARP row (0.25 B per lane and slot) and an 8-byte cell per lane only for non-unit
        partials, each written once and read once, + inputs + 8 (P+1) out."""
        cells = self.cells_total if self.cells_total else self.slots_total
        return 2 * (8 * cells + self.slots_total // 4) + self.input_bytes + 8 * (self.nparams + 1)


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous observation shard [start, start+count) of rank `rank`."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def describe(name: str, n: int, *, rank: int = 0, world: int = 1, steps: int = 33, p: int = 10) -> Workload:
    start, count = shard_range(n, rank, world)
    if name in ("linreg2_main", "linreg2_simple", "linreg2_private"):
        gen = synth.main_cpp_recipe if name == "linreg2_main" else synth.simple_cpp_recipe
        x, y, (a, b) = gen(n, start, count)
        if name == "linreg2_private":   # the 4 operators on a private stack, 2 pre-accumulated operators on the tape
            return Workload(name, "linreg2_private", np.array([a, b]), x, y, count, epilogue="half_n_log",
                            epilogue_n=n, max_ops=4, max_slots=6, ops_total=2 * count, slots_total=3 * count,
                            input_bytes=16 * count, linreg2=True)
        return Workload(name, "linreg2_p", np.array([a, b]), x, y, count, epilogue="half_n_log", epilogue_n=n,
                        max_ops=4, max_slots=6, ops_total=4 * count, slots_total=6 * count, cells_total=3 * count,
                        input_bytes=16 * count, linreg2=True)
    if name in ("regression_linear", "regression_logistic"):
        kind = name.split("_")[1]
        X, y, beta = synth.regression(n, p, kind, start, count)
        ops = (2 * p + 1) if kind == "linear" else (2 * p + 4)
        slots = (3 * p + 1) if kind == "linear" else (3 * p + 4)
        cells = (p + 2) if kind == "linear" else (p + 3)      # the rest are +1 / -1 partials of plus / minus
        return Workload(name, name, beta, X, y, count, i0=p,
                        epilogue="half_n_log" if kind == "linear" else "sum", epilogue_n=n,
                        max_ops=ops, max_slots=slots, ops_total=ops * count, slots_total=slots * count,
                        cells_total=cells * count, input_bytes=8 * (p + 1) * count)
    if name == "catch_at_age":
        # tape length varies 4x with age, so a contiguous observation range would give the ranks different
        # mixes.  When the 400 (year, age) cells divide evenly, every rank takes WHOLE cells, dealt round-robin
        # (cell = rank + world j: every age and a spread of years on every rank, and the long runs of identical
        # work-items that the warp-uniform recorder and the lock-step work-groups live on stay as long as on one
        # GPU); otherwise a slice of the REPLICATES of every cell (same mix on every rank, shorter runs).
        ncell = synth.CAA_YEARS * synth.CAA_AGES
        R = n // ncell
        per_cell = _caa_per_cell()
        if world > 1 and ncell % world == 0 and world < 1024:
            mine = np.arange(rank, ncell, world)
            obs, theta0, _ = synth.catch_at_age(n, 0, R, cells=mine)
            ops = R * sum(per_cell[c][0] for c in mine)
            slots = R * sum(per_cell[c][1] for c in mine)
            cells = R * sum(per_cell[c][2] for c in mine)
            return Workload(name, name, theta0, obs, np.array([float(R)]), len(obs), i0=R, i1=-(rank + 1024 * world) - 1,
                            max_ops=128, max_slots=176, ops_total=ops, slots_total=slots, cells_total=cells,
                            input_bytes=8 * len(obs))
        rep0, nrep = shard_range(R, rank, world)
        obs, theta0, _ = synth.catch_at_age(n, rep0, nrep)
        ops, slots = caa_totals(nrep)
        cells = nrep * sum(c for _, _, c in per_cell)
        return Workload(name, name, theta0, obs, np.array([float(R)]), len(obs), i0=nrep, i1=rep0, max_ops=128, max_slots=176,
                        ops_total=ops, slots_total=slots, cells_total=cells, input_bytes=8 * len(obs))
    if name == "tape_stress":
        x, theta = synth.tape_stress(n, start, count)
        theta = np.ones(p)
        ops, slots = 3 * steps + 1, 4 * steps + 1
        return Workload(name, name, theta, x, None, count, i0=steps, i1=p, max_ops=ops, max_slots=slots,
                        ops_total=ops * count, slots_total=slots * count, cells_total=(2 * steps + 1) * count,
                        input_bytes=8 * count, extra={"steps": steps})
    if name in ("matrix_product", "matrix_elementwise"):
        # config 2: test/matrix (matrix_mul.cpp:84-90 inputs); n is the matrix dimension here
        A, B = synth.msvcrt_rand_matrices(n)
        op = steps if name == "matrix_elementwise" else 0      # elementwise op selector rides on `steps`
        if name == "matrix_product":
            ops, slots = 2 * n + 1, 4 * n + 1
        else:
            ops, slots = {0: (1, 2), 1: (1, 2), 2: (1, 2), 3: (1, 1), 4: (2, 2)}[op]
        return Workload(name, name, np.concatenate([A, B]), None, None, n * n, i0=n, i1=op if name == "matrix_elementwise" else n,
                        max_ops=ops, max_slots=slots, ops_total=ops * n * n, slots_total=slots * n * n,
                        input_bytes=2 * n * n * 16, matrix=n)
    if name == "op_zoo":
        d0, theta = synth.op_zoo(n)
        return Workload(name, name, theta, d0[start:start + count], None, count, max_ops=48, max_slots=96,
                        ops_total=47 * count, slots_total=0, input_bytes=8 * count)
    raise KeyError(name)


def tape_stress_with_steps(w: Workload, steps: int) -> Workload:
    """The same tape_stress shard with another tape length (the data do not depend on it)."""
    import dataclasses
    ops, slots = 3 * steps + 1, 4 * steps + 1
    return dataclasses.replace(w, i0=steps, max_ops=ops, max_slots=slots, ops_total=ops * w.size,
                               slots_total=slots * w.size, cells_total=(2 * steps + 1) * w.size, extra={"steps": steps})


def caa_totals(nrep: int) -> tuple[int, int]:
    """AD operators and operand slots of catch_at_age over `nrep` replicates of every cell."""
    per_cell = _caa_per_cell()
    return nrep * sum(o for o, _, _ in per_cell), nrep * sum(sl for _, sl, _ in per_cell)


_CAA_CACHE: list | None = None


def _caa_per_cell():
    """(ops, slots, cells) one work-item of cell = year + 40*age records; counted from the
    kernel text of catch_at_age (objectives.cl) -- ops and slots are checked against the device's
    own counters in tests/test_gpu_parity.py.  cells = slots whose partial is not +1 / -1."""
    global _CAA_CACHE
    if _CAA_CACHE is None:
        out = []
        for age in range(synth.CAA_AGES):
            for year in range(synth.CAA_YEARS):
                o = s = c = 1
                have = False
                for kk in range(age + 1):
                    if year - age + kk >= 0:
                        o += 9
                        s += 12
                        c += 8          # minus_dv, plus_vd and plus (2 slots) carry unit partials
                    if kk < age:
                        if have:
                            o += 1
                            s += 2      # plus: both unit
                        have = True
                if have:
                    o += 1
                    s += 2              # minus: both unit
                o += 18
                s += 27
                c += 21                 # of the 27: minus_dv, plus (2), minus_dv, plus (2) are unit
                out.append((o, s, c))
        # cell index = year + 40*age: the loops above ran age-major, year-minor
        _CAA_CACHE = out
    return _CAA_CACHE


def bind(ctx, w: Workload, *, quirks: bool = False, **cfg):
    """Upload the workload's data and return a configured kernel."""
    from . import host
    k = ctx.builtin(w.kernel, quirks=quirks)
    d0 = ctx.buffer(w.d0) if w.d0 is not None else None
    d1 = ctx.buffer(w.d1) if w.d1 is not None else None
    if w.matrix:
        n = w.matrix
        k.set_args(None, None, host.params(0, n * n), host.params(n * n, n * n), None, w.i0, w.i1)
        opts = dict(max_ops=w.max_ops, max_slots=w.max_slots, local_size=(16, 16), acc="global")
        opts.update(cfg)
        k.configure((n, n), **opts)
        return k
    if w.linreg2:
        k.set_args(None, None, host.params(0), host.params(1), d0, d1, None, w.size)
    else:
        k.set_args(None, None, host.params(0, w.nparams), d0, d1, None, w.size, w.i0, w.i1)
    opts = dict(max_ops=w.max_ops, max_slots=w.max_slots, epilogue=w.epilogue, epilogue_n=w.epilogue_n)
    opts.update(cfg)
    k.configure(max(w.size, 1), **opts)
    k._buffers = (d0, d1)
    return k
