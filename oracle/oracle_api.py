"""ctypes bindings of the CPU oracles -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs
(cpu_baseline, --impl reference); never by the ad4cl_b200 package.

Two oracles, same call shapes:
  * `Reference(quirks)`  oracle/_ref/libad4cl_ref.so / libad4cl_reffix.so: the
    reference's OWN ad.cl + ad4cl.h + example kernels compiled from
    /root/reference by oracle/Makefile (kind "reference").
  * `Port(quirks)`       oracle/libad4cl_port.so: this repo's C restatement
    (oracle/ad_port.h), pinned against the former by tests/test_oracle.py
    (kind "port").
`best(quirks)` returns the reference build when its .so is present, else the port.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_longlong, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = POINTER(c_double)

FAMILIES = {"global": 0, "pad": 1, "private": 2}


class Entry(ctypes.Structure):
    """ad_entry, 40 bytes (ad.cl:75-79)."""
    _fields_ = [("dx0", c_double), ("id0", c_int), ("_p0", c_int),
                ("dx1", c_double), ("id1", c_int), ("_p1", c_int),
                ("id", c_int), ("size", c_int)]


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def build(target: str = "all") -> None:
    """make -C oracle <target> (port always; ref only where /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class _Base:
    kind = "?"

    def eval_objective(self, name, theta, d0, d1, size, i0=0, i1=0, *, global_size=None,
                       epilogue=0, threads=1, ops_per_item=None):
        """-> dict(f, grad[P], entries, t_ms[3] = record / host sum / sweep)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        d0 = None if d0 is None else np.ascontiguousarray(d0, dtype=np.float64)
        d1 = None if d1 is None else np.ascontiguousarray(d1, dtype=np.float64)
        gsz = size if global_size is None else global_size
        cap = int((ops_per_item or 128)) * max(size, 1) + size + 64
        f = c_double()
        grad = np.zeros(len(theta))
        ent = c_longlong()
        t = np.zeros(3)
        rc = self._eval(name.encode(), len(theta), _p(theta), _p(d0), _p(d1), size, i0, i1, gsz,
                        epilogue, threads, c_longlong(cap), byref(f), _p(grad), byref(ent), _p(t))
        if rc != 0:
            raise RuntimeError(f"oracle eval_objective({name}) failed rc={rc}")
        return {"f": f.value, "grad": grad, "entries": ent.value, "t_ms": t}

    def op_probe(self, family, op, a, b, ida=7, idb=9, recording=1):
        """-> dict(value, id, entry) or None when the family lacks the operator."""
        v, lhs = c_double(), c_double()
        rid = c_int()
        e = Entry()
        rc = self._probe(FAMILIES[family], op.encode(), c_double(a), c_double(b), ida, idb,
                         recording, byref(v), byref(rid), byref(e), byref(lhs))
        if rc != 0:
            return None
        return {"value": v.value, "id": rid.value, "lhs": lhs.value,
                "size": e.size, "eid": e.id, "dx": (e.dx0, e.dx1), "ids": (e.id0, e.id1)}


class Port(_Base):
    kind = "port"

    def __init__(self, quirks: bool = True):
        path = os.path.join(HERE, "libad4cl_port.so")
        if not os.path.exists(path):
            build("port")
        self.lib = ctypes.CDLL(path)
        pre = "port_q_" if quirks else "port_f_"
        self.quirks = quirks
        self._eval = getattr(self.lib, pre + "eval_objective")
        self._probe = getattr(self.lib, pre + "op_probe")
        self._matrix = getattr(self.lib, pre + "matrix")

    def matrix(self, n, A, B, want_grad=False):
        A = np.ascontiguousarray(A, dtype=np.float64)
        B = np.ascontiguousarray(B, dtype=np.float64)
        C = np.zeros(n * n)
        g = np.zeros(2 * n * n) if want_grad else None
        ent = c_longlong()
        rc = self._matrix(n, _p(A), _p(B), _p(C), byref(ent), _p(g))
        if rc != 0:
            raise RuntimeError(f"oracle matrix failed rc={rc}")
        return {"C": C, "entries": ent.value, "grad": g}


class Reference(_Base):
    kind = "reference"

    def __init__(self, quirks: bool = True):
        path = os.path.join(HERE, "_ref", "libad4cl_ref.so" if quirks else "libad4cl_reffix.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        self.quirks = quirks
        self._eval = self.lib.ref_eval_objective
        self._probe = self.lib.ref_op_probe

    def host_op_probe(self, op, a, b, ida=7, idb=9, recording=1):
        v, lhs = c_double(), c_double()
        rid = c_int()
        e = Entry()
        rc = self.lib.ref_host_op_probe(op.encode(), c_double(a), c_double(b), ida, idb, recording,
                                        byref(v), byref(rid), byref(e), byref(lhs))
        if rc != 0:
            return None
        return {"value": v.value, "id": rid.value, "lhs": lhs.value,
                "size": e.size, "eid": e.id, "dx": (e.dx0, e.dx1), "ids": (e.id0, e.id1)}

    def eval_weighted_log(self, name, theta, d0, d1, size, weights, i0=0, i1=0, ops_per_item=64, shift=0.0):
        """f = sum_i w_i log(out_i + shift) recorded with the reference's host operators on top of the kernel's out[]
        and swept by compute_gradient -> dict(f, grad, out)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        d0 = None if d0 is None else np.ascontiguousarray(d0, dtype=np.float64)
        d1 = None if d1 is None else np.ascontiguousarray(d1, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        cap = int(ops_per_item + 5) * max(size, 1) + 64
        f = c_double()
        grad, out = np.zeros(len(theta)), np.zeros(size)
        rc = self.lib.ref_eval_weighted_log(name.encode(), len(theta), _p(theta), _p(d0), _p(d1), size, i0, i1, size,
                                            _p(w), c_double(shift), c_longlong(cap), byref(f), _p(grad), _p(out))
        if rc != 0:
            raise RuntimeError(f"ref_eval_weighted_log rc={rc}")
        return {"f": f.value, "grad": grad, "out": out}

    def eval_simple(self, variant, a, b, x, y, *, global_size=None, threads=1):
        """The reference's own config-1 kernels: variant 0 kernel.cl, 1 simple.cl, 2 host AD()."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = len(x)
        gsz = n if global_size is None else global_size
        f = c_double()
        grad = np.zeros(2)
        ent = c_longlong()
        t = np.zeros(3)
        rc = self.lib.ref_eval_simple(variant, c_double(a), c_double(b), _p(x), _p(y), n, gsz, threads,
                                      byref(f), _p(grad), byref(ent), _p(t))
        if rc != 0:
            raise RuntimeError(f"ref_eval_simple rc={rc}")
        return {"f": f.value, "grad": grad, "entries": ent.value, "t_ms": t}

    def matrix(self, n, A, B, variant=-1, want_grad=False):
        """variant -1 host twin, 0 device GLOBAL, 1 device USE_LOCAL (matrixmul.cl)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        B = np.ascontiguousarray(B, dtype=np.float64)
        C = np.zeros(n * n)
        g = np.zeros(2 * n * n) if want_grad else None
        ent = c_longlong()
        t = np.zeros(3)
        rc = self.lib.ref_matrix(variant, n, _p(A), _p(B), _p(C), byref(ent), _p(g), _p(t))
        if rc != 0:
            raise RuntimeError(f"ref_matrix rc={rc}")
        return {"C": C, "entries": ent.value, "grad": g, "t_ms": t}


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libad4cl_ref.so"))


def best(quirks: bool = True) -> _Base:
    return Reference(quirks) if have_reference() else Port(quirks)
