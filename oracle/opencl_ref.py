"""The reference's UNMODIFIED OpenCL path on a real OpenCL device -- TEST INFRASTRUCTURE ONLY.

The GPU box carries NVIDIA's OpenCL platform (libnvidia-opencl.so.1; the ICD loader finds it with
OCL_ICD_FILENAMES because /etc/OpenCL/vendors is absent).  This module drives it through ctypes and
follows the reference's host protocol literally (test/admb/simple.cpp:111-268 set-up, :283-317
launch and read-back): program = ad.cl + kernel file concatenated as text and built at run time,
`gs` written as a few ints, NDRange with local 64 and global ceil(N/64+1)*64, then the whole stack,
`gs` and out[] read back.  The host half (gpu_restore, host sum, epilogue, compute_gradient) is the
reference's own ad4cl.h compiled in oracle/_ref/libad4cl_ref.so (ref_finish_stack).

Sources come from oracle/_ref/cl/ (copies made by oracle/Makefile where /root/reference is
mounted; git-ignored, they travel to the GPU box with the snapshot).
"""
from __future__ import annotations

import ctypes
import os
import time
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_longlong, c_size_t, c_uint, c_ulonglong, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CL_DIR = os.path.join(HERE, "_ref", "cl")
_dp = POINTER(c_double)

_cl = None


def _lib():
    global _cl
    if _cl is None:
        os.environ.setdefault("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1")
        lib = ctypes.CDLL("libOpenCL.so.1")
        for name in ("clCreateContext", "clCreateCommandQueue", "clCreateProgramWithSource", "clCreateKernel",
                     "clCreateBuffer"):
            getattr(lib, name).restype = c_void_p
        _cl = lib
    return _cl


def available() -> bool:
    """True when an OpenCL GPU device and the reference's .cl sources are both present."""
    if not os.path.exists(os.path.join(CL_DIR, "ad.cl")):
        return False
    try:
        cl = _lib()
        n = c_uint(0)
        if cl.clGetPlatformIDs(0, None, byref(n)) != 0 or n.value == 0:
            return False
        plats = (c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, plats, None)
        nd = c_uint(0)
        return cl.clGetDeviceIDs(c_void_p(plats[0]), 4, 0, None, byref(nd)) == 0 and nd.value > 0
    except OSError:
        return False


def _check(err, what):
    if err != 0:
        raise RuntimeError(f"OpenCL {what} failed: {err}")


class OpenCLReference:
    """quirks=True: the unmodified ad.cl; False: the copy patched at the cited lines (ad_fixed.cl)."""
    kind = "reference-opencl"

    def __init__(self, quirks: bool = True):
        cl = self.cl = _lib()
        n = c_uint(0)
        _check(cl.clGetPlatformIDs(0, None, byref(n)), "clGetPlatformIDs")
        plats = (c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, plats, None)
        dev = c_void_p()
        _check(cl.clGetDeviceIDs(c_void_p(plats[0]), 4, 1, byref(dev), None), "clGetDeviceIDs")
        self.dev = dev
        err = c_int()
        self.ctx = c_void_p(cl.clCreateContext(None, 1, byref(dev), None, None, byref(err)))
        _check(err.value, "clCreateContext")
        self.queue = c_void_p(cl.clCreateCommandQueue(self.ctx, dev, c_ulonglong(2), byref(err)))   # profiling on, simple.cpp:194-198
        _check(err.value, "clCreateCommandQueue")
        self.quirks = quirks
        self.ad_cl = open(os.path.join(CL_DIR, "ad.cl" if quirks else "ad_fixed.cl")).read()
        self._programs = {}
        import oracle_api as O
        self.host = O.Reference(quirks).lib
        name = ctypes.create_string_buffer(256)
        cl.clGetDeviceInfo(dev, 0x102B, 256, name, None)
        self.device_name = name.value.decode()

    # -- plumbing ---------------------------------------------------------------------------
    def _kernel(self, kernel_text: str, entry: str):
        key = (hash(kernel_text), entry)
        if key not in self._programs:
            cl = self.cl
            src = (self.ad_cl + "\n" + kernel_text).encode()       # simple.cpp:138-161: ad.cl + user file as one string
            err = c_int()
            sp = c_char_p(src)
            ln = c_size_t(len(src))
            prog = c_void_p(cl.clCreateProgramWithSource(self.ctx, 1, byref(sp), byref(ln), byref(err)))
            _check(err.value, "clCreateProgramWithSource")
            rc = cl.clBuildProgram(prog, 1, byref(self.dev), None, None, None)   # no options, simple.cpp:191
            if rc != 0:
                size = c_size_t()
                cl.clGetProgramBuildInfo(prog, self.dev, 0x1183, 0, None, byref(size))
                log = ctypes.create_string_buffer(size.value + 1)
                cl.clGetProgramBuildInfo(prog, self.dev, 0x1183, size.value, log, None)
                raise RuntimeError(f"clBuildProgram failed ({rc}):\n{log.value.decode(errors='replace')[-4000:]}")
            k = c_void_p(cl.clCreateKernel(prog, entry.encode(), byref(err)))
            _check(err.value, f"clCreateKernel({entry})")
            self._programs[key] = k
        return self._programs[key]

    def _buffer(self, nbytes: int, data: np.ndarray | None = None):
        err = c_int()
        if data is not None:
            data = np.ascontiguousarray(data)
            nbytes = max(nbytes, data.nbytes)
        b = c_void_p(self.cl.clCreateBuffer(self.ctx, c_ulonglong(1), c_size_t(max(nbytes, 16)), None, byref(err)))
        _check(err.value, "clCreateBuffer")
        if data is not None:
            _check(self.cl.clEnqueueWriteBuffer(self.queue, b, 1, c_size_t(0), c_size_t(data.nbytes),
                                                data.ctypes.data_as(c_void_p), 0, None, None), "clEnqueueWriteBuffer")
        return b

    def _read(self, buf, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        _check(self.cl.clEnqueueReadBuffer(self.queue, buf, 1, c_size_t(0), c_size_t(nbytes),
                                           out.ctypes.data_as(c_void_p), 0, None, None), "clEnqueueReadBuffer")
        return out

    def _run(self, kernel, args, global_size, local_size):
        """setArg + enqueueNDRangeKernel + event.wait(); returns the profiled kernel time in ms."""
        cl = self.cl
        for i, a in enumerate(args):
            if isinstance(a, c_void_p):
                _check(cl.clSetKernelArg(kernel, i, c_size_t(8), byref(a)), f"clSetKernelArg({i})")
            else:
                v = c_int(int(a))
                _check(cl.clSetKernelArg(kernel, i, c_size_t(4), byref(v)), f"clSetKernelArg({i})")
        dims = len(global_size)
        g = (c_size_t * dims)(*global_size)
        l = (c_size_t * dims)(*local_size)
        ev = c_void_p()
        _check(cl.clEnqueueNDRangeKernel(self.queue, kernel, dims, None, g, l, 0, None, byref(ev)), "clEnqueueNDRangeKernel")
        _check(cl.clWaitForEvents(1, byref(ev)), "clWaitForEvents")
        t0, t1 = c_ulonglong(), c_ulonglong()
        cl.clGetEventProfilingInfo(ev, 0x1282, 8, byref(t0), None)
        cl.clGetEventProfilingInfo(ev, 0x1283, 8, byref(t1), None)
        cl.clReleaseEvent(ev)
        return (t1.value - t0.value) * 1e-6

    def _release(self, *bufs):
        for b in bufs:
            if b is not None:
                self.cl.clReleaseMemObject(b)

    # -- the reference protocol ------------------------------------------------------------
    def _protocol(self, kernel, make_args, nparams, theta, n, global_size, local_size, stack_entries, epilogue):
        """write gs/params -> NDRange -> read gs, stack, out -> reference host finish."""
        gs = np.zeros(8, dtype=np.int32)            # device layout ad.cl:81-89: ptr, 4 ints, ptr (32 bytes)
        gs[2], gs[3], gs[4], gs[5] = nparams, 0, 1, 0   # current_ad_variable_id, stack_current, recording, counter
        gs_d = self._buffer(32, gs)
        stack_d = self._buffer(40 * stack_entries)
        params = np.zeros(nparams, dtype=[("value", "<f8"), ("id", "<i4"), ("pad", "<i4")])
        params["value"], params["id"] = theta, np.arange(nparams)      # ad_init_var, ad4cl.h:90-93
        out_d = self._buffer(16 * max(n, 1))
        extra, args = make_args(gs_d, stack_d, params, out_d)
        t0 = time.perf_counter()
        kernel_ms = self._run(kernel, args, global_size, local_size)
        gs_back = self._read(gs_d, 32).view(np.int32)
        counter = int(gs_back[5])
        if counter + n + 2 > stack_entries:
            raise RuntimeError(f"the reference's stack overflowed: {counter} entries recorded, {stack_entries} allocated")
        t1 = time.perf_counter()
        stack = np.zeros(40 * (counter + n + 2), dtype=np.uint8)
        stack[: 40 * counter] = self._read(stack_d, 40 * counter)       # simple.cpp:315 reads the WHOLE stack
        out = self._read(out_d, 16 * n)
        t2 = time.perf_counter()
        f, ent, grad, t = c_double(), c_longlong(), np.zeros(nparams), np.zeros(3)
        self.host.ref_finish_stack(stack.ctypes.data_as(c_void_p), counter, nparams, out.ctypes.data_as(c_void_p), n,
                                   epilogue, byref(f), grad.ctypes.data_as(_dp), byref(ent), t.ctypes.data_as(_dp))
        t3 = time.perf_counter()
        self._release(gs_d, stack_d, out_d, *extra)
        return {"f": f.value, "grad": grad, "entries": ent.value, "counter": counter, "kernel_ms": kernel_ms,
                "launch_ms": (t1 - t0) * 1e3, "readback_ms": (t2 - t1) * 1e3, "host_ms": (t3 - t2) * 1e3,
                "total_ms": (t3 - t0) * 1e3, "readback_bytes": 40 * counter + 16 * n + 32}

    def eval_simple(self, variant, a, b, x, y):
        """variant 0: kernel.cl (Q10 rename), 1: test/admb/simple.cl -- f = (n/2) log sum (a x + b - y)^2."""
        text = open(os.path.join(CL_DIR, "kernel_renamed.cl" if variant == 0 else "simple.cl")).read()
        k = self._kernel(text, "AD")
        n = len(x)

        def make_args(gs_d, stack_d, params, out_d):
            a_d, b_d = self._buffer(16, params[0:1]), self._buffer(16, params[1:2])
            x_d, y_d = self._buffer(8 * n, np.asarray(x, dtype=np.float64)), self._buffer(8 * n, np.asarray(y, dtype=np.float64))
            return (a_d, b_d, x_d, y_d), [gs_d, stack_d, a_d, b_d, x_d, y_d, out_d, n]

        gsz = (n + 63) // 64 * 64 + 64                                    # simple.cpp:235-238
        return self._protocol(k, make_args, 2, [a, b], n, (gsz,), (64,), 4 * n + n + 16, 1)

    def eval_objective(self, name, theta, d0, d1, size, i0=0, i1=0, *, epilogue=0, ops_per_item=128):
        """A kernel of ad4cl_b200/kernels/objectives.cl against the reference's ad.cl, on the OpenCL device."""
        text = open(os.path.join(HERE, "..", "ad4cl_b200", "kernels", "objectives.cl")).read()
        k = self._kernel(text, name)
        theta = np.asarray(theta, dtype=np.float64)
        P = len(theta)

        def make_args(gs_d, stack_d, params, out_d):
            th_d = self._buffer(16 * P, params)
            d0_d = self._buffer(8, None) if d0 is None else self._buffer(0, np.asarray(d0, dtype=np.float64))
            d1_d = self._buffer(8, None) if d1 is None else self._buffer(0, np.asarray(d1, dtype=np.float64))
            if name in ("linreg2_g", "linreg2_p"):
                a_d, b_d = self._buffer(16, params[0:1]), self._buffer(16, params[1:2])
                return (th_d, d0_d, d1_d, a_d, b_d), [gs_d, stack_d, a_d, b_d, d0_d, d1_d, out_d, size]
            return (th_d, d0_d, d1_d), [gs_d, stack_d, th_d, d0_d, d1_d, out_d, size, i0, i1]

        gsz = (size + 63) // 64 * 64
        return self._protocol(k, make_args, P, theta, size, (gsz,), (64,), ops_per_item * size + size + 64, epilogue)

    def matrix_forward(self, n, A, B):
        """test/matrix/matrixmul.cl (GLOBAL variant) as matrix_mul.cpp:210-243 launches it: forward values of C
        and the device entry counter (the gradient is not defined there: ids collide)."""
        text = open(os.path.join(CL_DIR, "matrixmul.cl")).read()
        k = self._kernel(text, "matrixMult")
        var_t = np.dtype([("value", "<f8"), ("id", "<i4"), ("pad", "<i4")])
        a = np.zeros(n * n, dtype=var_t)
        b = np.zeros(n * n, dtype=var_t)
        a["value"], a["id"] = A, np.arange(n * n)
        b["value"], b["id"] = B, n * n + np.arange(n * n)
        gs = np.zeros(8, dtype=np.int32)
        gs[2], gs[4] = 2 * n * n, 1
        gs_d, stack_d = self._buffer(32, gs), self._buffer(40 * (2 * n * n * n + 16))
        a_d, b_d, c_d = self._buffer(0, a), self._buffer(0, b), self._buffer(16 * n * n)
        ms = self._run(k, [gs_d, stack_d, a_d, b_d, c_d, n, n], (n, n), (16, 16))
        counter = int(self._read(gs_d, 32).view(np.int32)[5])
        C = self._read(c_d, 16 * n * n).view(var_t)["value"].copy()
        self._release(gs_d, stack_d, a_d, b_d, c_d)
        return {"C": C, "counter": counter, "kernel_ms": ms}
