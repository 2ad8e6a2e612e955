"""Python host side of the CUDA AD runtime: a thin ctypes mirror of the C ABI
(include/ad4cl_cuda.h), shaped like the host protocol every AD4CL example
follows (test/admb/simple.cpp:111-268 set-up, :270-384 one evaluation):

    ctx    = Context(device)                       # platform / context / queue
    x_d    = ctx.buffer(x); y_d = ctx.buffer(y)    # cl::Buffer + enqueueWriteBuffer
    kernel = ctx.builtin("linreg2_p")              # program build + cl::Kernel
    kernel.set_args(None, None, params(0), params(1), x_d, y_d, None, n)   # setArg 0..7
    kernel.configure(global_size=n, max_ops=4, epilogue="half_n_log")
    f, grad = kernel.eval([a, b])                  # the whole of userfunction()

There is NO CPU fallback: importing this module loads libad4cl_cuda.so and
raises if it is missing; creating a Context raises without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AD4CL_CUDA_LIB", os.path.join(_HERE, "libad4cl_cuda.so"))  # override: A/B builds

OK, ERR_CUDA, ERR_INVALID, ERR_TAPE_OVERFLOW, ERR_WINDOW, ERR_COMPILE, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
ACC = {"auto": -1, "thread": 0, "warp": 1, "global": 2}
TAPE = {"auto": -1, "hbm": 0, "shared": 1, "register": 2}
PREFETCH = {"auto": -1, "off": 0, "on": 1}
RECORDER = {"auto": -1, "uniform": 0, "lane": 1}
LOCKSTEP = {"auto": -1, "off": 0, "on": 1}
EPILOGUE = {"sum": 0, "half_n_log": 1}


class Ad4clError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ad4cl_cuda error {code}: {message}")
        self.code = code


class KernelConfig(ctypes.Structure):
    _fields_ = [("global_size", c_longlong * 2), ("local_size", c_int * 2), ("max_ops", c_int),
                ("max_slots", c_int), ("window", c_int), ("far_plane", c_int), ("acc_mode", c_int),
                ("tape_tier", c_int), ("epilogue", c_int), ("epilogue_n", c_double),
                ("blocks_per_sm", c_int), ("reference_quirks", c_int), ("sweep_prefetch", c_int), ("recorder", c_int),
                ("lockstep", c_int)]


class LbfgsOptions(ctypes.Structure):
    _fields_ = [("max_iterations", c_int), ("max_linesearch_iterations", c_int), ("history", c_int),
                ("gradient_tolerance", c_double), ("function_tolerance", c_double)]


class LbfgsResult(ctypes.Structure):
    _fields_ = [("f", c_double), ("gradient_max", c_double), ("converged", c_int), ("iterations", c_int),
                ("linesearch_iteration", c_int), ("evaluations", c_int), ("number_of_parameters", c_int)]


class EvalStats(ctypes.Structure):
    _fields_ = [("ops", c_double), ("slots", c_double), ("grid", c_int), ("block", c_int),
                ("smem_bytes", c_size_t), ("arena_bytes", c_size_t), ("kernel_ms", c_float),
                ("sweep_prefetch", c_int), ("tape_tier", c_int), ("recorder", c_int),
                ("lockstep", c_int)]


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C ad4cl_b200/csrc` (or __graft_entry__.build()). "
            "ad4cl_b200 has no CPU implementation to fall back to.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.ad4cl_cuda_last_error.restype = c_char_p
    lib.ad4cl_cuda_device_abi.restype = c_char_p
    lib.ad4cl_cuda_stream.restype = c_void_p
    lib.ad4cl_cuda_stream.argtypes = [c_void_p]
    lib.ad4cl_cuda_use_stream.argtypes = [c_void_p, c_void_p]
    lib.ad4cl_cuda_buffer_device_ptr.restype = c_void_p
    lib.ad4cl_cuda_buffer_device_ptr.argtypes = [c_void_p]
    lib.ad4cl_cuda_builtin_name.restype = c_char_p
    lib.ad4cl_cuda_builtin_signature.restype = c_char_p
    lib.ad4cl_cuda_kernel_signature.restype = c_char_p
    lib.ad4cl_cuda_kernel_signature.argtypes = [c_void_p]
    lib.ad4cl_cuda_init.argtypes = [c_int, POINTER(c_void_p)]
    lib.ad4cl_cuda_shutdown.argtypes = [c_void_p]
    lib.ad4cl_cuda_device_sm_count.argtypes = [c_void_p]
    lib.ad4cl_cuda_buffer_create.argtypes = [c_void_p, c_size_t, POINTER(c_void_p)]
    lib.ad4cl_cuda_buffer_wrap.argtypes = [c_void_p, c_void_p, c_size_t, POINTER(c_void_p)]
    lib.ad4cl_cuda_buffer_upload.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
    lib.ad4cl_cuda_buffer_download.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
    lib.ad4cl_cuda_buffer_free.argtypes = [c_void_p]
    lib.ad4cl_cuda_kernel_builtin.argtypes = [c_void_p, c_char_p, c_int, POINTER(c_void_p)]
    lib.ad4cl_cuda_kernel_from_source.argtypes = [c_void_p, c_char_p, c_char_p, c_char_p, c_char_p, c_int,
                                                  POINTER(c_void_p)]
    lib.ad4cl_cuda_kernel_free.argtypes = [c_void_p]
    lib.ad4cl_cuda_compile.argtypes = [c_char_p, c_char_p, c_char_p, c_char_p, c_int, POINTER(c_void_p), POINTER(c_size_t)]
    lib.ad4cl_cuda_free_host.argtypes = [c_void_p]
    lib.ad4cl_cuda_free_host.restype = None
    lib.ad4cl_cuda_kernel_from_cubin.argtypes = [c_void_p, c_char_p, c_size_t, c_char_p, c_char_p, c_int, POINTER(c_void_p)]
    lib.ad4cl_cuda_kernel_set_arg_buffer.argtypes = [c_void_p, c_int, c_void_p]
    lib.ad4cl_cuda_kernel_set_arg_int.argtypes = [c_void_p, c_int, c_int]
    lib.ad4cl_cuda_kernel_set_arg_double.argtypes = [c_void_p, c_int, c_double]
    lib.ad4cl_cuda_kernel_set_arg_params.argtypes = [c_void_p, c_int, c_int, c_int]
    lib.ad4cl_cuda_kernel_config_default.argtypes = [POINTER(KernelConfig)]
    lib.ad4cl_cuda_kernel_configure.argtypes = [c_void_p, POINTER(KernelConfig)]
    lib.ad4cl_cuda_eval.argtypes = [c_void_p, c_void_p, c_int, POINTER(c_double), c_void_p]
    lib.ad4cl_cuda_eval_device.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int]
    lib.ad4cl_cuda_check.argtypes = [c_void_p]
    lib.ad4cl_cuda_last_stats.argtypes = [c_void_p, POINTER(EvalStats)]
    lib.ad4cl_cuda_kernel_profile.argtypes = [c_void_p, c_int]
    lib.ad4cl_cuda_export_tape.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_longlong, POINTER(c_longlong), c_void_p, c_longlong]
    lib.ad4cl_cuda_measure_fp64_peak.argtypes = [c_void_p, POINTER(c_double)]
    lib.ad4cl_cuda_lbfgs_options_default.argtypes = [POINTER(LbfgsOptions)]
    lib.ad4cl_cuda_minimize_lbfgs.argtypes = [c_void_p, c_void_p, c_int, POINTER(LbfgsOptions), POINTER(LbfgsResult)]
    lib.ad4cl_cuda_matrix_product_dense.argtypes = [c_void_p, c_int, c_int] + [c_void_p] * 8
    lib.ad4cl_cuda_comm_create.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]
    lib.ad4cl_cuda_comm_handle.argtypes = [c_void_p, c_void_p]
    lib.ad4cl_cuda_comm_connect.argtypes = [c_void_p, c_void_p]
    lib.ad4cl_cuda_comm_allreduce.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double]
    lib.ad4cl_cuda_comm_check.argtypes = [c_void_p]
    lib.ad4cl_cuda_comm_destroy.argtypes = [c_void_p]
    lib.ad4cl_cuda_kernel_profile_read.argtypes = [c_void_p, POINTER(c_double), POINTER(c_longlong), c_int]
    lib.ad4cl_cuda_comm_create_local.argtypes = [POINTER(c_void_p), c_int, c_int, POINTER(c_void_p)]
    lib.ad4cl_cuda_comm_set_timeout.argtypes = [c_void_p, c_double]
    lib.ad4cl_cuda_eval_sharded.argtypes = [c_void_p, c_void_p, c_void_p, c_int, POINTER(c_double), c_void_p]
    lib.ad4cl_cuda_eval_device_sharded.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]
    lib.ad4cl_cuda_launches_per_eval.argtypes = [c_void_p]
    if hasattr(lib, "ad4cl_cuda_eval_general"):
        lib.ad4cl_cuda_eval_general.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, POINTER(c_double), c_void_p]
    if hasattr(lib, "ad4cl_cuda_minimize_lbfgs_device"):
        lib.ad4cl_cuda_minimize_lbfgs_device.argtypes = [c_void_p, c_void_p, c_int, POINTER(LbfgsOptions), POINTER(LbfgsResult)]
        lib.ad4cl_cuda_eval_device_if.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]
    return lib


lib = _load()


def _check(rc: int) -> None:
    if rc != OK:
        raise Ad4clError(rc, (lib.ad4cl_cuda_last_error() or b"").decode(errors="replace"))


def device_abi() -> str:
    """Hash of the device headers the library embeds; cubins are only valid for the same hash."""
    return lib.ad4cl_cuda_device_abi().decode()


def compile_source(source: str, entry: str, signature: str, options: str = "", quirks: bool = False) -> bytes:
    """NVRTC-compile kernel text written against ad.cl to an sm_100a cubin. Needs no GPU."""
    out, size = c_void_p(), c_size_t()
    _check(lib.ad4cl_cuda_compile(source.encode(), entry.encode(), signature.encode(),
                                  options.encode() if options else None, 1 if quirks else 0, byref(out), byref(size)))
    try:
        return ctypes.string_at(out.value, size.value)
    finally:
        lib.ad4cl_cuda_free_host(out)


def builtin_kernels() -> dict[str, str]:
    """name -> signature of the kernels compiled into the library."""
    return {lib.ad4cl_cuda_builtin_name(i).decode(): lib.ad4cl_cuda_builtin_signature(i).decode()
            for i in range(lib.ad4cl_cuda_builtin_count())}


class params:
    """Marks a kernel argument bound to `count` independents starting at id `first_id`
    (what ad_init_var + enqueueWriteBuffer(a_d, ...) set up in simple.cpp:278-286)."""

    def __init__(self, first_id: int, count: int = 1):
        self.first_id, self.count = first_id, count


class Buffer:
    def __init__(self, ctx: "Context", handle: c_void_p, nbytes: int):
        self.ctx, self._h, self.nbytes = ctx, handle, nbytes

    @property
    def device_ptr(self) -> int:
        return lib.ad4cl_cuda_buffer_device_ptr(self._h)

    def upload(self, host: np.ndarray, offset: int = 0) -> None:
        host = np.ascontiguousarray(host)
        _check(lib.ad4cl_cuda_buffer_upload(self._h, offset, host.ctypes.data, host.nbytes))

    def download(self, dtype=np.float64, count: int | None = None, offset: int = 0) -> np.ndarray:
        dt = np.dtype(dtype)
        n = (self.nbytes - offset) // dt.itemsize if count is None else count
        out = np.empty(n, dtype=dt)
        _check(lib.ad4cl_cuda_buffer_download(self._h, offset, out.ctypes.data, out.nbytes))
        return out

    def free(self) -> None:
        if self._h:
            lib.ad4cl_cuda_buffer_free(self._h)
            self._h = None


class Kernel:
    def __init__(self, ctx: "Context", handle: c_void_p, name: str):
        self.ctx, self._h, self.name = ctx, handle, name
        self.signature = lib.ad4cl_cuda_kernel_signature(handle).decode()
        self._keep = []

    def set_arg(self, index: int, value) -> None:
        """cl::Kernel::setArg: Buffer, params(...), int, float; None for gs / stack / out."""
        kind = self.signature[index] if 0 <= index < len(self.signature) else "?"
        if value is None:
            if kind in "GSO":
                return
            if kind in "DI":
                _check(lib.ad4cl_cuda_kernel_set_arg_buffer(self._h, index, None))
                return
            raise Ad4clError(ERR_INVALID, f"argument {index} of {self.name} (kind {kind}) cannot be None")
        if isinstance(value, params):
            _check(lib.ad4cl_cuda_kernel_set_arg_params(self._h, index, value.first_id, value.count))
        elif isinstance(value, Buffer):
            self._keep.append(value)
            _check(lib.ad4cl_cuda_kernel_set_arg_buffer(self._h, index, value._h))
        elif isinstance(value, (int, np.integer)):
            _check(lib.ad4cl_cuda_kernel_set_arg_int(self._h, index, int(value)))
        elif isinstance(value, (float, np.floating)):
            _check(lib.ad4cl_cuda_kernel_set_arg_double(self._h, index, float(value)))
        else:
            raise TypeError(f"unsupported kernel argument {value!r}")

    def set_args(self, *values) -> "Kernel":
        if len(values) != len(self.signature):
            raise Ad4clError(ERR_INVALID, f"{self.name} takes {len(self.signature)} arguments ({self.signature}), got {len(values)}")
        for i, v in enumerate(values):
            self.set_arg(i, v)
        return self

    def configure(self, global_size, max_ops: int, *, local_size=None, max_slots: int = 0, window: int = 0,
                  far_plane: bool = True, acc: str = "auto", tape: str = "auto", epilogue: str = "sum",
                  epilogue_n: float = 0.0, blocks_per_sm: int = 0, prefetch: str = "auto",
                  recorder: str = "auto", lockstep: str = "auto") -> "Kernel":
        cfg = KernelConfig()
        _check(lib.ad4cl_cuda_kernel_config_default(byref(cfg)))
        gs = (global_size, 1) if np.isscalar(global_size) else tuple(global_size)
        cfg.global_size[0], cfg.global_size[1] = int(gs[0]), int(gs[1]) if len(gs) > 1 else 1
        if local_size is not None:
            ls = (local_size, 1) if np.isscalar(local_size) else tuple(local_size)
            cfg.local_size[0], cfg.local_size[1] = int(ls[0]), int(ls[1]) if len(ls) > 1 else 1
        cfg.max_ops, cfg.max_slots, cfg.window = int(max_ops), int(max_slots), int(window)
        cfg.far_plane = 1 if far_plane else 0
        cfg.acc_mode, cfg.tape_tier, cfg.epilogue = ACC[acc], TAPE[tape], EPILOGUE[epilogue]
        cfg.epilogue_n = float(epilogue_n)
        cfg.blocks_per_sm = int(blocks_per_sm)
        cfg.sweep_prefetch = PREFETCH[prefetch]
        cfg.recorder = RECORDER[recorder]
        cfg.lockstep = LOCKSTEP[lockstep]
        _check(lib.ad4cl_cuda_kernel_configure(self._h, byref(cfg)))
        return self

    def eval(self, theta, want_grad: bool = True):
        """theta (host) -> (f, grad[P]) on the host; grad is None when want_grad is False."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        f = c_double()
        g = np.empty(len(th)) if want_grad else None
        _check(lib.ad4cl_cuda_eval(self._h, th.ctypes.data, len(th), byref(f), g.ctypes.data if want_grad else None))
        return f.value, g

    def eval_device(self, theta_dev_ptr: int, nparams: int, result_dev_ptr: int, *, want_grad: bool = True,
                    apply_epilogue: bool = True) -> None:
        """Asynchronous evaluation on the context's stream with device-resident theta and
        result ([grad..., f, ops, slots], nparams + 3 doubles)."""
        _check(lib.ad4cl_cuda_eval_device(self._h, theta_dev_ptr, nparams, result_dev_ptr,
                                          1 if want_grad else 0, 1 if apply_epilogue else 0))

    def eval_sharded(self, comm: "Comm", theta, want_grad: bool = True):
        """One sharded evaluation in one call (ad4cl_cuda_eval_sharded): host theta up, ONE kernel whose last
        CTA all-reduces the raw sums with the peers of `comm` and applies the epilogue, (f, grad) down."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        f = c_double()
        g = np.empty(len(th)) if want_grad else None
        _check(lib.ad4cl_cuda_eval_sharded(self._h, comm._h, th.ctypes.data, len(th), byref(f),
                                           g.ctypes.data if want_grad else None))
        return f.value, g

    def eval_device_sharded(self, comm: "Comm", theta_dev_ptr: int, nparams: int, result_dev_ptr: int, *,
                            want_grad: bool = True) -> None:
        """Device-resident, asynchronous form of eval_sharded: result = [grad..., f, ops, slots] summed over ranks."""
        _check(lib.ad4cl_cuda_eval_device_sharded(self._h, comm._h, theta_dev_ptr, nparams, result_dev_ptr,
                                                  1 if want_grad else 0))

    def eval_outputs(self, theta, n: int):
        """Pass 1 of a general epilogue: -> (sum_i out_i, Buffer holding out_i.value for the n work-items)."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        out = self.ctx.buffer(nbytes=8 * n)
        s = c_double()
        _check(lib.ad4cl_cuda_eval_general(self._h, th.ctypes.data, len(th), out.device_ptr, None, byref(s), None))
        return s.value, out

    def eval_seeded(self, theta, seeds: "Buffer"):
        """Pass 2: the gradient of sum_i seeds_i * out_i(theta) (seeds: device Buffer of n doubles) -> grad[P]."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        g = np.empty(len(th))
        s = c_double()
        _check(lib.ad4cl_cuda_eval_general(self._h, th.ctypes.data, len(th), None, seeds.device_ptr, byref(s), g.ctypes.data))
        return g

    @property
    def launches_per_eval(self) -> int:
        """Kernel launches of the last evaluation enqueued (1 unless acc = global)."""
        return lib.ad4cl_cuda_launches_per_eval(self._h)

    def check(self) -> None:
        _check(lib.ad4cl_cuda_check(self._h))

    def export_tape(self, theta, n_out: int, capacity: int, first_free_id: int | None = None):
        """Record without sweeping and return (entries, out): numpy structured arrays in the
        reference's layouts (40-byte ad_entry, 16-byte ad_variable)."""
        entry_t = np.dtype([("dx0", "<f8"), ("id0", "<i4"), ("_p0", "<i4"), ("dx1", "<f8"), ("id1", "<i4"),
                            ("_p1", "<i4"), ("id", "<i4"), ("size", "<i4")])
        var_t = np.dtype([("value", "<f8"), ("id", "<i4"), ("_p", "<i4")])
        th = np.ascontiguousarray(theta, dtype=np.float64)
        entries = np.zeros(capacity, dtype=entry_t)
        out = np.zeros(n_out, dtype=var_t)
        n = c_longlong()
        ffi = len(th) if first_free_id is None else first_free_id
        _check(lib.ad4cl_cuda_export_tape(self._h, th.ctypes.data, len(th), ffi, entries.ctypes.data, capacity,
                                          byref(n), out.ctypes.data, n_out))
        return entries[: n.value], out

    def minimize(self, theta0, device_resident: bool = True, **options):
        """L-BFGS inside the library -> (theta, dict of the result fields).  device_resident (default): the
        whole fit runs on the GPU as CUDA graphs of (evaluation, lbfgs_update) pairs, no host round trip
        between evaluations (ad4cl_cuda_minimize_lbfgs_device); False: the host loop over ad4cl_cuda_eval."""
        th = np.array(theta0, dtype=np.float64, copy=True)
        opt, res = LbfgsOptions(), LbfgsResult()
        _check(lib.ad4cl_cuda_lbfgs_options_default(byref(opt)))
        for key, v in options.items():
            setattr(opt, key, v)
        fn = lib.ad4cl_cuda_minimize_lbfgs_device if device_resident else lib.ad4cl_cuda_minimize_lbfgs
        _check(fn(self._h, th.ctypes.data, len(th), byref(opt), byref(res)))
        return th, {name: getattr(res, name) for name, _ in LbfgsResult._fields_}

    def stats(self) -> dict:
        s = EvalStats()
        _check(lib.ad4cl_cuda_last_stats(self._h, byref(s)))
        return {"ops": s.ops, "slots": s.slots, "grid": s.grid, "block": s.block, "smem_bytes": s.smem_bytes,
                "arena_bytes": s.arena_bytes, "kernel_ms": s.kernel_ms, "sweep_prefetch": s.sweep_prefetch,
                "tape_tier": {v: k for k, v in TAPE.items()}.get(s.tape_tier, str(s.tape_tier)),
                "recorder": {v: k for k, v in RECORDER.items()}.get(s.recorder, str(s.recorder)) if s.recorder >= 0 else "n/a",
                "lockstep": int(s.lockstep)}

    def profile(self, enable: bool = True) -> None:
        """CUDA-event timing of the fused kernel alone (CL_PROFILING analogue)."""
        _check(lib.ad4cl_cuda_kernel_profile(self._h, 1 if enable else 0))

    def profile_read(self, reset: bool = True) -> tuple[float, int]:
        """-> (mean ms per launch of the fused kernel, launches seen)."""
        ms, n = c_double(), c_longlong()
        _check(lib.ad4cl_cuda_kernel_profile_read(self._h, byref(ms), byref(n), 1 if reset else 0))
        return ms.value, n.value

    def free(self) -> None:
        """Release the kernel and the data buffers workloads.bind() uploaded for it."""
        if self._h:
            lib.ad4cl_cuda_kernel_free(self._h)
            self._h = None
        for b in getattr(self, "_buffers", ()) or ():
            if b is not None:
                b.free()
        self._buffers = ()


class Comm:
    """One-shot NVLink all-reduce of the short result vector (see include/ad4cl_cuda.h)."""

    def __init__(self, ctx: "Context", rank: int, world: int, max_doubles: int, _handle=None):
        h = c_void_p()
        if _handle is None:
            _check(lib.ad4cl_cuda_comm_create(ctx._h, rank, world, max_doubles, byref(h)))
        else:
            h = _handle
        self._h, self.rank, self.world = h, rank, world

    @staticmethod
    def create_local(ctxs: list["Context"], max_doubles: int) -> list["Comm"]:
        """All ranks in ONE process (one Context per GPU), connected through plain peer access."""
        n = len(ctxs)
        arr = (c_void_p * n)(*[c._h for c in ctxs])
        out = (c_void_p * n)()
        _check(lib.ad4cl_cuda_comm_create_local(arr, n, max_doubles, out))
        return [Comm(ctxs[r], r, n, max_doubles, _handle=c_void_p(out[r])) for r in range(n)]

    def set_timeout(self, seconds: float) -> None:
        _check(lib.ad4cl_cuda_comm_set_timeout(self._h, float(seconds)))

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(lib.ad4cl_cuda_comm_handle_bytes())
        _check(lib.ad4cl_cuda_comm_handle(self._h, buf))
        return buf.raw

    def connect(self, handles: list[bytes]) -> None:
        blob = b"".join(handles)
        _check(lib.ad4cl_cuda_comm_connect(self._h, blob))

    def allreduce(self, data_dev_ptr: int, n: int, nparams: int, epilogue: str = "sum", n_obs: float = 0.0) -> None:
        _check(lib.ad4cl_cuda_comm_allreduce(self._h, data_dev_ptr, n, nparams, EPILOGUE[epilogue], float(n_obs)))

    def check(self) -> None:
        _check(lib.ad4cl_cuda_comm_check(self._h))

    def destroy(self) -> None:
        if self._h:
            lib.ad4cl_cuda_comm_destroy(self._h)
            self._h = None


class Context:
    def __init__(self, device: int = 0):
        h = c_void_p()
        _check(lib.ad4cl_cuda_init(device, byref(h)))
        self._h, self.device = h, device

    @property
    def sm_count(self) -> int:
        return lib.ad4cl_cuda_device_sm_count(self._h)

    @property
    def stream(self) -> int:
        return lib.ad4cl_cuda_stream(self._h)

    def use_stream(self, cuda_stream: int | None) -> None:
        _check(lib.ad4cl_cuda_use_stream(self._h, cuda_stream))

    def buffer(self, data=None, nbytes: int | None = None) -> Buffer:
        """cl::Buffer (+ blocking enqueueWriteBuffer when `data` is given)."""
        if data is not None:
            data = np.ascontiguousarray(data)
            nbytes = data.nbytes
        h = c_void_p()
        _check(lib.ad4cl_cuda_buffer_create(self._h, nbytes, byref(h)))
        b = Buffer(self, h, nbytes)
        if data is not None and nbytes:
            b.upload(data)
        return b

    def wrap(self, device_ptr: int, nbytes: int) -> Buffer:
        h = c_void_p()
        _check(lib.ad4cl_cuda_buffer_wrap(self._h, device_ptr, nbytes, byref(h)))
        return Buffer(self, h, nbytes)

    def builtin(self, name: str, quirks: bool = False) -> Kernel:
        h = c_void_p()
        _check(lib.ad4cl_cuda_kernel_builtin(self._h, name.encode(), 1 if quirks else 0, byref(h)))
        return Kernel(self, h, name)

    def from_source(self, source: str, entry: str, signature: str, options: str = "", quirks: bool = False) -> Kernel:
        """JIT path (NVRTC), the analogue of cl::Program(context, ad.cl + kernel text).build()."""
        h = c_void_p()
        _check(lib.ad4cl_cuda_kernel_from_source(self._h, source.encode(), entry.encode(), signature.encode(),
                                                 options.encode() if options else None, 1 if quirks else 0, byref(h)))
        return Kernel(self, h, entry)

    def from_cubin(self, cubin: bytes, entry: str, signature: str, quirks: bool = False) -> Kernel:
        h = c_void_p()
        _check(lib.ad4cl_cuda_kernel_from_cubin(self._h, cubin, len(cubin), entry.encode(), signature.encode(),
                                                1 if quirks else 0, byref(h)))
        return Kernel(self, h, entry)

    def matrix_product_dense(self, n: int, A: Buffer, B: Buffer, *, G: Buffer | None = None, C: Buffer | None = None,
                             dA: Buffer | None = None, dB: Buffer | None = None, total: Buffer | None = None,
                             workspace: Buffer | None = None, variant: int = 2) -> None:
        """config 2 without a tape: C = A B, sum C, dL/dA = G B^T, dL/dB = A^T G (asynchronous).
        variant 2 (default) = fp64 tensor-core DMMA, all products in ONE persistent launch; 1 = DMMA, one launch per
        product; 0 = CUDA-core DFMA (2.3x slower than 1)."""
        p = lambda b: None if b is None else b.device_ptr  # noqa: E731
        _check(lib.ad4cl_cuda_matrix_product_dense(self._h, variant, n, p(A), p(B), p(G), p(C), p(dA), p(dB),
                                                   p(total), p(workspace)))

    def measure_fp64_peak(self) -> float:
        """Measured fp64 FMA rate in TFLOP/s (microbenchmark, ~20 ms)."""
        v = c_double()
        _check(lib.ad4cl_cuda_measure_fp64_peak(self._h, byref(v)))
        return v.value

    def synchronize(self) -> None:
        """Wait for everything queued on the context's stream (a 0-byte blocking download)."""
        if not hasattr(self, "_sync_buf"):
            self._sync_buf = self.buffer(nbytes=8)
        self._sync_buf.download(np.uint8, 0)

    def close(self) -> None:
        if self._h:
            lib.ad4cl_cuda_shutdown(self._h)
            self._h = None
