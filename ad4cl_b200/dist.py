"""Multi-GPU evaluation: observations are sharded contiguously over the ranks
(one process per GPU), every rank produces the raw sums of its shard
[d S/d theta_0..P-1, S, #ops, #slots] with the single-GPU path, ONE all-reduce
of those P+3 doubles combines them, and the scalar epilogue f = h(S),
grad = h'(S) dS/dtheta is applied identically on every rank.

The objective of every AD4CL example has this form: f = h(sum_i out_i), with
h = identity or (n/2) log (test/admb/simple.cpp:357-365).  Nothing else crosses
the interconnect; there is no data-path collective.

`torch.distributed` is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.distributed as dist

from .workloads import shard_range  # noqa: F401  (re-exported: the sharding rule)


def apply_epilogue(raw: torch.Tensor, nparams: int, epilogue: str, n_total: float) -> torch.Tensor:
    """raw = [g_0..g_{P-1}, S, ops, slots] (sums over all shards) -> same layout with
    f in place of S and the gradient scaled.  (n/2) log S: simple.cpp:365, main.cpp:417."""
    if epilogue == "sum":
        return raw
    if epilogue == "half_n_log":
        out = raw.clone()
        S = raw[nparams]
        out[:nparams] = raw[:nparams] * ((n_total / 2.0) * (1.0 / S))
        out[nparams] = (n_total / 2.0) * torch.log(S)
        return out
    raise ValueError(epilogue)


class ShardedObjective:
    """local_eval(theta, raw_out) must fill raw_out (P+3 doubles, on `device`) with this
    rank's raw sums for parameters `theta` (P doubles on `device`), asynchronously on
    the current stream.

    `comm` (a connected host.Comm) switches the reduction from torch.distributed's
    all_reduce (NCCL / gloo) to the library's one-shot NVLink all-reduce, which also applies
    the epilogue on the device."""

    def __init__(self, local_eval: Callable[[torch.Tensor, torch.Tensor], None], nparams: int,
                 epilogue: str, n_total: int, device: torch.device, group=None, comm=None, kernel=None):
        self.local_eval, self.P, self.epilogue, self.n_total = local_eval, nparams, epilogue, n_total
        self.device, self.group, self.comm, self.kernel = device, group, comm, kernel
        self.raw = torch.zeros(nparams + 3, dtype=torch.float64, device=device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def eval_raw(self, theta: torch.Tensor) -> torch.Tensor:
        """-> all-reduced raw sums (device tensor; no host synchronisation)."""
        self.local_eval(theta, self.raw)
        if self.world > 1:
            dist.all_reduce(self.raw, op=dist.ReduceOp.SUM, group=self.group)
        return self.raw

    def eval_device(self, theta: torch.Tensor) -> torch.Tensor:
        """-> [grad..., f, ops, slots] on the device, epilogue applied, no host synchronisation."""
        if self.comm is not None and self.world > 1:
            if self.kernel is not None:
                # ONE launch: the evaluation kernel's last CTA exchanges the raw sums with the peers over
                # NVLink and applies the epilogue (ad4cl_cuda_eval_device_sharded)
                self.kernel.eval_device_sharded(self.comm, theta.data_ptr(), self.P, self.raw.data_ptr())
                return self.raw
            self.local_eval(theta, self.raw)
            self.comm.allreduce(self.raw.data_ptr(), self.P + 3, self.P, self.epilogue, float(self.n_total))
            return self.raw
        return apply_epilogue(self.eval_raw(theta), self.P, self.epilogue, float(self.n_total))

    def eval(self, theta: torch.Tensor):
        """-> (f, grad) as a python float and a host tensor.  A peer that never arrived (one-shot
        all-reduce time-out) raises instead of returning the NaN-filled vector."""
        out = self.eval_device(theta).cpu()
        if self.comm is not None:
            self.comm.check()
        return float(out[self.P]), out[: self.P].clone()


def connect_oneshot(ctx, nparams: int, group=None):
    """Create the one-shot all-reduce communicator of this rank and connect it to its peers:
    the IPC handles are exchanged with torch.distributed (plumbing), the data path is the
    library's own kernel over NVLink peer memory.  Returns None when world size is 1."""
    from . import host
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    comm = host.Comm(ctx, rank, world, nparams + 3)
    handles = [None] * world
    dist.all_gather_object(handles, comm.handle(), group=group)
    comm.connect(handles)
    dist.barrier(group)
    return comm


def from_kernel(kernel, workload, device: torch.device, group=None, comm=None) -> ShardedObjective:
    """Wrap a configured host.Kernel bound to this rank's shard of `workload`."""
    P = workload.nparams

    def local_eval(theta: torch.Tensor, raw: torch.Tensor) -> None:
        kernel.eval_device(theta.data_ptr(), P, raw.data_ptr(), want_grad=True, apply_epilogue=False)

    n_total = int(workload.epilogue_n) if workload.epilogue_n else workload.size
    if device.type == "cuda":
        # the library launches on the context's stream: make it the stream torch is using on this device,
        # so the kernel, the collective and the caller's copies of theta / the result are ordered
        kernel.ctx.use_stream(torch.cuda.current_stream(device).cuda_stream)
    return ShardedObjective(local_eval, P, workload.epilogue, n_total, device, group, comm,
                            kernel=kernel if comm is not None else None)
