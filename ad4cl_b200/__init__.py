"""ad4cl_b200 -- B200-native (sm_100a CUDA) implementation of AD4CL's data-parallel hot path.

    host       ctypes mirror of the C ABI (include/ad4cl_cuda.h): Context, Buffer, Kernel, Comm
    workloads  the BASELINE.json configurations as (data, kernel arguments, tape budget) recipes
    dist       multi-GPU evaluation: observation shards + one all-reduce of P+3 doubles
    synth      frozen, counter-based synthetic inputs

Importing `host` loads ad4cl_b200/libad4cl_cuda.so and raises if it is missing: there is no CPU
implementation to fall back to.  (`workloads`, `synth` and `dist` import without it.)
"""
__all__ = ["host", "workloads", "dist", "synth"]
