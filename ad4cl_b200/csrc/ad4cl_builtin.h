/* ad4cl_builtin.h -- row type of the ahead-of-time kernel tables (builtin_kernels.cu). */
#pragma once
struct ad4cl_builtin_row {
    const char* name;        /* user kernel name */
    const void* entry[3];    /* host handles of NAME__entry_m0/_m1/_m2 (accumulator mode) for cudaLaunchKernel */
    const char* signature;   /* user kernel parameter kinds, see builtin_kernels.cu */
};
