/*
 * builtin_static.cu -- REGISTER-tier instances (ad4cl_tape.cuh, AD4CL_TAPE_STATIC) of the short
 * straight-line objectives of ../kernels/objectives.cl: the 4-operator kernel of the shipped example
 * (kernel.cl:17-25, test/admb/simple.cl:123-137; P = 2) and the regressions of BASELINE config 3
 * (P = 10).  Compiled once per (AD4CL_P, AD4CL_REFERENCE_QUIRKS) by the Makefile; each build exports
 * one table that shim.cu consults when a kernel is configured with tape tier AUTO or REGISTER.
 *
 * The entries take the SAME parameter list as the dynamic ones (the shim marshals both alike) but
 * hand the user kernel `th`, local copies of the independents with literal ids, and literals for the
 * integer arguments that bound its loops -- the row records which argument was frozen to what.
 */
#include "ad4cl_opencl_compat.cuh"
#include "ad4cl_builtin.h"

#include "../kernels/objectives.cl"

#define SROW(NAME, SIG, FROZEN_ARG, FROZEN_VALUE, SLOTS) \
    {#NAME, (const void*) &AD4CL_STATIC_ENTRY_NAME(NAME), SIG, AD4CL_P, FROZEN_ARG, FROZEN_VALUE, SLOTS}

#if AD4CL_P == 2
AD4CL_STATIC_ENTRY(linreg2_g,
    (struct ad_variable* a, struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, th + 0, th + 1, x, y, out, size))
AD4CL_STATIC_ENTRY(linreg2_p,
    (const struct ad_variable* a, const struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, th + 0, th + 1, x, y, out, size))
extern "C" const ad4cl_static_row AD4CL_STATIC_TABLE[] = {
    SROW(linreg2_g, "GSVVDDOi", -1, 0, 6),
    SROW(linreg2_p, "GSVVDDOi", -1, 0, 6),
    {nullptr, nullptr, nullptr, 0, -1, 0, 0},
};
#elif AD4CL_P == 10
/* i0 (argument 7 of "GSVDDOiii") is the number of parameters = the trip count of the kernel's loop */
AD4CL_STATIC_ENTRY(regression_linear,
    (struct ad_variable* theta, double* d0, double* d1, int size, int i0, int i1),
    (gs, gradient_stack, th, d0, d1, out, size, AD4CL_P, i1))
AD4CL_STATIC_ENTRY(regression_logistic,
    (struct ad_variable* theta, double* d0, double* d1, int size, int i0, int i1),
    (gs, gradient_stack, th, d0, d1, out, size, AD4CL_P, i1))
extern "C" const ad4cl_static_row AD4CL_STATIC_TABLE[] = {
    SROW(regression_linear, "GSVDDOiii", 7, AD4CL_P, 3 * AD4CL_P + 1),
    SROW(regression_logistic, "GSVDDOiii", 7, AD4CL_P, 3 * AD4CL_P + 4),
    {nullptr, nullptr, nullptr, 0, -1, 0, 0},
};
#else
#error "no register-tier instances for this AD4CL_P"
#endif
