/*
 * builtin_kernels.cu -- ahead-of-time (nvcc, sm_100a) instances of the
 * objective kernels of ../kernels/objectives.cl behind the fused evaluation
 * entry of ad4cl_entry.cuh.  Compiled twice by the Makefile:
 *   AD4CL_REFERENCE_QUIRKS=0, AD4CL_ENTRY_SUFFIX=_fixed,  AD4CL_BUILTIN_TABLE=ad4cl_builtin_table_fixed
 *   AD4CL_REFERENCE_QUIRKS=1, AD4CL_ENTRY_SUFFIX=_quirks, AD4CL_BUILTIN_TABLE=ad4cl_builtin_table_quirks
 * Each build exports one table {name, entry, signature} that shim.cu searches.
 *
 * Signature letters describe the USER kernel's parameter list, in order (it is
 * what ad4cl_cuda_kernel_set_arg_* indexes, like cl::Kernel::setArg):
 *   G gs   S gradient_stack   O out        -- provided by the runtime
 *   V ad_variable* (independents)   D double*   i int
 */
#include "ad4cl_opencl_compat.cuh"
#include "ad4cl_builtin.h"
#include "ad4cl_var.cuh"

#include "../kernels/objectives.cl"
#include "../kernels/matrix.cl"
#include "../kernels/op_probe.cl"
#include "../kernels/var_kernels.cl"

#define OBJECTIVE_ABI(NAME)                                                                  \
    AD4CL_OBJECTIVE_ENTRY(NAME,                                                               \
        (struct ad_variable* theta, double* d0, double* d1, int size, int i0, int i1),        \
        (gs, gradient_stack, theta, d0, d1, out, size, i0, i1))

AD4CL_OBJECTIVE_ENTRY(linreg2_g,
    (struct ad_variable* a, struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, a, b, x, y, out, size))
AD4CL_OBJECTIVE_ENTRY(linreg2_p,
    (const struct ad_variable* a, const struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, a, b, x, y, out, size))
AD4CL_OBJECTIVE_ENTRY(linreg2_private,
    (const struct ad_variable* a, const struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, a, b, x, y, out, size))
AD4CL_OBJECTIVE_ENTRY(linreg2_var,
    (struct ad_variable* a, struct ad_variable* b, double* x, double* y, int size),
    (gs, gradient_stack, a, b, x, y, out, size))
OBJECTIVE_ABI(regression_logistic_var)
OBJECTIVE_ABI(regression_linear)
OBJECTIVE_ABI(regression_logistic)
OBJECTIVE_ABI(catch_at_age)
OBJECTIVE_ABI(tape_stress)
OBJECTIVE_ABI(op_zoo)
OBJECTIVE_ABI(op_zoo_pad)
OBJECTIVE_ABI(op_zoo_private)
OBJECTIVE_ABI(op_probe)
AD4CL_OBJECTIVE_ENTRY(matrix_product,
    (struct ad_variable* A, struct ad_variable* B, int widthA, int widthB),
    (gs, gradient_stack, A, B, out, widthA, widthB))
AD4CL_OBJECTIVE_ENTRY(matrix_elementwise,
    (struct ad_variable* A, struct ad_variable* B, int width, int op),
    (gs, gradient_stack, A, B, out, width, op))

#define ROW(NAME, SIG) {#NAME, {{(const void*) &AD4CL_ENTRY_NAME2(NAME, u, 0), (const void*) &AD4CL_ENTRY_NAME2(NAME, u, 1), \
                              (const void*) &AD4CL_ENTRY_NAME2(NAME, u, 2)},                                       \
                             {(const void*) &AD4CL_ENTRY_NAME2(NAME, m, 0), (const void*) &AD4CL_ENTRY_NAME2(NAME, m, 1), \
                              (const void*) &AD4CL_ENTRY_NAME2(NAME, m, 2)}}, SIG}
extern "C" const ad4cl_builtin_row AD4CL_BUILTIN_TABLE[] = {
    ROW(linreg2_g, "GSVVDDOi"),
    ROW(linreg2_p, "GSVVDDOi"),
    ROW(linreg2_private, "GSVVDDOi"),
    ROW(linreg2_var, "GSVVDDOi"),
    ROW(regression_logistic_var, "GSVDDOiii"),
    ROW(regression_linear, "GSVDDOiii"),
    ROW(regression_logistic, "GSVDDOiii"),
    ROW(catch_at_age, "GSVDDOiii"),
    ROW(tape_stress, "GSVDDOiii"),
    ROW(op_zoo, "GSVDDOiii"),
    ROW(op_zoo_pad, "GSVDDOiii"),
    ROW(op_zoo_private, "GSVDDOiii"),
    ROW(op_probe, "GSVDDOiii"),
    ROW(matrix_product, "GSVVOii"),
    ROW(matrix_elementwise, "GSVVOii"),
    {nullptr, {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}, nullptr},
};
