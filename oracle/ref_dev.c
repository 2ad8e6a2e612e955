/*
 * oracle/ref_dev.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the REFERENCE'S OWN device code, where it lies under
 * /root/reference, as C99 (through clshim.h) and exports plain-C entry points
 * that run its kernels one work-item at a time.  Nothing from the reference
 * is copied into this repository: the sources are #included by absolute path
 * at build time (see oracle/Makefile) and the objects land in oracle/_ref/.
 *
 * Build-time macros (set by oracle/Makefile):
 *   REF_AD_CL        path of the device runtime   (/root/reference/ad.cl, or
 *                    the quirk-patched copy oracle/_ref/ad_fixed.cl)
 *   REF_SIMPLE_CL    /root/reference/test/admb/simple.cl
 *   REF_KERNEL_CL    oracle/_ref/kernel_renamed.cl (kernel.cl after the Q10
 *                    type rename of SURVEY.md 2.3, done by sed in the Makefile)
 *   REF_MATRIX_CL    /root/reference/test/matrix/matrixmul.cl   (GLOBAL variant)
 *   REF_MATRIX_LOCAL_CL  oracle/_ref/matrixmul_local.cl (same file with the
 *                    leading "#define GLOBAL" turned into USE_LOCAL by sed)
 *   OBJECTIVES_CL    ad4cl_b200/kernels/objectives.cl (this repo's kernels)
 *   SYM(x)           symbol prefix: ref_##x or reffix_##x
 */
#include "clshim.h"
#include <string.h>
#include <stdlib.h>

#include REF_AD_CL

/* --- the reference's example kernels, renamed so they can share one TU --- */
#define AD AD_simple_cl
#include REF_SIMPLE_CL
#undef AD

#define AD AD_kernel_cl
#define compute_gradient compute_gradient_unused_dup
#include REF_KERNEL_CL
#undef AD
#undef compute_gradient

#define matrixMult matrixMult_global
#include REF_MATRIX_CL
#undef matrixMult
#undef GLOBAL

#define matrixMult matrixMult_local
#include REF_MATRIX_LOCAL_CL
#undef matrixMult

/* --- this repo's objective kernels against the reference runtime --- */
#include OBJECTIVES_CL

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct ad_gradient_structure dev_gs_t; /* 32 bytes (ad.cl:81-89) */

int SYM(sizeof_dev_gs)(void) { return (int) sizeof (dev_gs_t); }
int SYM(sizeof_entry)(void) { return (int) sizeof (struct ad_entry); }
int SYM(sizeof_variable)(void) { return (int) sizeof (struct ad_variable); }
int SYM(sizeof_private_gs)(void) { return (int) sizeof (struct ad_private_gradient_structure); }

/*
 * NDRange drivers.  `gs` is the device-layout structure; the host side
 * (ref_host.cpp) copies its 24-byte prefix in and out exactly as
 * test/admb/simple.cpp:284,312 does with enqueueWrite/ReadBuffer.
 * threads <= 1 runs the work-items serially in id order (deterministic tape).
 */
#define NDRANGE_1D(global_size, threads, CALL)                                   \
    do {                                                                          \
        if ((threads) <= 1) {                                                     \
            clshim_gsize[0] = (global_size);                                      \
            for (int wi_ = 0; wi_ < (global_size); wi_++) {                       \
                clshim_gid[0] = wi_;                                              \
                CALL;                                                             \
            }                                                                     \
        } else {                                                                  \
            _Pragma("omp parallel for schedule(static) num_threads(threads)")     \
            for (int wi_ = 0; wi_ < (global_size); wi_++) {                       \
                clshim_gsize[0] = (global_size);                                  \
                clshim_gid[0] = wi_;                                              \
                CALL;                                                             \
            }                                                                     \
        }                                                                         \
    } while (0)

/* variant 0: kernel.cl (global-stack family); 1: simple.cl (pad family) */
void SYM(launch_simple)(int variant, dev_gs_t* gs, struct ad_entry* stack,
        struct ad_variable* a, struct ad_variable* b, double* x, double* y,
        struct ad_variable* out, int size, int global_size, int threads) {
    if (variant == 0) {
        NDRANGE_1D(global_size, threads, AD_kernel_cl(gs, stack, a, b, x, y, out, size));
    } else {
        NDRANGE_1D(global_size, threads, AD_simple_cl(gs, stack, a, b, x, y, out, size));
    }
}

/* variant 0: GLOBAL (matrixmul.cl:5-28); 1: USE_LOCAL (matrixmul.cl:31-61).
   2-D NDRange (widthA, heightB), id(0) fastest, as matrix_mul.cpp:219 enqueues it. */
void SYM(launch_matrix)(int variant, dev_gs_t* gs, struct ad_entry* stack,
        struct ad_variable* A, struct ad_variable* B, struct ad_variable* C,
        int widthA, int widthB, int g0, int g1) {
    clshim_gsize[0] = g0;
    clshim_gsize[1] = g1;
    for (int j = 0; j < g1; j++) {
        for (int i = 0; i < g0; i++) {
            clshim_gid[0] = i;
            clshim_gid[1] = j;
            if (variant == 0) {
                matrixMult_global(gs, stack, A, B, C, widthA, widthB);
            } else {
                matrixMult_local(gs, stack, A, B, C, widthA, widthB);
            }
        }
    }
}

/* this repo's kernels: by name */
int SYM(launch_objective)(const char* name, dev_gs_t* gs, struct ad_entry* stack,
        struct ad_variable* theta, double* d0, double* d1,
        struct ad_variable* out, int size, int i0, int i1, int global_size, int threads) {
    if (strcmp(name, "linreg2_g") == 0) {
        NDRANGE_1D(global_size, threads, linreg2_g(gs, stack, &theta[0], &theta[1], d0, d1, out, size));
    } else if (strcmp(name, "linreg2_p") == 0) {
        NDRANGE_1D(global_size, threads, linreg2_p(gs, stack, &theta[0], &theta[1], d0, d1, out, size));
    } else if (strcmp(name, "regression_linear") == 0) {
        NDRANGE_1D(global_size, threads, regression_linear(gs, stack, theta, d0, d1, out, size, i0, i1));
    } else if (strcmp(name, "regression_logistic") == 0) {
        NDRANGE_1D(global_size, threads, regression_logistic(gs, stack, theta, d0, d1, out, size, i0, i1));
    } else if (strcmp(name, "catch_at_age") == 0) {
        NDRANGE_1D(global_size, threads, catch_at_age(gs, stack, theta, d0, d1, out, size, i0, i1));
    } else if (strcmp(name, "tape_stress") == 0) {
        NDRANGE_1D(global_size, threads, tape_stress(gs, stack, theta, d0, d1, out, size, i0, i1));
    } else if (strcmp(name, "op_zoo") == 0) {
        NDRANGE_1D(global_size, threads, op_zoo(gs, stack, theta, d0, d1, out, size, i0, i1));
    } else {
        return -1;
    }
    return 0;
}

/*
 * Single-op probes for the op-level known-answer tests: run ONE operator of
 * ONE family on (a, b) with operand ids (ida, idb) and hand back the value,
 * the result id and the tape entry it recorded.
 *   family 0: global ad_*   1: block-reserved pad_*   2: private *_p
 * Returns 0, or -1 if that family has no such operator.
 */
#define OPS_BIN_VV(X)  X(plus) X(minus) X(times) X(divide)
#define OPS_MATH1(X)   X(cos) X(sin) X(tan) X(acos) X(asin) X(atan) X(cosh) X(sinh) \
                       X(tanh) X(exp) X(log) X(log10) X(sqrt)

int SYM(op_probe)(int family, const char* op, double a, double b, int ida, int idb,
        int recording, double* value, int* id, struct ad_entry* entry, double* lhs_after) {
    struct ad_entry* stack = (struct ad_entry*) calloc(8, sizeof (struct ad_entry));
    dev_gs_t g;
    memset(&g, 0, sizeof g);
    g.gradient_stack = stack;
    g.current_ad_variable_id = 100;
    g.stack_current = 2;
    g.recording = recording;
    g.counter = 1;
    dev_gs_t p; /* private structure of the pad family */
    memset(&p, 0, sizeof p);
    if (family == 1) pad_init(1, &p, &g, stack);
    struct ad_private_gradient_structure* q =
            (struct ad_private_gradient_structure*) calloc(1, sizeof (struct ad_private_gradient_structure));
    ad_init_p(q);
    q->recording = recording;
    q->current_ad_variable_id = 100;
    q->stack_current = 2;
    q->counter = 1;

    struct ad_variable va = {a, ida}, vb = {b, idb}, r = {0.0, 0};
    int ok = -1, inplace = 0;

#define TRY_BIN(nm)                                                               \
    if (strcmp(op, #nm) == 0) {                                                   \
        ok = 0;                                                                   \
        if (family == 0) r = ad_##nm(&g, va, vb);                                 \
        else if (family == 1) r = pad_##nm(&p, va, vb);                           \
        else r = ad_##nm##_p(q, va, vb);                                          \
    }                                                                             \
    if (strcmp(op, #nm "_vd") == 0) {                                             \
        ok = 0;                                                                   \
        if (family == 0) r = ad_##nm##_vd(&g, va, b);                             \
        else if (family == 1) r = pad_##nm##_vd(&p, va, b);                       \
        else r = ad_##nm##_vd_p(q, va, b);                                        \
    }                                                                             \
    if (strcmp(op, #nm "_dv") == 0) {                                             \
        ok = 0;                                                                   \
        if (family == 0) r = ad_##nm##_dv(&g, a, vb);                             \
        else if (family == 1) r = pad_##nm##_dv(&p, a, vb);                       \
        else r = ad_##nm##_dv_p(q, a, vb);                                        \
    }
    OPS_BIN_VV(TRY_BIN)

#define TRY_MATH1(nm)                                                             \
    if (strcmp(op, #nm) == 0) {                                                   \
        if (family == 0) { ok = 0; r = ad_##nm(&g, va); }                         \
        else if (family == 2) { ok = 0; r = ad_##nm##_p(q, va); }                 \
    }
    OPS_MATH1(TRY_MATH1)

    if (strcmp(op, "pow") == 0) {
        if (family == 0) { ok = 0; r = ad_pow(&g, va, vb); }
        else if (family == 2) { ok = 0; r = ad_pow_p(q, va, vb); }
    }
    if (strcmp(op, "pow_vd") == 0) {
        if (family == 0) { ok = 0; r = ad_pow_vd(&g, va, b); }
        else if (family == 2) { ok = 0; r = ad_pow_vd_p(q, va, b); }
    }
    if (strcmp(op, "pow_dv") == 0) {
        if (family == 0) { ok = 0; r = ad_pow_dv(&g, a, vb); }
        else if (family == 2) { ok = 0; r = ad_pow_dv_p(q, a, vb); }
    }
    if (strcmp(op, "plus_eq") == 0) {
        ok = 0; inplace = 1;
        if (family == 0) ad_plus_eq(&g, &va, vb);
        else if (family == 1) pad_plus_eq(&p, &va, vb);
        else ad_plus_eq_p(q, &va, vb);
    }
    if (strcmp(op, "plus_eq_d") == 0) {
        ok = 0; inplace = 1;
        if (family == 0) ad_plus_eq_d(&g, &va, b);
        else if (family == 1) pad_plus_eq_d(&p, &va, b);
        else ad_plus_eq_d_p(q, &va, b);
    }
    if (inplace) r = va;

    if (ok == 0) {
        *value = r.value;
        *id = r.id;
        *lhs_after = va.value;
        /* all three families were primed so that the op lands in slot 1+2 */
        if (family == 2) *entry = q->gradient_stack[3];
        else *entry = stack[3];
    }
    free(q);
    free(stack);
    return ok;
}
