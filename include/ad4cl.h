/*
 * ad4cl.h -- host-side API of AD4CL, source compatible with the reference's
 * ad4cl.h (same type, field, macro and function names; same ownership rules),
 * for host programs that now run their kernels through libad4cl_cuda.so.
 *
 * What this header is: the SCALAR host tape of the reference (ad4cl.h:39-802).
 * Host programs use it for the few operations that follow a launch -- the
 * epilogue f = (N/2) log(sum), test/admb/simple.cpp:357-365, main.cpp:408-417 --
 * and for compute_gradient().  It is not a fallback for the device path: the
 * per-work-item recording, the adjoint sweep over the work-items' tapes and the
 * gradient reduction exist only in CUDA (ad4cl_cuda.h).  The bridge at the end,
 * ad4cl_cuda_eval_to_tape(), replaces the reference's "launch, read the whole
 * stack back, gpu_restore, sum out[] on the host" (simple.cpp:283-361) by one
 * device evaluation whose result lands on this host tape as P entries.
 *
 * Differences from the reference header, all fixes of defects listed in
 * SURVEY.md 2.3:
 *   Q16  every function is `static inline` (the reference defines three
 *        non-inline functions in the header: ODR violation in 2 TUs);
 *   Q4   ad_plus_vd records size 1 (the reference sets size 2 with coeff[1]
 *        uninitialised);
 *   Q5   ad_plus_dv / ad_divide_vd take ONE id (the reference increments twice);
 *   Q1/Q3 product rule and cos are correct unless AD4CL_REFERENCE_QUIRKS is 1;
 *   create_gradient_structure initialises `counter`; the structs also compile as C.
 */
#ifndef ADCL_H
#define ADCL_H
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef DEFAULT_ENTRY_SIZE
#define DEFAULT_ENTRY_SIZE 10000000 /* ad4cl.h:15-17 */
#endif

#ifndef MAX_VARIABLE_IN_EXPESSION
#define MAX_VARIABLE_IN_EXPESSION 2 /* ad4cl.h:19-21, spelling kept */
#endif

#ifndef AD4CL_REFERENCE_QUIRKS
#define AD4CL_REFERENCE_QUIRKS 0
#endif

#ifdef __cplusplus
extern "C" {
#endif

struct ad_variable { /* ad4cl.h:39-42 */
    double value;
    int id;
};

struct ad_pair { /* ad4cl.h:44-47 */
    double dx;
    int id;
};

struct ad_entry { /* ad4cl.h:49-54 */
    struct ad_pair coeff[MAX_VARIABLE_IN_EXPESSION];
    int id;
    int size;
};

struct ad_gradient_structure { /* ad4cl.h:55-61 */
    struct ad_entry* gradient_stack;
    int current_variable_id;
    int stack_current;
    int recording;
    int counter;
};

/* ad4cl.h:68-75: malloc'd, owned by the caller (the reference never frees it) */
static inline struct ad_gradient_structure* create_gradient_structure(int size) {
    struct ad_gradient_structure* gs = (struct ad_gradient_structure*) malloc(sizeof (struct ad_gradient_structure));
    gs->current_variable_id = 0;
    gs->recording = 1;
    gs->stack_current = 0;
    gs->counter = 0;
    gs->gradient_stack = (struct ad_entry*) malloc(sizeof (struct ad_entry) * (size_t) size);
    return gs;
}

/* ad4cl.h:763-773 */
static inline struct ad_entry* create_entries(int size) {
    struct ad_entry* e = (struct ad_entry*) malloc((size_t) size * sizeof (struct ad_entry));
    for (int i = 0; i < size; i++) {
        e[i].id = 0;
        e[i].size = 0;
    }
    return e;
}

/* ad4cl.h:83-87: fold what the device recorded into the host counters.  With the
   CUDA path nothing is recorded on the host stack by a launch, so counter stays 0. */
static inline void gpu_restore(struct ad_gradient_structure* gs) {
    gs->current_variable_id += gs->counter;
    gs->stack_current += gs->counter;
    gs->counter = 0;
}

/* ad4cl.h:90-93 */
static inline void ad_init_var(struct ad_gradient_structure* gs, struct ad_variable* var, double value) {
    var->id = gs->current_variable_id++;
    var->value = value;
}

static inline void ad4cl_host_store(struct ad_gradient_structure* gs, int id, int size,
        double dx0, int id0, double dx1, int id1) {
    struct ad_entry* e = &gs->gradient_stack[gs->stack_current++];
    e->coeff[0].dx = dx0;
    e->coeff[0].id = id0;
    e->coeff[1].dx = dx1;
    e->coeff[1].id = id1;
    e->size = size;
    e->id = id;
}

/* One row per operator: value and partial expressions as in ad4cl.h:104-761. */
#define AD4CL_HOST_VV(NAME, VAL, DA, DB)                                                         \
    static inline struct ad_variable NAME(struct ad_gradient_structure* gs, struct ad_variable a, struct ad_variable b) { \
        const double av = a.value, bv = b.value;                                                  \
        const double val = (VAL);                                                                 \
        struct ad_variable ret;                                                                   \
        ret.value = val;                                                                          \
        ret.id = 0;                                                                               \
        (void) av; (void) bv;                                                                     \
        if (gs->recording == 1) {                                                                 \
            ret.id = gs->current_variable_id++;                                                   \
            ad4cl_host_store(gs, ret.id, 2, (DA), a.id, (DB), b.id);                              \
        }                                                                                         \
        return ret;                                                                               \
    }
#define AD4CL_HOST_VD(NAME, VAL, DA)                                                             \
    static inline struct ad_variable NAME(struct ad_gradient_structure* gs, struct ad_variable a, double bv) { \
        const double av = a.value;                                                                \
        const double val = (VAL);                                                                 \
        struct ad_variable ret;                                                                   \
        ret.value = val;                                                                          \
        ret.id = 0;                                                                               \
        (void) av;                                                                                \
        if (gs->recording == 1) {                                                                 \
            ret.id = gs->current_variable_id++;                                                   \
            ad4cl_host_store(gs, ret.id, 1, (DA), a.id, 0.0, 0);                                  \
        }                                                                                         \
        return ret;                                                                               \
    }
#define AD4CL_HOST_DV(NAME, VAL, DB)                                                             \
    static inline struct ad_variable NAME(struct ad_gradient_structure* gs, double av, struct ad_variable b) { \
        const double bv = b.value;                                                                \
        const double val = (VAL);                                                                 \
        struct ad_variable ret;                                                                   \
        ret.value = val;                                                                          \
        ret.id = 0;                                                                               \
        (void) bv;                                                                                \
        if (gs->recording == 1) {                                                                 \
            ret.id = gs->current_variable_id++;                                                   \
            ad4cl_host_store(gs, ret.id, 1, (DB), b.id, 0.0, 0);                                  \
        }                                                                                         \
        return ret;                                                                               \
    }
#define AD4CL_HOST_V(NAME, VAL, DA)                                                              \
    static inline struct ad_variable NAME(struct ad_gradient_structure* gs, struct ad_variable a) { \
        const double av = a.value;                                                                \
        const double val = (VAL);                                                                 \
        struct ad_variable ret;                                                                   \
        ret.value = val;                                                                          \
        ret.id = 0;                                                                               \
        if (gs->recording == 1) {                                                                 \
            ret.id = gs->current_variable_id++;                                                   \
            ad4cl_host_store(gs, ret.id, 1, (DA), a.id, 0.0, 0);                                  \
        }                                                                                         \
        return ret;                                                                               \
    }

#if AD4CL_REFERENCE_QUIRKS
#define AD4CL_HOST_TIMES_DA av /* ad4cl.h:321-322 as written */
#define AD4CL_HOST_TIMES_DB bv
#define AD4CL_HOST_COS log(av) /* ad4cl.h:469 */
#else
#define AD4CL_HOST_TIMES_DA bv
#define AD4CL_HOST_TIMES_DB av
#define AD4CL_HOST_COS cos(av)
#endif

AD4CL_HOST_VV(ad_plus, av + bv, 1.0, 1.0)                                   /* ad4cl.h:104 */
AD4CL_HOST_VD(ad_plus_vd, av + bv, 1.0)                                     /* :131 */
AD4CL_HOST_DV(ad_plus_dv, av + bv, 1.0)                                     /* :160 */
AD4CL_HOST_VV(ad_minus, av - bv, 1.0, -1.0)                                 /* :232 */
AD4CL_HOST_VD(ad_minus_vd, av - bv, 1.0)                                    /* :260 */
AD4CL_HOST_DV(ad_minus_dv, av - bv, -1.0)                                   /* :287 */
AD4CL_HOST_VV(ad_times, av * bv, AD4CL_HOST_TIMES_DA, AD4CL_HOST_TIMES_DB)  /* :314 */
AD4CL_HOST_VD(ad_times_vd, av * bv, bv)                                     /* :341 */
AD4CL_HOST_DV(ad_times_dv, av * bv, av)                                     /* :368 */
AD4CL_HOST_VV(ad_divide, av / bv, 1.0 / bv, -1.0 * val * (1.0 / bv))        /* :395 */
AD4CL_HOST_VD(ad_divide_vd, av / bv, 1.0 / bv)                              /* :423 */
AD4CL_HOST_DV(ad_divide_dv, av / bv, -1.0 * val * (1.0 / bv))               /* :449 */
AD4CL_HOST_V(ad_cos, AD4CL_HOST_COS, -1.0 * sin(av))                        /* :468 */
AD4CL_HOST_V(ad_sin, sin(av), cos(av))                                      /* :486 */
AD4CL_HOST_V(ad_tan, tan(av), (1.0 / cos(av)) * (1.0 / cos(av)))            /* :504 */
AD4CL_HOST_V(ad_acos, acos(av), (-1.0) / pow((1.0) - pow(av, 2.0), 0.5))    /* :522 */
AD4CL_HOST_V(ad_asin, asin(av), (1.0) / pow((1.0) - pow(av, 2.0), 0.5))     /* :543 */
AD4CL_HOST_V(ad_atan, atan(av), (1.0) / (av * av + (1.0)))                  /* :564 */
AD4CL_HOST_V(ad_cosh, cosh(av), sinh(av))                                   /* :582 */
AD4CL_HOST_V(ad_sinh, sinh(av), cosh(av))                                   /* :599 */
AD4CL_HOST_V(ad_tanh, tanh(av), (1.0 / cosh(av)) * (1.0 / cosh(av)))        /* :616 */
AD4CL_HOST_V(ad_exp, exp(av), val)                                          /* :634 */
AD4CL_HOST_V(ad_log, log(av), 1.0 / av)                                     /* :651 */
AD4CL_HOST_V(ad_log10, log10(av), 1.0 / (av * 2.30258509299404590109361379290930926799774169921875)) /* :669 */
AD4CL_HOST_VV(ad_pow, pow(av, bv), bv * pow(av, bv - (1.0)), log(av) * val) /* :687 */
AD4CL_HOST_VD(ad_pow_vd, pow(av, bv), bv * pow(av, bv - (1.0)))             /* :707 */
AD4CL_HOST_DV(ad_pow_dv, pow(av, bv), log(av) * val)                        /* :726 */
AD4CL_HOST_V(ad_sqrt, sqrt(av), .5 / val)                                   /* :745 */

/* ad4cl.h:186-207: in place, the entry re-uses the left-hand id */
static inline void ad_plus_eq_v(struct ad_gradient_structure* gs, struct ad_variable* a, struct ad_variable b) {
    a->value += b.value;
    if (gs->recording == 1) ad4cl_host_store(gs, a->id, 2, 1.0, a->id, 1.0, b.id);
}

/* ad4cl.h:209-230 */
static inline void ad_plus_eq_d(struct ad_gradient_structure* gs, struct ad_variable* a, double b) {
    a->value += b;
    if (gs->recording == 1) ad4cl_host_store(gs, a->id, 1, 1.0, a->id, 0.0, 0);
}

/*
 * ad4cl.h:775-802: the adjoint sweep over the HOST stack.  Seeds the id of the
 * last entry, returns a malloc'd array of *size = current_variable_id + 1
 * adjoints indexed by variable id; the caller frees it (main.cpp:429).
 * (C callers: pass pointers; C++ callers may use the reference-style overload below.)
 */
static inline double* ad4cl_compute_gradient(struct ad_gradient_structure* gs, int* size) {
    double* gradient = NULL;
    if (gs->recording == 1 && gs->stack_current > 0) {
        *size = gs->current_variable_id + 1;
        gradient = (double*) calloc((size_t) *size, sizeof (double));
        gradient[gs->gradient_stack[gs->stack_current - 1].id] = 1.0;
        for (int j = gs->stack_current - 1; j >= 0; j--) {
            const struct ad_entry* e = &gs->gradient_stack[j];
            const double w = gradient[e->id];
            gradient[e->id] = 0.0;
            for (int i = 0; i < e->size; i++) gradient[e->coeff[i].id] += w * e->coeff[i].dx;
        }
    }
    return gradient;
}

#ifdef AD4CL_CUDA_H
/*
 * Device evaluation onto the host tape.  Runs `kernel` (configured with
 * AD4CL_EPILOGUE_SUM) at the current values of the P independents `params`
 * (ids as given by ad_init_var), and returns S = sum_i out_i as a host variable
 * whose P tape entries carry dS/dparams -- exactly what the host sum of out[]
 * over the read-back device stack leaves behind in the reference
 * (simple.cpp:283-361), minus the 40 N + 16 N bytes of PCIe traffic.
 * Returns AD4CL_OK or the error of ad4cl_cuda_eval.
 */
static inline int ad4cl_cuda_eval_to_tape(ad4cl_kernel* kernel, struct ad_gradient_structure* gs,
        const struct ad_variable* params, int nparams, struct ad_variable* sum) {
    double* theta = (double*) malloc(sizeof (double) * (size_t) (2 * nparams + 1));
    double* grad = theta + nparams;
    for (int p = 0; p < nparams; p++) theta[p] = params[p].value;
    double S = 0.0;
    int rc = ad4cl_cuda_eval(kernel, theta, nparams, &S, grad);
    if (rc == AD4CL_OK) {
        sum->value = S;
        sum->id = 0;
        if (gs->recording == 1) {
            sum->id = gs->current_variable_id++;
            for (int p = 0; p < nparams; p++)
                ad4cl_host_store(gs, sum->id, 2, 1.0, sum->id, grad[p], params[p].id);
        }
    }
    free(theta);
    return rc;
}
#endif

#ifdef __cplusplus
}

/* the reference's signature (ad4cl.h:775): references, C++ only */
static inline double* compute_gradient(struct ad_gradient_structure& gs, int& size) {
    return ad4cl_compute_gradient(&gs, &size);
}
#endif

#endif /* ADCL_H */
