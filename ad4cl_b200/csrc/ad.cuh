/*
 * ad.cuh -- CUDA (sm_100a) device AD runtime with the API of the reference's ad.cl.
 *
 * Drop-in surface: the five struct types of ad.cl:65-97 with the same field
 * names, the init functions of ad.cl:129-161 and all three operator families
 *     ad_<op>(gs, ...)        global-stack family     ad.cl:174-525, 926-1200
 *     pad_<op>(pgs, ...)      block-reserved family   ad.cl:536-924
 *     ad_<op>_p(pgs, ...)     private-stack family    ad.cl:1216-1829
 * with identical names, argument order and value/partial expressions, so the
 * BODY of a kernel written for ad.cl compiles unchanged (see
 * ad4cl_opencl_compat.cuh for the OpenCL keywords).  What changes is where an
 * operator records: `gs` is a per-thread handle onto the thread-private tape
 * of ad4cl_tape.cuh, not a pointer to one global stack, so no operator issues
 * an atomic and the adjoint sweep happens on the device.
 *
 * Differences a kernel can observe, all deliberate:
 *   - ids are per-thread: independents keep the ids the host gave them
 *     (0..P-1), every temporary gets P + the index of its operator's last
 *     operand slot on this thread's tape.
 *   - in-place operators (ad_plus_eq, ...) give the left-hand side a NEW id
 *     instead of re-using it (ad.cl:259); the sweep result is identical.
 *   - pad_init does not reserve a block: pad_* operators append in execution
 *     order to the same thread tape as ad_*.
 *   - the pad_ family also has the math functions (the reference lacks them,
 *     SURVEY.md Q12).
 *
 * AD4CL_REFERENCE_QUIRKS (default 0):
 *   1  reproduce the reference's arithmetic as written, including the swapped
 *      product rule (ad.cl:393-394, 772-773, 1421-1422), the wrong partial of
 *      ad_times_dv in the global/private families (ad.cl:443, 1471) and
 *      ad_cos returning log (ad.cl:927, 1556);
 *   0  mathematically correct partials.
 */
#pragma once

#include "ad4cl_tape.cuh"

#ifndef AD4CL_REFERENCE_QUIRKS
#define AD4CL_REFERENCE_QUIRKS 0
#endif

#ifndef PRIVATE_STACK_SIZE
#define PRIVATE_STACK_SIZE 100 /* ad.cl:51-53 */
#endif
#ifndef PRIVATE_GRADIENT_SIZE
#define PRIVATE_GRADIENT_SIZE 150 /* ad.cl:55-57 */
#endif
#ifndef MAX_VARIABLE_IN_EXPESSION
#define MAX_VARIABLE_IN_EXPESSION 2 /* ad.cl:59-61, spelling kept */
#endif

typedef double real_t; /* ad.cl:29: fp64 only (the structs hard-code double, Q15) */

struct ad_variable { /* ad.cl:65-68 */
    double value;
    int id;
};

struct ad_pair { /* ad.cl:70-73 */
    double dx;
    int id;
};

struct ad_entry { /* ad.cl:75-79: the reference's 40-byte record; export format only */
    struct ad_pair coeff[MAX_VARIABLE_IN_EXPESSION];
    int id;
    int size;
};

struct ad_gradient_structure { /* ad.cl:81-89 + the tape handle */
    struct ad_entry* gradient_stack;
    int current_ad_variable_id;
    int stack_current;
    int recording;
    int counter;
    struct ad_entry* index;
    ad4cl::ThreadTape* tape;
};

/* ad.cl:91-97.  The private stack keeps the compact slot form: 2 slots per entry. */
struct ad_private_gradient_structure {
    double slot_dx[2 * PRIVATE_STACK_SIZE];
    int slot_id[2 * PRIVATE_STACK_SIZE];
    int nslots;
    int current_ad_variable_id;
    int stack_current;
    int recording;
    int counter;
    int overflow;   /* more than PRIVATE_STACK_SIZE entries were recorded: the stack stopped at the limit */
};

#define AD4CL_DEV __device__ __forceinline__

/* feature macros kernel sources can test (both are additions over the reference's ad.cl) */
#define AD4CL_HAVE_PAD_MATH 1
#define AD4CL_HAVE_COMMIT_P 1

/* ------------------------------------------------------------------ init */

/* ad.cl:129-131 */
AD4CL_DEV void ad_init(struct ad_gradient_structure* gs, struct ad_entry* gradient_stack) {
    gs->gradient_stack = gradient_stack;
}

/* ad.cl:133-142 */
AD4CL_DEV void ad_init_p(struct ad_private_gradient_structure* gs) {
    gs->nslots = 0;
    gs->overflow = 0;
    gs->counter = 0;
    gs->current_ad_variable_id = 0;
    gs->recording = 1;
    gs->stack_current = 0;
}

/* ad.cl:144-156: the private structure becomes a second handle on the same tape */
AD4CL_DEV void pad_init(int operations, struct ad_gradient_structure* pgs,
        struct ad_gradient_structure* gs, struct ad_entry* gradient_stack) {
    (void) operations;
    if (gs->recording == 1) {
        pgs->gradient_stack = gradient_stack;
        pgs->counter = gs->counter;
        pgs->current_ad_variable_id = gs->current_ad_variable_id;
        pgs->stack_current = gs->stack_current;
        pgs->recording = gs->recording;
        pgs->index = gs->gradient_stack;
        pgs->tape = gs->tape;
    } else {
        pgs->recording = 0;
        pgs->tape = gs->tape;
    }
}

/* ad.cl:158-161: a fresh thread-local leaf variable */
AD4CL_DEV void ad_init_var_g(struct ad_gradient_structure* gs, struct ad_variable* var, double value) {
    var->value = value;
    if (gs->recording == 1) {
        ad4cl::tape_push<1>(gs->tape, 0.0, 0, ad4cl::kEndFlag | ad4cl::kNopFlag);
        var->id = ad4cl::tape_close_op(gs->tape);
        gs->counter++;
    } else {
        var->id = 0;
    }
}

/* ------------------------------------------------------------------ recorders */

namespace ad4cl {

/* UA / UB = +1 / -1: the partial is the compile-time constant +1.0 / -1.0 (a unit slot owns no cell
   on the tape, ad4cl_tape.cuh); 0: a general partial.  The dx arguments always carry the value. */
template <int UA>
AD4CL_DEV int record1(struct ad_gradient_structure* gs, double dx, int id) {
    ThreadTape* t = gs->tape;
    tape_push<UA>(t, dx, id, kEndFlag);
    const int r = tape_close_op(t);
    gs->counter++;
    return r;
}

template <int UA, int UB>
AD4CL_DEV int record2(struct ad_gradient_structure* gs, double dx0, int id0, double dx1, int id1) {
    ThreadTape* t = gs->tape;
    tape_push2<UA, UB>(t, dx0, id0, dx1, id1);
    const int r = tape_close_op(t);
    gs->counter++;
    return r;
}

/* private family: result id = counter + current_ad_variable_id as in ad.cl:1220-1221.  A stack that
   is full stops recording and says so (ad_commit_p reports it) instead of pairing later results
   with the wrong slots. */
template <int UA>
AD4CL_DEV int record1(struct ad_private_gradient_structure* gs, double dx, int id) {
    const int index = gs->counter++;
    if (gs->nslots < 2 * PRIVATE_STACK_SIZE) {
        gs->slot_dx[gs->nslots] = dx;
        gs->slot_id[gs->nslots] = id | kEndFlag;
        gs->nslots++;
    } else {
        gs->overflow = 1;
    }
    return index + gs->current_ad_variable_id;
}

template <int UA, int UB>
AD4CL_DEV int record2(struct ad_private_gradient_structure* gs, double dx0, int id0, double dx1, int id1) {
    const int index = gs->counter++;
    if (gs->nslots + 1 < 2 * PRIVATE_STACK_SIZE) {
        gs->slot_dx[gs->nslots] = dx0;
        gs->slot_id[gs->nslots] = id0;
        gs->slot_dx[gs->nslots + 1] = dx1;
        gs->slot_id[gs->nslots + 1] = id1 | kEndFlag;
        gs->nslots += 2;
    } else {
        gs->overflow = 1;
    }
    return index + gs->current_ad_variable_id;
}

} // namespace ad4cl

/* ------------------------------------------------------------------ operator table
 * One row per operator: NAME, value expression, partial(s).  Inside a row
 * `av`, `bv` are the operand values and `val` the result value.  The three
 * family generators below stamp each row out for ad_* / pad_* / *_p.
 */
/* exp and log of the operator table, optionally out of line: the fused kernel is bound by instruction
   fetch (profiles/README.md), and one shared copy of a ~60-instruction libdevice expansion instead of one
   per call site shrinks the hot code of a statistical objective by a third.  Same code, same bits. */
#ifndef AD4CL_OUTLINE_MATH
#define AD4CL_OUTLINE_MATH 0
#endif
#if AD4CL_OUTLINE_MATH
static __device__ __noinline__ double ad4cl_m_exp(double x) { return exp(x); }
static __device__ __noinline__ double ad4cl_m_log(double x) { return log(x); }
#else
#define ad4cl_m_exp exp
#define ad4cl_m_log log
#endif

#if AD4CL_REFERENCE_QUIRKS
#define AD4CL_TIMES_DA av /* as written in the reference: the operand's own value */
#define AD4CL_TIMES_DB bv
#define AD4CL_TIMES_DV_DB_QUIRKY bv /* global + private families */
#define AD4CL_COS_VALUE ad4cl_m_log(av)
#else
#define AD4CL_TIMES_DA bv
#define AD4CL_TIMES_DB av
#define AD4CL_TIMES_DV_DB_QUIRKY av
#define AD4CL_COS_VALUE cos(av)
#endif

#define AD4CL_OP_VV(GS_T, NAME, VAL, DA, DB, UA, UB)                                            \
    AD4CL_DEV struct ad_variable NAME(GS_T* gs, const struct ad_variable a, const struct ad_variable b) { \
        const double av = a.value, bv = b.value;                                            \
        const double val = (VAL);                                                           \
        struct ad_variable ret = {val, 0};                                                  \
        (void) av; (void) bv;                                                               \
        if (gs->recording == 1) ret.id = ad4cl::record2<UA, UB>(gs, (DA), a.id, (DB), b.id); \
        return ret;                                                                         \
    }

#define AD4CL_OP_VD(GS_T, NAME, VAL, DA, UA)                                                \
    AD4CL_DEV struct ad_variable NAME(GS_T* gs, const struct ad_variable a, double bv) {   \
        const double av = a.value;                                                          \
        const double val = (VAL);                                                           \
        struct ad_variable ret = {val, 0};                                                  \
        (void) av;                                                                          \
        if (gs->recording == 1) ret.id = ad4cl::record1<UA>(gs, (DA), a.id);                \
        return ret;                                                                         \
    }

#define AD4CL_OP_DV(GS_T, NAME, VAL, DB, UB)                                                \
    AD4CL_DEV struct ad_variable NAME(GS_T* gs, double av, const struct ad_variable b) {   \
        const double bv = b.value;                                                          \
        const double val = (VAL);                                                           \
        struct ad_variable ret = {val, 0};                                                  \
        (void) bv;                                                                          \
        if (gs->recording == 1) ret.id = ad4cl::record1<UB>(gs, (DB), b.id);                \
        return ret;                                                                         \
    }

#define AD4CL_OP_V(GS_T, NAME, VAL, DA)                                                   \
    AD4CL_DEV struct ad_variable NAME(GS_T* gs, const struct ad_variable a) {              \
        const double av = a.value;                                                          \
        const double val = (VAL);                                                           \
        struct ad_variable ret = {val, 0};                                                  \
        if (gs->recording == 1) ret.id = ad4cl::record1<0>(gs, (DA), a.id);                 \
        return ret;                                                                         \
    }

/* in-place forms: ad.cl:249-261 (+=v), 285-297 (+=d) */
#define AD4CL_OP_EQ(GS_T, NAME_V, NAME_D)                                                 \
    AD4CL_DEV void NAME_V(GS_T* gs, struct ad_variable* a, const struct ad_variable b) {   \
        a->value += b.value;                                                                \
        if (gs->recording == 1) a->id = ad4cl::record2<1, 1>(gs, 1.0, a->id, 1.0, b.id);    \
    }                                                                                       \
    AD4CL_DEV void NAME_D(GS_T* gs, struct ad_variable* a, double b) {                     \
        a->value += b;                                                                      \
        if (gs->recording == 1) a->id = ad4cl::record1<1>(gs, 1.0, a->id);                  \
    }

/* Arithmetic rows.  Reference lines (global / pad / private):
 *   plus 174/536/1216  plus_vd 200/572/1242  plus_dv 225/597/1267
 *   minus 307/679/1335 minus_vd 333/705/1361 minus_dv 358/732/1386
 *   times 383/757/1411 times_vd 410/797/1438 times_dv 435/832/1463
 *   divide 460/857/1488 divide_vd 486/883/1514 divide_dv 510/909/1538
 * The divide partials keep the reference's expression order:
 *   inv = 1.0 / b ;  d/da = inv ;  d/db = -1.0 * result * inv.
 */
#define AD4CL_ARITHMETIC(GS_T, PRE, SUF, TIMES_DV_DB)                                              \
    AD4CL_OP_VV(GS_T, PRE##plus##SUF, av + bv, 1.0, 1.0, 1, 1)                                         \
    AD4CL_OP_VD(GS_T, PRE##plus_vd##SUF, av + bv, 1.0, 1)                                            \
    AD4CL_OP_DV(GS_T, PRE##plus_dv##SUF, av + bv, 1.0, 1)                                            \
    AD4CL_OP_EQ(GS_T, PRE##plus_eq##SUF, PRE##plus_eq_d##SUF)                                       \
    AD4CL_OP_VV(GS_T, PRE##minus##SUF, av - bv, 1.0, -1.0, 1, -1)                                      \
    AD4CL_OP_VD(GS_T, PRE##minus_vd##SUF, av - bv, 1.0, 1)                                           \
    AD4CL_OP_DV(GS_T, PRE##minus_dv##SUF, av - bv, -1.0, -1)                                         \
    AD4CL_OP_VV(GS_T, PRE##times##SUF, av * bv, AD4CL_TIMES_DA, AD4CL_TIMES_DB, 0, 0)               \
    AD4CL_OP_VD(GS_T, PRE##times_vd##SUF, av * bv, bv, 0)                                           \
    AD4CL_OP_DV(GS_T, PRE##times_dv##SUF, av * bv, TIMES_DV_DB, 0)                                  \
    AD4CL_OP_VV(GS_T, PRE##divide##SUF, av / bv, 1.0 / bv, -1.0 * val * (1.0 / bv), 0, 0)           \
    AD4CL_OP_VD(GS_T, PRE##divide_vd##SUF, av / bv, 1.0 / bv, 0)                                    \
    AD4CL_OP_DV(GS_T, PRE##divide_dv##SUF, av / bv, -1.0 * val * (1.0 / bv), 0)

/* Math rows (ad.cl:926-1200 global, 1555-1829 private); partial expressions
 * are the reference's, e.g. pow(1 - pow(v,2), 0.5) rather than sqrt. */
#define AD4CL_LN10 2.30258509299404590109361379290930926799774169921875

#define AD4CL_MATH(GS_T, PRE, SUF)                                                                 \
    AD4CL_OP_V(GS_T, PRE##cos##SUF, AD4CL_COS_VALUE, -1.0 * sin(av))                                \
    AD4CL_OP_V(GS_T, PRE##sin##SUF, sin(av), cos(av))                                               \
    AD4CL_OP_V(GS_T, PRE##tan##SUF, tan(av), (1.0 / cos(av)) * (1.0 / cos(av)))                     \
    AD4CL_OP_V(GS_T, PRE##acos##SUF, acos(av), (-1.0) / pow((1.0) - pow(av, 2.0), 0.5))             \
    AD4CL_OP_V(GS_T, PRE##asin##SUF, asin(av), (1.0) / pow((1.0) - pow(av, 2.0), 0.5))              \
    AD4CL_OP_V(GS_T, PRE##atan##SUF, atan(av), (1.0) / (av * av + (1.0)))                           \
    AD4CL_OP_V(GS_T, PRE##cosh##SUF, cosh(av), sinh(av))                                            \
    AD4CL_OP_V(GS_T, PRE##sinh##SUF, sinh(av), cosh(av))                                            \
    AD4CL_OP_V(GS_T, PRE##tanh##SUF, tanh(av), (1.0 / cosh(av)) * (1.0 / cosh(av)))                 \
    AD4CL_OP_V(GS_T, PRE##exp##SUF, ad4cl_m_exp(av), val)                                                   \
    AD4CL_OP_V(GS_T, PRE##log##SUF, ad4cl_m_log(av), 1.0 / av)                                              \
    AD4CL_OP_V(GS_T, PRE##log10##SUF, log10(av), 1.0 / (av * AD4CL_LN10))                           \
    AD4CL_OP_VV(GS_T, PRE##pow##SUF, pow(av, bv), bv * pow(av, bv - (1.0)), log(av) * val, 0, 0)       \
    AD4CL_OP_VD(GS_T, PRE##pow_vd##SUF, pow(av, bv), bv * pow(av, bv - (1.0)), 0)                   \
    AD4CL_OP_DV(GS_T, PRE##pow_dv##SUF, pow(av, bv), log(av) * val, 0)                              \
    AD4CL_OP_V(GS_T, PRE##sqrt##SUF, sqrt(av), .5 / val)

AD4CL_ARITHMETIC(struct ad_gradient_structure, ad_, , AD4CL_TIMES_DV_DB_QUIRKY)
AD4CL_MATH(struct ad_gradient_structure, ad_, )
AD4CL_ARITHMETIC(struct ad_gradient_structure, pad_, , av) /* ad.cl:840 is correct in the reference */
AD4CL_MATH(struct ad_gradient_structure, pad_, )           /* completes the family (Q12) */
AD4CL_ARITHMETIC(struct ad_private_gradient_structure, ad_, _p, AD4CL_TIMES_DV_DB_QUIRKY)
AD4CL_MATH(struct ad_private_gradient_structure, ad_, _p)

/* the __global-lhs twins of ad.cl:263-283, 635-655 */
AD4CL_DEV void ad_plus_eq_g(struct ad_gradient_structure* gs, struct ad_variable* a, const struct ad_variable b) {
    ad_plus_eq(gs, a, b);
}
AD4CL_DEV void pad_plus_eq_g(struct ad_gradient_structure* gs, struct ad_variable* a, const struct ad_variable b) {
    pad_plus_eq(gs, a, b);
}

/* Test / debug accessor: the `back`-th most recent slot of the thread tape (0 = last pushed). */
AD4CL_DEV void ad4cl_tape_peek(const struct ad_gradient_structure* gs, int back, double* dx, int* word) {
    ad4cl::tape_read_row(gs->tape, gs->tape->next_id - gs->tape->first_temp - 1 - back, dx, word);
}

/*
 * Sweep of a private stack (the reference declares the structure, ad.cl:91-97,
 * but has no sweep for it).  gradient[] must hold nvars zeros, nvars >
 * every id on the tape; seeds `out`, walks the slots backwards.
 */
AD4CL_DEV void ad_gradient_p(const struct ad_private_gradient_structure* gs, const struct ad_variable out,
        double* gradient) {
    int p = gs->counter; /* operators recorded */
    int s = gs->nslots;
    if (gs->overflow) { /* the stack stopped at the limit: count the operators it holds */
        p = 0;
        for (int i = 0; i < s; i++) p += (gs->slot_id[i] & ad4cl::kEndFlag) ? 1 : 0;
    }
    gradient[out.id] = 1.0;
    double w = 0.0;
    while (s > 0) {
        --s;
        const int idw = gs->slot_id[s];
        if (idw & ad4cl::kEndFlag) {
            --p;
            const int rid = p + gs->current_ad_variable_id;
            w = gradient[rid];
            gradient[rid] = 0.0;
        }
        gradient[idw & ad4cl::kIdMask] += w * gs->slot_dx[s];
    }
}

/*
 * Pre-accumulation: fold a finished private stack into the thread tape.  Sweeps `pgs`
 * seeded at `out` into a private gradient of PRIVATE_GRADIENT_SIZE adjoints (the macro the
 * reference declares for exactly this, ad.cl:55-57, and never uses) and records
 * out = sum_p g_p * theta_p-linearisation as one short chain of operators whose operands
 * are the independents (ids below gs->current_ad_variable_id).  The private stack lives in
 * registers / local memory, so only P slots per work-item reach the arena instead of one
 * slot per operand.  Ids on the private stack must be independents' ids or the private
 * stack's own results, i.e. pgs->current_ad_variable_id >= gs->current_ad_variable_id.
 * A private stack that overflowed (more than PRIVATE_STACK_SIZE operators) or whose ids do
 * not fit the private gradient (independents + operators > PRIVATE_GRADIENT_SIZE) is reported
 * as a tape overflow (sticky status, the evaluation fails with AD4CL_ERR_TAPE_OVERFLOW)
 * instead of being swept out of bounds.
 */
AD4CL_DEV struct ad_variable ad_commit_p(struct ad_gradient_structure* gs,
        const struct ad_private_gradient_structure* pgs, const struct ad_variable out) {
    struct ad_variable ret = {out.value, 0};
    if (gs->recording != 1) return ret;
    const int used = pgs->current_ad_variable_id + pgs->counter; /* ids the private stack can refer to */
    if (pgs->overflow || used > PRIVATE_GRADIENT_SIZE || out.id < 0 || out.id >= used) {
        gs->tape->status |= ad4cl::kErrTapeOverflow;
        ad_init_var_g(gs, &ret, out.value);
        return ret;
    }
    double g[PRIVATE_GRADIENT_SIZE];
    for (int i = 0; i < used; i++) g[i] = 0.0;
    ad_gradient_p(pgs, out, g);
    const int nparams = gs->current_ad_variable_id;
    int have = 0;
    for (int p = 0; p < nparams && p < PRIVATE_GRADIENT_SIZE; p++) {
        if (g[p] == 0.0) continue;
        ret.id = have ? ad4cl::record2<1, 0>(gs, 1.0, ret.id, g[p], p) : ad4cl::record1<0>(gs, g[p], p);
        have = 1;
    }
    if (!have) ad_init_var_g(gs, &ret, out.value); /* constant: a leaf with no partials */
    return ret;
}
