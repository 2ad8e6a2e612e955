/*
 * oracle/ad_port.h -- TEST INFRASTRUCTURE ONLY (never included by the product).
 *
 * CPU restatement ("port") of the reference's device AD runtime, the algorithm
 * of /root/reference/ad.cl: a Jacobian tape of 40-byte entries on ONE stack
 * shared by all work-items, one entry per overloaded operator, slot reserved
 * with a counter (ad.cl:178-181), ids = slot index + current_ad_variable_id.
 * It exists so that parity can be checked where the reference tree itself is
 * not mounted; it is pinned against the compiled reference (oracle/_ref) and
 * against test/matrix/host.txt by tests/test_oracle.py.
 *
 * Written independently of both the reference text and ad4cl_b200/csrc/ad.cuh:
 * every operator is one row of a table (value, partials, arity) expanded by
 * three family generators.  Each row cites the reference lines it restates.
 *
 *   AD_PORT_QUIRKS=1  reproduce the reference's arithmetic as written, incl.
 *                     Q1 swapped product rule, Q2 times_dv, Q3 cos
 *   AD_PORT_QUIRKS=0  mathematically correct partials (== oracle/_ref/*fix*)
 */
#ifndef AD4CL_ORACLE_AD_PORT_H
#define AD4CL_ORACLE_AD_PORT_H

#include <math.h>

#ifndef AD_PORT_QUIRKS
#define AD_PORT_QUIRKS 1
#endif

#define PRIVATE_STACK_SIZE 100          /* ad.cl:51-53 */
#define MAX_VARIABLE_IN_EXPESSION 2     /* ad.cl:59-61 (spelling is the reference's) */

/* ad.cl:65-79 -- 16, 16 and 40 bytes */
struct ad_variable { double value; int id; };
struct ad_pair { double dx; int id; };
struct ad_entry { struct ad_pair coeff[MAX_VARIABLE_IN_EXPESSION]; int id; int size; };

/* ad.cl:81-89 */
struct ad_gradient_structure {
    struct ad_entry* gradient_stack;
    int current_ad_variable_id;
    int stack_current;
    int recording;
    int counter;
    struct ad_entry* index;
};

/* ad.cl:91-97 */
struct ad_private_gradient_structure {
    struct ad_entry gradient_stack[PRIVATE_STACK_SIZE];
    int current_ad_variable_id;
    int stack_current;
    int recording;
    int counter;
};

/* ad.cl:129-131 */
static inline void ad_init(struct ad_gradient_structure* gs, struct ad_entry* gradient_stack) {
    gs->gradient_stack = gradient_stack;
}

/* ad.cl:133-142 */
static inline void ad_init_p(struct ad_private_gradient_structure* gs) {
    for (int i = 0; i < PRIVATE_STACK_SIZE; i++) {
        gs->gradient_stack[i].id = 0;
        gs->gradient_stack[i].size = 0;
    }
    gs->counter = 0;
    gs->current_ad_variable_id = 0;
    gs->recording = 1;
    gs->stack_current = 0;
}

/* ad.cl:144-156: ONE atomic reserves `operations` consecutive slots */
static inline void pad_init(int operations, struct ad_gradient_structure* pgs,
        struct ad_gradient_structure* gs, struct ad_entry* gradient_stack) {
    if (gs->recording == 1) {
        pgs->gradient_stack = gradient_stack;
        pgs->counter = __sync_fetch_and_add(&gs->counter, operations);
        pgs->current_ad_variable_id = gs->current_ad_variable_id;
        pgs->stack_current = gs->stack_current;
        pgs->recording = gs->recording;
        pgs->index = &gs->gradient_stack[gs->stack_current];
    } else {
        pgs->recording = 0;
    }
}

/* ad.cl:158-161 (id taken from the variable counter, not the slot counter: Q6) */
static inline void ad_init_var_g(struct ad_gradient_structure* gs, struct ad_variable* var, double value) {
    var->id = __sync_fetch_and_add(&gs->current_ad_variable_id, 1);
    var->value = value;
}

/* ---- slot reservation per family ---- */
#define PORT_SLOT_g(gs) __sync_fetch_and_add(&(gs)->counter, 1)   /* ad.cl:178 */
#define PORT_SLOT_pad(gs) ((gs)->counter++)                        /* ad.cl:540 */
#define PORT_SLOT_p(gs) ((gs)->counter++)                          /* ad.cl:1220 */

static inline void port_store(struct ad_entry* e, int id, int size,
        double dx0, int id0, double dx1, int id1) {
    e->coeff[0].dx = dx0;
    e->coeff[0].id = id0;
    if (size == 2) {
        e->coeff[1].dx = dx1;
        e->coeff[1].id = id1;
    }
    e->id = id;
    e->size = size;
}

/*
 * Generators.  FAM is g / pad / p; GS_T the structure type; NAME the exported
 * function name.  VAL, DA, DB are expressions over (av, bv) = operand values
 * and `val` = the result value.
 */
#define PORT_VV(FAM, GS_T, NAME, VAL, DA, DB)                                          \
    static inline struct ad_variable NAME(GS_T* gs, struct ad_variable a, struct ad_variable b) { \
        const double av = a.value, bv = b.value;                                        \
        const double val = (VAL);                                                       \
        struct ad_variable ret = {val, 0};                                              \
        (void) av; (void) bv;                                                           \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            ret.id = index + gs->current_ad_variable_id;                                \
            port_store(&gs->gradient_stack[index + gs->stack_current], ret.id, 2,       \
                    (DA), a.id, (DB), b.id);                                            \
        }                                                                               \
        return ret;                                                                     \
    }

#define PORT_VD(FAM, GS_T, NAME, VAL, DA)                                              \
    static inline struct ad_variable NAME(GS_T* gs, struct ad_variable a, double bv) { \
        const double av = a.value;                                                      \
        const double val = (VAL);                                                       \
        struct ad_variable ret = {val, 0};                                              \
        (void) av;                                                                      \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            ret.id = index + gs->current_ad_variable_id;                                \
            port_store(&gs->gradient_stack[index + gs->stack_current], ret.id, 1,       \
                    (DA), a.id, 0.0, 0);                                                \
        }                                                                               \
        return ret;                                                                     \
    }

#define PORT_DV(FAM, GS_T, NAME, VAL, DB)                                              \
    static inline struct ad_variable NAME(GS_T* gs, double av, struct ad_variable b) { \
        const double bv = b.value;                                                      \
        const double val = (VAL);                                                       \
        struct ad_variable ret = {val, 0};                                              \
        (void) bv;                                                                      \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            ret.id = index + gs->current_ad_variable_id;                                \
            port_store(&gs->gradient_stack[index + gs->stack_current], ret.id, 1,       \
                    (DB), b.id, 0.0, 0);                                                \
        }                                                                               \
        return ret;                                                                     \
    }

#define PORT_V(FAM, GS_T, NAME, VAL, DA)                                               \
    static inline struct ad_variable NAME(GS_T* gs, struct ad_variable a) {            \
        const double av = a.value;                                                      \
        const double val = (VAL);                                                       \
        struct ad_variable ret = {val, 0};                                              \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            ret.id = index + gs->current_ad_variable_id;                                \
            port_store(&gs->gradient_stack[index + gs->stack_current], ret.id, 1,       \
                    (DA), a.id, 0.0, 0);                                                \
        }                                                                               \
        return ret;                                                                     \
    }

/* in-place: the entry re-uses the left-hand id (ad.cl:249-261, 285-297) */
#define PORT_EQ_V(FAM, GS_T, NAME)                                                     \
    static inline void NAME(GS_T* gs, struct ad_variable* a, struct ad_variable b) {    \
        a->value += b.value;                                                            \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            port_store(&gs->gradient_stack[index + gs->stack_current], a->id, 2,        \
                    1.0, a->id, 1.0, b.id);                                             \
        }                                                                               \
    }
#define PORT_EQ_D(FAM, GS_T, NAME)                                                     \
    static inline void NAME(GS_T* gs, struct ad_variable* a, double b) {                \
        a->value += b;                                                                  \
        if (gs->recording == 1) {                                                       \
            int index = PORT_SLOT_##FAM(gs);                                            \
            port_store(&gs->gradient_stack[index + gs->stack_current], a->id, 1,        \
                    1.0, a->id, 0.0, 0);                                                \
        }                                                                               \
    }

/* ---- the operator table ---- */
#if AD_PORT_QUIRKS
#define PORT_TIMES_DA av            /* Q1: ad.cl:393, 772, 1421 store the operand's OWN value */
#define PORT_TIMES_DB bv            /* Q1: ad.cl:394, 773, 1422 */
#define PORT_TIMES_DV_DB_g bv       /* Q2: ad.cl:443 */
#define PORT_TIMES_DV_DB_p bv       /* Q2: ad.cl:1471 */
#define PORT_COS_VAL log(av)        /* Q3: ad.cl:927, 1556 */
#else
#define PORT_TIMES_DA bv
#define PORT_TIMES_DB av
#define PORT_TIMES_DV_DB_g av
#define PORT_TIMES_DV_DB_p av
#define PORT_COS_VAL cos(av)
#endif
#define PORT_TIMES_DV_DB_pad av     /* ad.cl:840 is correct in the reference */

#define PORT_LN10 2.30258509299404590109361379290930926799774169921875 /* ad.cl:1122 */

/* arithmetic: all three families.
 *   plus 174/536/1216   plus_vd 200/572/1242   plus_dv 225/597/1267
 *   minus 307/679/1335  minus_vd 333/705/1361  minus_dv 358/732/1386
 *   times 383/757/1411  times_vd 410/797/1438  times_dv 435/832/1463
 *   divide 460/857/1488 divide_vd 486/883/1514 divide_dv 510/909/1538  (ad.cl lines)
 * divide partials exactly as written: inv = 1/b ; -1.0 * ret * inv.
 */
#define PORT_ARITH(FAM, GS_T, PRE, SUF)                                                          \
    PORT_VV(FAM, GS_T, PRE##plus##SUF, av + bv, 1.0, 1.0)                                         \
    PORT_VD(FAM, GS_T, PRE##plus_vd##SUF, av + bv, 1.0)                                           \
    PORT_DV(FAM, GS_T, PRE##plus_dv##SUF, av + bv, 1.0)                                           \
    PORT_EQ_V(FAM, GS_T, PRE##plus_eq##SUF)                                                       \
    PORT_EQ_D(FAM, GS_T, PRE##plus_eq_d##SUF)                                                     \
    PORT_VV(FAM, GS_T, PRE##minus##SUF, av - bv, 1.0, -1.0)                                       \
    PORT_VD(FAM, GS_T, PRE##minus_vd##SUF, av - bv, 1.0)                                          \
    PORT_DV(FAM, GS_T, PRE##minus_dv##SUF, av - bv, -1.0)                                         \
    PORT_VV(FAM, GS_T, PRE##times##SUF, av * bv, PORT_TIMES_DA, PORT_TIMES_DB)                    \
    PORT_VD(FAM, GS_T, PRE##times_vd##SUF, av * bv, bv)                                           \
    PORT_DV(FAM, GS_T, PRE##times_dv##SUF, av * bv, PORT_TIMES_DV_DB_##FAM)                       \
    PORT_VV(FAM, GS_T, PRE##divide##SUF, av / bv, 1.0 / bv, -1.0 * val * (1.0 / bv))              \
    PORT_VD(FAM, GS_T, PRE##divide_vd##SUF, av / bv, 1.0 / bv)                                    \
    PORT_DV(FAM, GS_T, PRE##divide_dv##SUF, av / bv, -1.0 * val * (1.0 / bv))

/* math: global (ad.cl:926-1200) and private (ad.cl:1555-1829) families only (Q12) */
#define PORT_MATH(FAM, GS_T, SUF)                                                                 \
    PORT_V(FAM, GS_T, ad_cos##SUF, PORT_COS_VAL, -1.0 * sin(av))                                  \
    PORT_V(FAM, GS_T, ad_sin##SUF, sin(av), cos(av))                                              \
    PORT_V(FAM, GS_T, ad_tan##SUF, tan(av), (1.0 / cos(av)) * (1.0 / cos(av)))                    \
    PORT_V(FAM, GS_T, ad_acos##SUF, acos(av), (-1.0) / pow((1.0) - pow(av, 2.0), 0.5))            \
    PORT_V(FAM, GS_T, ad_asin##SUF, asin(av), (1.0) / pow((1.0) - pow(av, 2.0), 0.5))             \
    PORT_V(FAM, GS_T, ad_atan##SUF, atan(av), (1.0) / (av * av + (1.0)))                          \
    PORT_V(FAM, GS_T, ad_cosh##SUF, cosh(av), sinh(av))                                           \
    PORT_V(FAM, GS_T, ad_sinh##SUF, sinh(av), cosh(av))                                           \
    PORT_V(FAM, GS_T, ad_tanh##SUF, tanh(av), (1.0 / cosh(av)) * (1.0 / cosh(av)))                \
    PORT_V(FAM, GS_T, ad_exp##SUF, exp(av), val)                                                  \
    PORT_V(FAM, GS_T, ad_log##SUF, log(av), 1.0 / av)                                             \
    PORT_V(FAM, GS_T, ad_log10##SUF, log10(av), 1.0 / (av * PORT_LN10))                           \
    PORT_VV(FAM, GS_T, ad_pow##SUF, pow(av, bv), bv * pow(av, bv - (1.0)), log(av) * val)         \
    PORT_VD(FAM, GS_T, ad_pow_vd##SUF, pow(av, bv), bv * pow(av, bv - (1.0)))                     \
    PORT_DV(FAM, GS_T, ad_pow_dv##SUF, pow(av, bv), log(av) * val)                                \
    PORT_V(FAM, GS_T, ad_sqrt##SUF, sqrt(av), .5 / val)

PORT_ARITH(g, struct ad_gradient_structure, ad_, )
PORT_MATH(g, struct ad_gradient_structure, )
PORT_ARITH(pad, struct ad_gradient_structure, pad_, )
PORT_ARITH(p, struct ad_private_gradient_structure, ad_, _p)
PORT_MATH(p, struct ad_private_gradient_structure, _p)

/* the __global-lhs twins (ad.cl:263-283, 635-655) */
#define ad_plus_eq_g ad_plus_eq
#define pad_plus_eq_g pad_plus_eq

/*
 * The adjoint sweep, ad4cl.h:775-802 (== simple.cpp:30-47, simple.cl:2-29):
 * seed the id of the LAST entry, walk the stack downwards; for each entry take
 * w = g[id], zero it, scatter w*dx to the operands.  g must hold nvars+1 zeros.
 */
static inline void port_sweep(const struct ad_entry* stack, int stack_current, double* g) {
    if (stack_current <= 0) return;
    g[stack[stack_current - 1].id] = 1.0;
    for (int j = stack_current - 1; j >= 0; j--) {
        const int id = stack[j].id;
        const double w = g[id];
        g[id] = 0.0;
        for (int i = 0; i < stack[j].size; i++) {
            g[stack[j].coeff[i].id] += w * stack[j].coeff[i].dx;
        }
    }
}

#endif
