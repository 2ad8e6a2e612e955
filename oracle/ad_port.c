/*
 * oracle/ad_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives oracle/ad_port.h (the CPU restatement of ad.cl) through the
 * reference's per-evaluation protocol, test/admb/simple.cpp:270-384:
 *   record kernel over all work-items -> gpu_restore (ad4cl.h:83-87)
 *   -> host sum of out[] (simple.cpp:357-361) -> scalar epilogue (simple.cpp:365)
 *   -> serial adjoint sweep (ad4cl.h:775-802) -> g[param.id].
 * Compiled twice by oracle/Makefile: PSYM(x) = port_q_##x with the reference's
 * arithmetic quirks, port_f_##x without.
 *
 * Callers: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Never the product path.
 */
/* two builds of this TU (quirks on/off) share one .so: keep the kernels file-local */
#define __kernel static
#include "clshim.h"
#include "ad_port.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include OBJECTIVES_CL

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double) ts.tv_sec + 1e-6 * (double) ts.tv_nsec;
}

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

/* host-side tape append used by the epilogue: ad4cl.h host ops take their slot
   from stack_current and their id from current_variable_id (ad4cl.h:108-109) */
static void host_plus_eq(struct ad_gradient_structure* gs, struct ad_variable* a, struct ad_variable b) {
    a->value += b.value;                                             /* ad4cl.h:186-199 */
    port_store(&gs->gradient_stack[gs->stack_current++], a->id, 2, 1.0, a->id, 1.0, b.id);
}

static struct ad_variable host_log(struct ad_gradient_structure* gs, struct ad_variable v) {
    struct ad_variable r = {log(v.value), gs->current_ad_variable_id++};  /* ad4cl.h:651-667 */
    port_store(&gs->gradient_stack[gs->stack_current++], r.id, 1, 1.0 / v.value, v.id, 0.0, 0);
    return r;
}

static struct ad_variable host_times_dv(struct ad_gradient_structure* gs, double a, struct ad_variable b) {
    struct ad_variable r = {a * b.value, gs->current_ad_variable_id++};   /* ad4cl.h:368-384: dx = a */
    port_store(&gs->gradient_stack[gs->stack_current++], r.id, 1, a, b.id, 0.0, 0);
    return r;
}

/*
 * One objective+gradient evaluation of a kernel of objectives.cl.
 *   epilogue 0: f = sum_i out_i        1: f = (size/2) log sum_i out_i
 * Returns 0; -1 unknown kernel; -2 out of memory; -3 tape would overflow.
 */
int PSYM(eval_objective)(const char* name, int nparams, const double* theta,
        double* d0, double* d1, int size, int i0, int i1, int global_size,
        int epilogue, int threads, long long stack_entries,
        double* f, double* grad, long long* entries, double* t_ms) {
    struct ad_gradient_structure gs;
    memset(&gs, 0, sizeof gs);
    gs.recording = 1;
    struct ad_entry* stack = (struct ad_entry*) calloc((size_t) stack_entries, sizeof (struct ad_entry));
    struct ad_variable* th = (struct ad_variable*) malloc(sizeof (struct ad_variable) * (size_t) (nparams > 0 ? nparams : 1));
    struct ad_variable* out = (struct ad_variable*) calloc((size_t) (size > 0 ? size : 1), sizeof (struct ad_variable));
    if (!stack || !th || !out) { free(stack); free(th); free(out); return -2; }
    gs.gradient_stack = stack;
    for (int p = 0; p < nparams; p++) {              /* ad_init_var, ad4cl.h:90-93 */
        th[p].value = theta[p];
        th[p].id = gs.current_ad_variable_id++;
    }

    int rc = 0;
    double t0 = now_ms();
    if (!strcmp(name, "linreg2_g")) {
        NDRANGE_1D(global_size, threads, linreg2_g(&gs, stack, &th[0], &th[1], d0, d1, out, size));
    } else if (!strcmp(name, "linreg2_p")) {
        NDRANGE_1D(global_size, threads, linreg2_p(&gs, stack, &th[0], &th[1], d0, d1, out, size));
    } else if (!strcmp(name, "regression_linear")) {
        NDRANGE_1D(global_size, threads, regression_linear(&gs, stack, th, d0, d1, out, size, i0, i1));
    } else if (!strcmp(name, "regression_logistic")) {
        NDRANGE_1D(global_size, threads, regression_logistic(&gs, stack, th, d0, d1, out, size, i0, i1));
    } else if (!strcmp(name, "catch_at_age")) {
        NDRANGE_1D(global_size, threads, catch_at_age(&gs, stack, th, d0, d1, out, size, i0, i1));
    } else if (!strcmp(name, "tape_stress")) {
        NDRANGE_1D(global_size, threads, tape_stress(&gs, stack, th, d0, d1, out, size, i0, i1));
    } else if (!strcmp(name, "op_zoo")) {
        NDRANGE_1D(global_size, threads, op_zoo(&gs, stack, th, d0, d1, out, size, i0, i1));
    } else {
        rc = -1;
    }
    double t1 = now_ms();
    if (rc == 0) {
        /* gpu_restore, ad4cl.h:83-87 */
        gs.current_ad_variable_id += gs.counter;
        gs.stack_current += gs.counter;
        gs.counter = 0;
        if ((long long) gs.stack_current + size + 2 > stack_entries) rc = -3;
    }
    if (rc == 0) {
        struct ad_variable sum = {0.0, gs.current_ad_variable_id++};   /* simple.cpp:357 */
        for (int i = 0; i < size; i++) host_plus_eq(&gs, &sum, out[i]);
        struct ad_variable ff = sum;
        if (epilogue == 1) ff = host_times_dv(&gs, (double) size / 2.0, host_log(&gs, sum));
        double t2 = now_ms();
        double* g = (double*) calloc((size_t) gs.current_ad_variable_id + 1, sizeof (double));
        if (!g) { rc = -2; }
        else {
            port_sweep(stack, gs.stack_current, g);
            double t3 = now_ms();
            *f = ff.value;
            for (int p = 0; p < nparams; p++) grad[p] = g[p];
            *entries = gs.stack_current;
            if (t_ms) { t_ms[0] = t1 - t0; t_ms[1] = t2 - t1; t_ms[2] = t3 - t2; }
            free(g);
        }
    }
    free(stack); free(th); free(out);
    return rc;
}

/*
 * test/matrix: C = A*B, restating the host twin MatrixMultHost
 * (matrix_mul.cpp:39-57) with ad_plus_eq / ad_times entries, then the host sum
 * of C (matrix_mul.cpp:253-261).  A ids 0..n^2-1, B ids n^2..2n^2-1; every C
 * accumulator and the final sum take a fresh id.  grad may be NULL.
 */
int PSYM(matrix)(int n, const double* A, const double* B, double* C_out,
        long long* entries, double* grad) {
    const long long cap = 2LL * n * n * n + 2LL * n * n + 16;
    struct ad_entry* stack = (struct ad_entry*) calloc((size_t) cap, sizeof (struct ad_entry));
    if (!stack) return -2;
    int next_id = 2 * n * n;
    long long top = 0;
    int* cid = (int*) malloc(sizeof (int) * (size_t) n * n);
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) {
            double value = 0.0;
            const int vid = next_id++;
            for (int k = 0; k < n; k++) {
                const int ia = k + j * n, ib = n * n + k * n + i;
                const double av = A[k + j * n], bv = B[k * n + i];
                const int tid = next_id++;
                port_store(&stack[top++], tid, 2, PORT_TIMES_DA, ia, PORT_TIMES_DB, ib);
                value += av * bv;
                port_store(&stack[top++], vid, 2, 1.0, vid, 1.0, tid);
            }
            C_out[i + n * j] = value;
            cid[i + n * j] = vid;
        }
    }
    const int sid = next_id++;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            port_store(&stack[top++], sid, 2, 1.0, sid, 1.0, cid[i * n + j]);
    *entries = top;
    if (grad) {
        double* g = (double*) calloc((size_t) next_id + 1, sizeof (double));
        port_sweep(stack, (int) top, g);
        memcpy(grad, g, sizeof (double) * 2 * (size_t) n * n);
        free(g);
    }
    free(cid);
    free(stack);
    return 0;
}

/* single-operator probe: family 0 global, 1 pad, 2 private (mirrors ref_op_probe) */
int PSYM(op_probe)(int family, const char* op, double a, double b, int ida, int idb,
        int recording, double* value, int* id, struct ad_entry* entry, double* lhs_after) {
    struct ad_entry* stack = (struct ad_entry*) calloc(8, sizeof (struct ad_entry));
    struct ad_gradient_structure g;
    memset(&g, 0, sizeof g);
    g.gradient_stack = stack;
    g.current_ad_variable_id = 100;
    g.stack_current = 2;
    g.recording = recording;
    g.counter = 1;
    struct ad_gradient_structure p;
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
    int ok = -1;
#define TRY3(nm, ARGS_G, ARGS_P, ARGS_Q)                                           \
    if (!strcmp(op, #nm)) {                                                          \
        ok = 0;                                                                      \
        if (family == 0) r = ad_##nm ARGS_G;                                         \
        else if (family == 1) r = pad_##nm ARGS_P;                                   \
        else r = ad_##nm##_p ARGS_Q;                                                 \
    }
#define TRY_BIN(nm)                                                                 \
    TRY3(nm, (&g, va, vb), (&p, va, vb), (q, va, vb))                                \
    TRY3(nm##_vd, (&g, va, b), (&p, va, b), (q, va, b))                              \
    TRY3(nm##_dv, (&g, a, vb), (&p, a, vb), (q, a, vb))
    TRY_BIN(plus) TRY_BIN(minus) TRY_BIN(times) TRY_BIN(divide)
#define TRY2(nm, ARGS_G, ARGS_Q)                                                    \
    if (!strcmp(op, #nm)) {                                                          \
        if (family == 0) { ok = 0; r = ad_##nm ARGS_G; }                             \
        else if (family == 2) { ok = 0; r = ad_##nm##_p ARGS_Q; }                    \
    }
#define TRY_M(nm) TRY2(nm, (&g, va), (q, va))
    TRY_M(cos) TRY_M(sin) TRY_M(tan) TRY_M(acos) TRY_M(asin) TRY_M(atan) TRY_M(cosh)
    TRY_M(sinh) TRY_M(tanh) TRY_M(exp) TRY_M(log) TRY_M(log10) TRY_M(sqrt)
    TRY2(pow, (&g, va, vb), (q, va, vb))
    TRY2(pow_vd, (&g, va, b), (q, va, b))
    TRY2(pow_dv, (&g, a, vb), (q, a, vb))
    if (!strcmp(op, "plus_eq")) {
        ok = 0;
        if (family == 0) ad_plus_eq(&g, &va, vb);
        else if (family == 1) pad_plus_eq(&p, &va, vb);
        else ad_plus_eq_p(q, &va, vb);
        r = va;
    }
    if (!strcmp(op, "plus_eq_d")) {
        ok = 0;
        if (family == 0) ad_plus_eq_d(&g, &va, b);
        else if (family == 1) pad_plus_eq_d(&p, &va, b);
        else ad_plus_eq_d_p(q, &va, b);
        r = va;
    }
    if (ok == 0) {
        *value = r.value;
        *id = r.id;
        *lhs_after = va.value;
        *entry = (family == 2) ? q->gradient_stack[3] : stack[3];
    }
    free(q);
    free(stack);
    return ok;
}

int PSYM(sizeof_entry)(void) { return (int) sizeof (struct ad_entry); }
