/*
 * oracle/ref_host.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * The host half of the reference oracle: #includes the reference's own
 * ad4cl.h (REF_AD4CL_H, by absolute path, g++ -fpermissive -w) and drives the
 * per-evaluation protocol of test/admb/simple.cpp:270-384 around the device
 * kernels compiled in ref_dev.c:
 *     write gs (24 bytes) -> NDRange -> read gs -> gpu_restore
 *     -> host sum of out[] with ad_plus_eq_v -> scalar epilogue
 *     -> compute_gradient (ad4cl.h:775-802) -> read g[param.id].
 * Only the OpenCL transport is replaced (a memcpy of the same 24 bytes).
 */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include REF_AD4CL_H

typedef struct ad_variable var_t;
typedef struct ad_entry entry_t;
typedef struct ad_gradient_structure host_gs_t; /* 24 bytes (ad4cl.h:55-61) */

extern "C" {
/* device half (ref_dev.c) */
int ref_sizeof_dev_gs(void);
void ref_launch_simple(int variant, void* gs, entry_t* stack, var_t* a, var_t* b,
        double* x, double* y, var_t* out, int size, int global_size, int threads);
void ref_launch_matrix(int variant, void* gs, entry_t* stack, var_t* A, var_t* B, var_t* C,
        int widthA, int widthB, int g0, int g1);
int ref_launch_objective(const char* name, void* gs, entry_t* stack, var_t* theta,
        double* d0, double* d1, var_t* out, int size, int i0, int i1, int global_size, int threads);
}

namespace {

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* "enqueueWriteBuffer(gs_d, ..., sizeof(ad_gradient_structure), gs)": the host
   structure is a 24-byte prefix of the 32-byte device one. */
struct dev_gs_image {
    alignas(8) unsigned char bytes[64];
    void write(const host_gs_t* gs) { memset(bytes, 0, sizeof bytes); memcpy(bytes, gs, sizeof (host_gs_t)); }
    void read(host_gs_t* gs) const { memcpy(gs, bytes, sizeof (host_gs_t)); }
};

/* host sum + epilogue + sweep: simple.cpp:357-369 */
void finish(host_gs_t* gs, var_t* out, int n, int epilogue, int nparams,
        double* f, double* grad, long long* entries, double* t_ms) {
    double t0 = now_ms();
    var_t sum = {.value = 0.0, .id = gs->current_variable_id++};
    for (int i = 0; i < n; i++) {
        ad_plus_eq_v(gs, &sum, out[i]);
    }
    var_t ff = sum;
    if (epilogue == 1) {
        ff = ad_times_dv(gs, static_cast<double> (n) / 2.0, ad_log(gs, sum));
    }
    double t1 = now_ms();
    int gsize = 0;
    double* g = compute_gradient(*gs, gsize);
    double t2 = now_ms();
    *f = ff.value;
    for (int p = 0; p < nparams; p++) {
        grad[p] = g[p];
    }
    *entries = gs->stack_current;
    free(g);
    if (t_ms) {
        t_ms[1] = t1 - t0;
        t_ms[2] = t2 - t1;
    }
}

} // namespace

extern "C" {

int ref_sizeof_host_gs(void) { return (int) sizeof (host_gs_t); }

/*
 * The host half of the protocol on a stack that was recorded ELSEWHERE -- by the reference's
 * unmodified OpenCL kernels on a real OpenCL device (oracle/opencl_ref.py) and read back with
 * enqueueReadBuffer, exactly as simple.cpp:312-317 does: gpu_restore, host sum of out[], epilogue,
 * compute_gradient.  `stack` must have room for counter + n + 2 entries.
 */
int ref_finish_stack(entry_t* stack, int counter, int nparams, var_t* out, int n, int epilogue,
        double* f, double* grad, long long* entries, double* t_ms) {
    host_gs_t gs_s;
    host_gs_t* gs = &gs_s;
    gs->gradient_stack = stack;
    gs->current_variable_id = nparams;
    gs->stack_current = 0;
    gs->recording = 1;
    gs->counter = counter;
    gpu_restore(gs);
    finish(gs, out, n, epilogue, nparams, f, grad, entries, t_ms);
    return 0;
}

/*
 * Config 1 through the reference's own kernels.
 *   variant 0: kernel.cl   1: test/admb/simple.cl   2: host twin AD()
 *   (simple.cpp:14-28, host ops of ad4cl.h -- the AD4CL_HOST method)
 * f = (n/2) log sum_i ((a x_i + b) - y_i)^2 ; grad = {df/da, df/db}.
 */
int ref_eval_simple(int variant, double a, double b, double* x, double* y, int n,
        int global_size, int threads, double* f, double* grad, long long* entries, double* t_ms) {
    const long long cap = 5LL * n + 16;
    host_gs_t gs_s;
    host_gs_t* gs = &gs_s;
    gs->current_variable_id = 0;
    gs->stack_current = 0;
    gs->recording = 1;
    gs->counter = 0;
    entry_t* stack = create_entries((int) cap);
    gs->gradient_stack = stack;
    var_t aa, bb;
    ad_init_var(gs, &aa, a);
    ad_init_var(gs, &bb, b);
    std::vector<var_t> out(n > 0 ? n : 1);

    double t0 = now_ms();
    if (variant == 2) {
        for (int i = 0; i < n; i++) {
            var_t temp = ad_minus_vd(gs, ad_plus(gs, ad_times_vd(gs, aa, x[i]), bb), y[i]);
            out[i] = ad_times(gs, temp, temp);
        }
    } else {
        dev_gs_image img;
        img.write(gs);
        ref_launch_simple(variant, img.bytes, stack, &aa, &bb, x, y, out.data(), n, global_size, threads);
        img.read(gs);
        gs->gradient_stack = stack;
        gpu_restore(gs);
    }
    if (t_ms) t_ms[0] = now_ms() - t0;
    finish(gs, out.data(), n, 1, 2, f, grad, entries, t_ms);
    free(stack);
    return 0;
}

/*
 * Any kernel of ad4cl_b200/kernels/objectives.cl through the reference
 * runtime.  epilogue 0: f = sum out_i ; 1: f = (n/2) log sum out_i.
 */
int ref_eval_objective(const char* name, int nparams, const double* theta,
        double* d0, double* d1, int size, int i0, int i1, int global_size,
        int epilogue, int threads, long long stack_entries,
        double* f, double* grad, long long* entries, double* t_ms) {
    host_gs_t gs_s;
    host_gs_t* gs = &gs_s;
    gs->current_variable_id = 0;
    gs->stack_current = 0;
    gs->recording = 1;
    gs->counter = 0;
    entry_t* stack = (entry_t*) calloc((size_t) stack_entries, sizeof (entry_t));
    if (!stack) return -2;
    gs->gradient_stack = stack;
    std::vector<var_t> th(nparams);
    for (int p = 0; p < nparams; p++) {
        ad_init_var(gs, &th[p], theta[p]);
    }
    std::vector<var_t> out(size > 0 ? size : 1);

    double t0 = now_ms();
    dev_gs_image img;
    img.write(gs);
    int rc = ref_launch_objective(name, img.bytes, stack, th.data(), d0, d1, out.data(),
            size, i0, i1, global_size, threads);
    if (rc != 0) {
        free(stack);
        return rc;
    }
    img.read(gs);
    gs->gradient_stack = stack;
    gpu_restore(gs);
    if (t_ms) t_ms[0] = now_ms() - t0;
    if ((long long) gs->stack_current + size + 2 > stack_entries) {
        free(stack);
        return -3; /* the reference has no bounds check (Q8); the oracle refuses instead */
    }
    finish(gs, out.data(), size, epilogue, nparams, f, grad, entries, t_ms);
    free(stack);
    return 0;
}

/*
 * A NON-UNIFORM-SEED epilogue on the reference's own runtime: the kernel's out[] as above,
 * then the host tape records  f = sum_i w_i * log(out_i + shift)  with the reference's ad_plus_vd, ad_log,
 * ad_times_dv and ad_plus_eq_v (ad4cl.h:131, 651, 368, 186) and compute_gradient (ad4cl.h:775-802) sweeps the
 * whole stack.  d f / d out_i = w_i / (out_i + shift) differs per work-item: the case the (n/2) log sum
 * epilogue never exercises.  (ad_plus_vd has the reference's Q4 defect; use the patched build.)
 */
int ref_eval_weighted_log(const char* name, int nparams, const double* theta,
        double* d0, double* d1, int size, int i0, int i1, int global_size, const double* weights, double shift,
        long long stack_entries, double* f, double* grad, double* out_values) {
    host_gs_t gs_s;
    host_gs_t* gs = &gs_s;
    gs->current_variable_id = 0;
    gs->stack_current = 0;
    gs->recording = 1;
    gs->counter = 0;
    entry_t* stack = (entry_t*) calloc((size_t) stack_entries, sizeof (entry_t));
    if (!stack) return -2;
    gs->gradient_stack = stack;
    std::vector<var_t> th(nparams);
    for (int p = 0; p < nparams; p++) ad_init_var(gs, &th[p], theta[p]);
    std::vector<var_t> out(size > 0 ? size : 1);
    dev_gs_image img;
    img.write(gs);
    int rc = ref_launch_objective(name, img.bytes, stack, th.data(), d0, d1, out.data(), size, i0, i1, global_size, 1);
    if (rc != 0) {
        free(stack);
        return rc;
    }
    img.read(gs);
    gs->gradient_stack = stack;
    gpu_restore(gs);
    if ((long long) gs->stack_current + 4LL * size + 2 > stack_entries) {
        free(stack);
        return -3;
    }
    var_t sum = {.value = 0.0, .id = gs->current_variable_id++};
    for (int i = 0; i < size; i++) {
        if (out_values) out_values[i] = out[i].value;
        ad_plus_eq_v(gs, &sum, ad_times_dv(gs, weights[i], ad_log(gs, ad_plus_vd(gs, out[i], shift))));
    }
    int gsize = 0;
    double* g = compute_gradient(*gs, gsize);
    *f = sum.value;
    for (int p = 0; p < nparams; p++) grad[p] = g[p];
    free(g);
    free(stack);
    return 0;
}

/*
 * test/matrix: C = A*B on n x n ad_variables, then the host sum of C
 * (matrix_mul.cpp:253-261).  A gets ids 0..n^2-1, B ids n^2..2n^2-1
 * (matrix_mul.cpp:84-90).
 *   variant -1: host twin MatrixMultHost (matrix_mul.cpp:39-57)
 *            0: device GLOBAL kernel   1: device USE_LOCAL kernel
 * C_out receives the n^2 forward values, *entries the tape length printed at
 * matrix_mul.cpp:272.  grad (2 n^2 doubles, may be NULL) is d(sum C)/d(A,B);
 * it is only meaningful for variant -1 (the device kernels' ids collide, Q6).
 */
int ref_matrix(int variant, int n, const double* A, const double* B, double* C_out,
        long long* entries, double* grad, double* t_ms) {
    const long long cap = 2LL * n * n * n + 2LL * n * n + 16;
    host_gs_t gs_s;
    host_gs_t* gs = &gs_s;
    gs->current_variable_id = 0;
    gs->stack_current = 0;
    gs->recording = 1;
    gs->counter = 0;
    entry_t* stack = (entry_t*) calloc((size_t) cap, sizeof (entry_t));
    if (!stack) return -2;
    gs->gradient_stack = stack;
    std::vector<var_t> a(n * n), b(n * n), c(n * n);
    for (int i = 0; i < n * n; i++) ad_init_var(gs, &a[i], A[i]);
    for (int i = 0; i < n * n; i++) ad_init_var(gs, &b[i], B[i]);

    double t0 = now_ms();
    if (variant < 0) {
        for (int j = 0; j < n; j++) {
            for (int i = 0; i < n; i++) {
                var_t value;
                ad_init_var(gs, &value, 0.0);
                for (int k = 0; k < n; k++) {
                    ad_plus_eq_v(gs, &value, ad_times(gs, a[k + j * n], b[k * n + i]));
                }
                c[i + n * j] = value;
            }
        }
    } else {
        dev_gs_image img;
        img.write(gs);
        ref_launch_matrix(variant, img.bytes, stack, a.data(), b.data(), c.data(), n, n, n, n);
        img.read(gs);
        gs->gradient_stack = stack;
        gpu_restore(gs);
    }
    if (t_ms) t_ms[0] = now_ms() - t0;

    var_t sum;
    ad_init_var(gs, &sum, 0.0);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) {
            ad_plus_eq_v(gs, &sum, c[i * n + j]);
        }
    }
    for (int i = 0; i < n * n; i++) C_out[i] = c[i].value;
    *entries = gs->stack_current;
    if (grad) {
        int gsize = 0;
        double* g = compute_gradient(*gs, gsize);
        for (int i = 0; i < 2 * n * n; i++) grad[i] = g[i];
        free(g);
    }
    free(stack);
    return 0;
}

/*
 * Host-op probe (family "host" of the op table): one ad4cl.h
 * operator on (a, b); returns value, id and the recorded entry.
 * ad_plus_vd is probed on a zeroed stack because of Q4 (coeff[1] is read
 * uninitialised by the sweep; here it reads zeros).
 */
int ref_host_op_probe(const char* op, double a, double b, int ida, int idb, int recording,
        double* value, int* id, entry_t* entry, double* lhs_after) {
    entry_t stack[8];
    memset(stack, 0, sizeof stack);
    host_gs_t g;
    g.gradient_stack = stack;
    g.current_variable_id = 101;
    g.stack_current = 3;
    g.recording = recording;
    g.counter = 0;
    var_t va = {a, ida}, vb = {b, idb}, r = {0.0, 0};
    int ok = -1;
#define HB(nm)                                                         \
    if (!strcmp(op, #nm)) { ok = 0; r = ad_##nm(&g, va, vb); }         \
    if (!strcmp(op, #nm "_vd")) { ok = 0; r = ad_##nm##_vd(&g, va, b); } \
    if (!strcmp(op, #nm "_dv")) { ok = 0; r = ad_##nm##_dv(&g, a, vb); }
    HB(plus) HB(minus) HB(times) HB(divide)
#define HM(nm) if (!strcmp(op, #nm)) { ok = 0; r = ad_##nm(&g, va); }
    HM(cos) HM(sin) HM(tan) HM(acos) HM(asin) HM(atan) HM(cosh) HM(sinh) HM(tanh)
    HM(exp) HM(log) HM(log10) HM(sqrt)
    if (!strcmp(op, "pow")) { ok = 0; r = ad_pow(&g, va, vb); }
    if (!strcmp(op, "pow_vd")) { ok = 0; r = ad_pow_vd(&g, va, b); }
    if (!strcmp(op, "pow_dv")) { ok = 0; r = ad_pow_dv(&g, a, vb); }
    if (!strcmp(op, "plus_eq")) { ok = 0; ad_plus_eq_v(&g, &va, vb); r = va; }
    if (!strcmp(op, "plus_eq_d")) { ok = 0; ad_plus_eq_d(&g, &va, b); r = va; }
    if (ok == 0) {
        *value = r.value;
        *id = r.id;
        *entry = stack[3];
        *lhs_after = va.value;
    }
    return ok;
}

} // extern "C"
