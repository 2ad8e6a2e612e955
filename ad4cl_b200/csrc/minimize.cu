/*
 * minimize.cu -- L-BFGS driver over ad4cl_cuda_eval.
 *
 * The reference declares an L-BFGS for its device runtime -- struct lbfgs_parameters_{g,p}
 * (ad.cl:99-123: linesearch, converged, max_iterations, linesearch_iteration,
 * max_linesearch_iterations, gradient_size, number_of_parameters) and lbfgs_update_{g,p}
 * (ad.cl:125-127) -- and never defines it; its examples are driven by ADMB's
 * function_minimizer instead (test/admb/simple.cpp:481-492).  This is that missing caller:
 * the two-loop recursion and a backtracking (Armijo) line search on P doubles run on the host,
 * every objective+gradient evaluation is one ad4cl_cuda_eval (the whole of it on the GPU).
 * Field names follow the reference's struct.
 */
#include "../../include/ad4cl_cuda.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

extern "C" int ad4cl_cuda_set_error_(int code, const char* msg);

namespace {

double dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}

double norm_inf(const std::vector<double>& a) {
    double m = 0.0;
    for (double v : a) m = std::fmax(m, std::fabs(v));
    return m;
}

} // namespace

extern "C" {
#pragma GCC visibility push(default)

int ad4cl_cuda_lbfgs_options_default(ad4cl_lbfgs_options* o) {
    if (!o) return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "options is NULL");
    o->max_iterations = 200;
    o->max_linesearch_iterations = 30;
    o->history = 8;
    o->gradient_tolerance = 1e-8;
    o->function_tolerance = 1e-14;
    return AD4CL_OK;
}

int ad4cl_cuda_minimize_lbfgs(ad4cl_kernel* kernel, double* theta, int number_of_parameters,
        const ad4cl_lbfgs_options* options, ad4cl_lbfgs_result* result) {
    if (!kernel || !theta || !result || number_of_parameters < 1)
        return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "NULL argument");
    ad4cl_lbfgs_options opt;
    if (options) opt = *options;
    else ad4cl_cuda_lbfgs_options_default(&opt);
    const int P = number_of_parameters, m = opt.history > 0 ? opt.history : 8;
    std::vector<double> x(theta, theta + P), g(P), xn(P), gn(P), d(P), q(P);
    std::vector<std::vector<double>> S, Y;
    std::vector<double> rho;
    memset(result, 0, sizeof *result);

    double f = 0.0;
    int rc = ad4cl_cuda_eval(kernel, x.data(), P, &f, g.data());
    if (rc != AD4CL_OK) return rc;
    result->evaluations = 1;
    for (int it = 0; it < opt.max_iterations; it++) {
        const double gmax = norm_inf(g);
        if (gmax <= opt.gradient_tolerance) {
            result->converged = 1;
            break;
        }
        /* two-loop recursion: d = -H g */
        q = g;
        const int hist = (int) S.size();
        std::vector<double> alpha(hist);
        for (int i = hist - 1; i >= 0; i--) {
            alpha[i] = rho[i] * dot(S[i], q);
            for (int j = 0; j < P; j++) q[j] -= alpha[i] * Y[i][j];
        }
        double gamma = 1.0;
        if (hist > 0) gamma = dot(S[hist - 1], Y[hist - 1]) / dot(Y[hist - 1], Y[hist - 1]);
        for (int j = 0; j < P; j++) q[j] *= gamma;
        for (int i = 0; i < hist; i++) {
            const double beta = rho[i] * dot(Y[i], q);
            for (int j = 0; j < P; j++) q[j] += (alpha[i] - beta) * S[i][j];
        }
        for (int j = 0; j < P; j++) d[j] = -q[j];
        double slope = dot(g, d);
        if (!(slope < 0.0)) { /* not a descent direction: restart from steepest descent */
            S.clear(); Y.clear(); rho.clear();
            for (int j = 0; j < P; j++) d[j] = -g[j];
            slope = -dot(g, g);
        }
        /* backtracking line search, sufficient decrease c1 = 1e-4 */
        double step = hist == 0 ? std::fmin(1.0, 1.0 / std::sqrt(-slope)) : 1.0;
        double fn = f;
        bool ok = false;
        int ls = 0;
        for (; ls < opt.max_linesearch_iterations; ls++) {
            for (int j = 0; j < P; j++) xn[j] = x[j] + step * d[j];
            rc = ad4cl_cuda_eval(kernel, xn.data(), P, &fn, gn.data());
            if (rc != AD4CL_OK) return rc;
            result->evaluations++;
            if (std::isfinite(fn) && fn <= f + 1e-4 * step * slope) {
                ok = true;
                break;
            }
            step *= 0.5;
        }
        result->linesearch_iteration += ls + 1;
        if (!ok) {
            result->converged = -1; /* line search failed: keep the last accepted point */
            break;
        }
        std::vector<double> s(P), y(P);
        for (int j = 0; j < P; j++) {
            s[j] = xn[j] - x[j];
            y[j] = gn[j] - g[j];
        }
        const double sy = dot(s, y);
        if (sy > 1e-12 * std::sqrt(dot(s, s) * dot(y, y))) {
            if ((int) S.size() == m) {
                S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin());
            }
            S.push_back(s); Y.push_back(y); rho.push_back(1.0 / sy);
        }
        const double df = f - fn;
        x = xn;
        g = gn;
        f = fn;
        result->iterations = it + 1;
        if (df <= opt.function_tolerance * std::fmax(1.0, std::fabs(f))) {
            result->converged = 2;
            break;
        }
    }
    memcpy(theta, x.data(), sizeof (double) * (size_t) P);
    result->f = f;
    result->gradient_max = norm_inf(g);
    result->number_of_parameters = P;
    return AD4CL_OK;
}

#pragma GCC visibility pop
} // extern "C"

/* ======================================================================================================
 * Device-resident L-BFGS (ad4cl_cuda_minimize_lbfgs_device): the reference's lbfgs_parameters_g /
 * lbfgs_update_g (ad.cl:99-127), defined.  One single-CTA kernel, `lbfgs_update`, runs between two
 * evaluations and is a state machine over the SAME algorithm as the host driver above (two-loop
 * recursion, Armijo backtracking with c1 = 1e-4, curvature-guarded history update, steepest-descent
 * restart); every dot product is a fixed-order block reduction, so a fit is bit-reproducible.
 * ====================================================================================================== */
namespace {

/* field names follow ad.cl:99-123 where the reference has them */
struct lbfgs_parameters_g {
    int phase;                      /* 0: the evaluation of the start point is pending, 1: a line-search trial is pending */
    int done;                       /* != 0: finished; the evaluations still queued skip themselves on this flag */
    int converged;                  /* ad.cl:105 */
    int max_iterations;             /* ad.cl:106 */
    int linesearch_iteration;       /* ad.cl:107: trials in total */
    int max_linesearch_iterations;  /* ad.cl:108 */
    int number_of_parameters;       /* ad.cl:110 */
    int iterations, evaluations, ls, hist, head, m;
    double f, step, slope, gradient_max;
    double gradient_tolerance, function_tolerance;
    double *x, *g, *d, *q, *theta, *res;   /* theta: the trial point the next evaluation reads; res: [grad, f, ops, slots] */
    double *S, *Y, *rho, *alpha;           /* history ring: logical pair i is physical (head + i) % m */
};

constexpr int kLbfgsThreads = 256;

__device__ double block_dot(const double* a, const double* b, int n, double* scratch) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kLbfgsThreads) s += a[i] * b[i];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kLbfgsThreads / 32; w++) t += scratch[w];
    return t;
}

__device__ double block_max_abs(const double* a, int n, double* scratch) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kLbfgsThreads) s = fmax(s, fabs(a[i]));
    for (int off = 16; off > 0; off >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kLbfgsThreads / 32; w++) t = fmax(t, scratch[w]);
    return t;
}

/* d = -H g by the two-loop recursion; returns the slope g.d (falls back to steepest descent) */
__device__ double new_direction(lbfgs_parameters_g* p, double* scratch) {
    const int P = p->number_of_parameters, m = p->m;
    for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->q[j] = p->g[j];
    __syncthreads();
    for (int i = p->hist - 1; i >= 0; i--) {
        const int k = (p->head + i) % m;
        const double a = p->rho[k] * block_dot(p->S + (size_t) k * P, p->q, P, scratch);
        if (threadIdx.x == 0) p->alpha[k] = a;
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->q[j] -= a * p->Y[(size_t) k * P + j];
        __syncthreads();
    }
    double gamma = 1.0;
    if (p->hist > 0) {
        const int k = (p->head + p->hist - 1) % m;
        gamma = block_dot(p->S + (size_t) k * P, p->Y + (size_t) k * P, P, scratch)
                / block_dot(p->Y + (size_t) k * P, p->Y + (size_t) k * P, P, scratch);
    }
    for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->q[j] *= gamma;
    __syncthreads();
    for (int i = 0; i < p->hist; i++) {
        const int k = (p->head + i) % m;
        const double beta = p->rho[k] * block_dot(p->Y + (size_t) k * P, p->q, P, scratch);
        const double a = p->alpha[k];
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->q[j] += (a - beta) * p->S[(size_t) k * P + j];
        __syncthreads();
    }
    for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->d[j] = -p->q[j];
    __syncthreads();
    double slope = block_dot(p->g, p->d, P, scratch);
    if (!(slope < 0.0)) { /* not a descent direction: restart from steepest descent */
        __syncthreads();
        if (threadIdx.x == 0) p->hist = 0;
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->d[j] = -p->g[j];
        __syncthreads();
        slope = -block_dot(p->g, p->g, P, scratch);
    }
    return slope;
}

/* start the line search along d from x: first trial point into theta */
__device__ void start_linesearch(lbfgs_parameters_g* p, double slope) {
    const int P = p->number_of_parameters;
    const double step = p->hist == 0 ? fmin(1.0, 1.0 / sqrt(-slope)) : 1.0;
    for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->theta[j] = p->x[j] + step * p->d[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        p->step = step;
        p->slope = slope;
        p->ls = 0;
        p->phase = 1;
    }
}

__device__ void finish(lbfgs_parameters_g* p, int converged, double* scratch) {
    const int P = p->number_of_parameters;
    const double gmax = block_max_abs(p->g, P, scratch);
    for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->theta[j] = p->x[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        p->converged = converged;
        p->gradient_max = gmax;
        p->done = 1;
    }
}

/* ad.cl:125: `void lbfgs_update_g(...)` -- consumes the evaluation that just finished */
__global__ void __launch_bounds__(kLbfgsThreads) lbfgs_update(lbfgs_parameters_g* p) {
    __shared__ double scratch[kLbfgsThreads / 32];
    if (p->done) return;
    const int P = p->number_of_parameters;
    const double fn = p->res[P];
    const double* gn = p->res;
    __syncthreads();
    if (threadIdx.x == 0) p->evaluations++;
    if (p->phase == 0) { /* the start point */
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) {
            p->x[j] = p->theta[j];
            p->g[j] = gn[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) p->f = fn;
        __syncthreads();
    } else {
        const bool accept = isfinite(fn) && fn <= p->f + 1e-4 * p->step * p->slope;
        __syncthreads();
        if (threadIdx.x == 0) p->linesearch_iteration++;
        if (!accept) {
            if (p->ls + 1 >= p->max_linesearch_iterations) { /* line search failed: keep the last accepted point */
                finish(p, -1, scratch);
                return;
            }
            const double step = p->step * 0.5;
            for (int j = threadIdx.x; j < P; j += kLbfgsThreads) p->theta[j] = p->x[j] + step * p->d[j];
            __syncthreads();
            if (threadIdx.x == 0) {
                p->step = step;
                p->ls++;
            }
            return;
        }
        /* accepted: s = theta - x, y = gn - g, first into the scratch vectors d and q (both are recomputed
           before their next use); the pair enters the history -- after the newest pair, or over the oldest
           when the ring is full -- only if it passes the curvature test */
        const int m = p->m;
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) {
            p->d[j] = p->theta[j] - p->x[j];
            p->q[j] = gn[j] - p->g[j];
        }
        __syncthreads();
        const double sy = block_dot(p->d, p->q, P, scratch);
        const double ss = block_dot(p->d, p->d, P, scratch);
        const double yy = block_dot(p->q, p->q, P, scratch);
        const double df = p->f - fn;
        const bool keep = sy > 1e-12 * sqrt(ss * yy);
        const int k = p->hist == m ? p->head : (p->head + p->hist) % m;
        __syncthreads();
        for (int j = threadIdx.x; j < P; j += kLbfgsThreads) {
            if (keep) {
                p->S[(size_t) k * P + j] = p->d[j];
                p->Y[(size_t) k * P + j] = p->q[j];
            }
            p->x[j] = p->theta[j];
            p->g[j] = gn[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (keep) {
                p->rho[k] = 1.0 / sy;
                if (p->hist == m) p->head = (p->head + 1) % m;
                else p->hist++;
            }
            p->f = fn;
            p->iterations++;
        }
        __syncthreads();
        if (df <= p->function_tolerance * fmax(1.0, fabs(fn))) {
            finish(p, 2, scratch);
            return;
        }
        if (p->iterations >= p->max_iterations) {
            finish(p, 0, scratch);
            return;
        }
    }
    if (block_max_abs(p->g, P, scratch) <= p->gradient_tolerance) {
        finish(p, 1, scratch);
        return;
    }
    const double slope = new_direction(p, scratch);
    __syncthreads();
    start_linesearch(p, slope);
}

#define LB_TRY(expr)                                                                                 \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            rc = ad4cl_cuda_set_error_(AD4CL_ERR_CUDA, cudaGetErrorString(e_));                      \
            goto out;                                                                                \
        }                                                                                            \
    } while (0)

} // namespace

extern "C" {
#pragma GCC visibility push(default)

int ad4cl_cuda_minimize_lbfgs_device(ad4cl_kernel* kernel, double* theta, int number_of_parameters,
        const ad4cl_lbfgs_options* options, ad4cl_lbfgs_result* result) {
    if (!kernel || !theta || !result || number_of_parameters < 1)
        return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "NULL argument");
    ad4cl_lbfgs_options opt;
    if (options) opt = *options;
    else ad4cl_cuda_lbfgs_options_default(&opt);
    const int P = number_of_parameters, m = opt.history > 0 ? opt.history : 8;
    constexpr int kPairsPerGraph = 128;
    memset(result, 0, sizeof *result);
    ad4cl_context* ctx = ad4cl_cuda_kernel_context(kernel);
    cudaStream_t st = (cudaStream_t) ad4cl_cuda_stream(ctx);
    int rc = AD4CL_OK;
    double* pool = nullptr;
    lbfgs_parameters_g* state = nullptr;
    lbfgs_parameters_g h{};
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const size_t doubles = (size_t) P * 5 + (size_t) (P + 4) + 2 * (size_t) m * P + 2 * (size_t) m;

    LB_TRY(cudaMalloc(&pool, sizeof (double) * doubles));
    LB_TRY(cudaMalloc(&state, sizeof (lbfgs_parameters_g)));
    LB_TRY(cudaMemsetAsync(pool, 0, sizeof (double) * doubles, st));
    {
        double* c = pool;
        h.x = c; c += P;
        h.g = c; c += P;
        h.d = c; c += P;
        h.q = c; c += P;
        h.theta = c; c += P;
        h.res = c; c += P + 4;
        h.S = c; c += (size_t) m * P;
        h.Y = c; c += (size_t) m * P;
        h.rho = c; c += m;
        h.alpha = c;
    }
    h.max_iterations = opt.max_iterations;
    h.max_linesearch_iterations = opt.max_linesearch_iterations;
    h.number_of_parameters = P;
    h.m = m;
    h.gradient_tolerance = opt.gradient_tolerance;
    h.function_tolerance = opt.function_tolerance;
    LB_TRY(cudaMemcpyAsync(state, &h, sizeof h, cudaMemcpyHostToDevice, st));
    LB_TRY(cudaMemcpyAsync(h.theta, theta, sizeof (double) * (size_t) P, cudaMemcpyHostToDevice, st));
    LB_TRY(cudaStreamSynchronize(st));

    /* one uncaptured evaluation first: it sizes the kernel's resources (allocations cannot be captured) */
    rc = ad4cl_cuda_eval_device_if(kernel, h.theta, P, h.res, 1, 1, &state->done);
    if (rc != AD4CL_OK) goto out;
    lbfgs_update<<<1, kLbfgsThreads, 0, st>>>(state);
    rc = ad4cl_cuda_check(kernel);
    if (rc != AD4CL_OK) goto out;

    LB_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < kPairsPerGraph && rc == AD4CL_OK; i++) {
        rc = ad4cl_cuda_eval_device_if(kernel, h.theta, P, h.res, 1, 1, &state->done);
        lbfgs_update<<<1, kLbfgsThreads, 0, st>>>(state);
    }
    {
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc != AD4CL_OK) goto out;
        LB_TRY(e);
    }
    LB_TRY(cudaGraphInstantiate(&exec, graph, 0));
    for (int launches = 0; launches < 64; launches++) {
        LB_TRY(cudaGraphLaunch(exec, st));
        LB_TRY(cudaMemcpyAsync(&h, state, sizeof h, cudaMemcpyDeviceToHost, st));
        LB_TRY(cudaStreamSynchronize(st));
        if (h.done) break;
    }
    rc = ad4cl_cuda_check(kernel);
    if (rc != AD4CL_OK) goto out;
    /* not done after 64 x 128 evaluations: report the last accepted point */
    LB_TRY(cudaMemcpyAsync(theta, h.done ? h.theta : h.x, sizeof (double) * (size_t) P, cudaMemcpyDeviceToHost, st));
    LB_TRY(cudaStreamSynchronize(st));
    result->f = h.f;
    result->gradient_max = h.gradient_max;
    result->converged = h.done ? h.converged : 0;
    result->iterations = h.iterations;
    result->linesearch_iteration = h.linesearch_iteration;
    result->evaluations = h.evaluations;
    result->number_of_parameters = P;
out:
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    cudaFree(pool);
    cudaFree(state);
    return rc;
}

#pragma GCC visibility pop
} // extern "C"
