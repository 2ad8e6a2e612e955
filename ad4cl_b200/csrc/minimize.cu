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
