/*
 * sharded_cuda.cpp -- a C++ host (the reference's host language) driving EVERY GPU of the box from one
 * process through the C ABI: the shipped example's objective (test/admb/simple.cl:106-141 =
 * linreg2_p; data recipe main.cpp:267-274) with the observations sharded over `ranks` ranks, one
 * evaluation = one call of ad4cl_cuda_eval_sharded per rank (host theta in -> f, grad out; the kernel's
 * last CTA all-reduces the raw sums with the peers and applies the (N/2) log epilogue).
 *
 *   g++ -O2 -pthread -Iinclude examples/sharded_cuda.cpp -o sharded_cuda ad4cl_b200/libad4cl_cuda.so
 *   ./sharded_cuda [N] [ranks] [evaluations]      ranks may exceed the number of GPUs (ranks share devices)
 *
 * Prints the sharded result of every rank and the single-GPU result of the whole problem.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ad4cl_cuda.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        if ((call) != AD4CL_OK) {                                                     \
            std::fprintf(stderr, "%s\n  -> %s\n", #call, ad4cl_cuda_last_error());    \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)

static ad4cl_kernel* make_kernel(ad4cl_context* ctx, const double* x, const double* y, int count, int n_total,
        ad4cl_buffer** x_d, ad4cl_buffer** y_d) {
    ad4cl_kernel* k;
    CHECK(ad4cl_cuda_buffer_create(ctx, sizeof (double) * count, x_d));
    CHECK(ad4cl_cuda_buffer_create(ctx, sizeof (double) * count, y_d));
    CHECK(ad4cl_cuda_buffer_upload(*x_d, 0, x, sizeof (double) * count));
    CHECK(ad4cl_cuda_buffer_upload(*y_d, 0, y, sizeof (double) * count));
    CHECK(ad4cl_cuda_kernel_builtin(ctx, "linreg2_p", 0, &k));
    CHECK(ad4cl_cuda_kernel_set_arg_params(k, 2, 0, 1));
    CHECK(ad4cl_cuda_kernel_set_arg_params(k, 3, 1, 1));
    CHECK(ad4cl_cuda_kernel_set_arg_buffer(k, 4, *x_d));
    CHECK(ad4cl_cuda_kernel_set_arg_buffer(k, 5, *y_d));
    CHECK(ad4cl_cuda_kernel_set_arg_int(k, 7, count));
    ad4cl_kernel_config cfg;
    CHECK(ad4cl_cuda_kernel_config_default(&cfg));
    cfg.global_size[0] = count;
    cfg.max_ops = 4;
    cfg.epilogue = AD4CL_EPILOGUE_HALF_N_LOG;   /* simple.cpp:365 */
    cfg.epilogue_n = n_total;                   /* n of the WHOLE problem */
    CHECK(ad4cl_cuda_kernel_configure(k, &cfg));
    return k;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? std::atoi(argv[1]) : 1000003;
    const int ranks = argc > 2 ? std::atoi(argv[2]) : 2;
    const int evals = argc > 3 ? std::atoi(argv[3]) : 3;
    std::vector<double> x(N), y(N);
    for (int i = 0; i < N; i++) {
        x[i] = static_cast<double> (i) / N;
        y[i] = 4.1919 * x[i] + 3.2123 + 0.01 * std::sin(12345.678 * i);
    }
    double theta[2] = {4.1919 - .005, 3.2123 - .0051};

    /* how many devices?  ad4cl_cuda_init fails beyond the last one */
    std::vector<ad4cl_context*> ctxs(ranks);
    int ndev = 0;
    {
        ad4cl_context* probe;
        while (ndev < ranks && ad4cl_cuda_init(ndev, &probe) == AD4CL_OK) {
            ad4cl_cuda_shutdown(probe);
            ndev++;
        }
        if (ndev == 0) {
            std::fprintf(stderr, "no CUDA device: %s\n", ad4cl_cuda_last_error());
            return 1;
        }
    }
    for (int r = 0; r < ranks; r++) CHECK(ad4cl_cuda_init(r % ndev, &ctxs[r]));
    std::vector<ad4cl_comm*> comms(ranks);
    CHECK(ad4cl_cuda_comm_create_local(ctxs.data(), ranks, 2 + 3, comms.data()));

    std::vector<ad4cl_kernel*> kernels(ranks);
    std::vector<ad4cl_buffer*> xb(ranks), yb(ranks);
    for (int r = 0; r < ranks; r++) {   /* contiguous observation shards */
        const int base = N / ranks, rem = N % ranks;
        const int start = r * base + (r < rem ? r : rem), count = base + (r < rem ? 1 : 0);
        kernels[r] = make_kernel(ctxs[r], x.data() + start, y.data() + start, count, N, &xb[r], &yb[r]);
    }
    std::vector<double> f(ranks), g(2 * ranks);
    for (int e = 0; e < evals; e++) {
        std::vector<std::thread> th;
        for (int r = 0; r < ranks; r++)   /* one host thread per rank: the call blocks until the peers have arrived */
            th.emplace_back([&, r] { CHECK(ad4cl_cuda_eval_sharded(kernels[r], comms[r], theta, 2, &f[r], &g[2 * r])); });
        for (auto& t : th) t.join();
        theta[0] += 1e-7;
        theta[1] += 1e-7;
    }
    theta[0] -= 1e-7;
    theta[1] -= 1e-7;
    for (int r = 0; r < ranks; r++)
        std::printf("rank %d of %d on device %d: f = %.17g  df/da = %.17g  df/db = %.17g  launches/eval = %d\n", r, ranks,
                r % ndev, f[r], g[2 * r], g[2 * r + 1], ad4cl_cuda_launches_per_eval(kernels[r]));

    ad4cl_buffer *x1, *y1;
    ad4cl_kernel* single = make_kernel(ctxs[0], x.data(), y.data(), N, N, &x1, &y1);
    double f1, g1[2];
    CHECK(ad4cl_cuda_eval(single, theta, 2, &f1, g1));
    std::printf("single GPU: f = %.17g  df/da = %.17g  df/db = %.17g\n", f1, g1[0], g1[1]);

    ad4cl_cuda_kernel_free(single);
    ad4cl_cuda_buffer_free(x1);
    ad4cl_cuda_buffer_free(y1);
    for (int r = 0; r < ranks; r++) {
        ad4cl_cuda_kernel_free(kernels[r]);
        ad4cl_cuda_buffer_free(xb[r]);
        ad4cl_cuda_buffer_free(yb[r]);
        ad4cl_cuda_comm_destroy(comms[r]);
    }
    for (int r = 0; r < ranks; r++) ad4cl_cuda_shutdown(ctxs[r]);
    return 0;
}
