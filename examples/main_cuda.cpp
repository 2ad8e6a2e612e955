/*
 * main_cuda.cpp -- the reference's toy driver main.cpp (data main.cpp:267-274, parameters
 * :290-291, 37 evaluations each nudging a and b by 1e-7, :328, :448-449; prints f, df/da,
 * df/db, :426-428) with the OpenCL plumbing and the host-side sum/sweep over the read-back
 * stack replaced by libad4cl_cuda.  The host tape (include/ad4cl.h) still records the
 * scalar epilogue f = N/2 * log(sum) exactly as main.cpp:417 does, and compute_gradient
 * still delivers g[a.id], g[b.id].
 *
 *   g++ -O2 -Iinclude examples/main_cuda.cpp -o main_cuda ad4cl_b200/libad4cl_cuda.so
 *   ./main_cuda [N] [iterations]
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ad4cl_cuda.h"
#include "ad4cl.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        if ((call) != AD4CL_OK) {                                                     \
            std::fprintf(stderr, "%s\n  -> %s\n", #call, ad4cl_cuda_last_error());    \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char** argv) {
    const int DATA_SIZE = argc > 1 ? std::atoi(argv[1]) : 1000000;
    const int iterations = argc > 2 ? std::atoi(argv[2]) : 37;
    std::vector<double> x(DATA_SIZE), y(DATA_SIZE);
    const double aa = 4.1919, bb = 3.2123;
    for (int i = 0; i < DATA_SIZE; i++) {
        x[i] = static_cast<double> (i);
        y[i] = aa * x[i] + bb;
    }

    struct ad_gradient_structure* gs = create_gradient_structure(1024);   // host tape: P + 3 entries per evaluation
    struct ad_variable params[2];
    ad_init_var(gs, &params[0], aa - .005);
    ad_init_var(gs, &params[1], bb - .0051);
    const int first_free_id = gs->current_variable_id;

    ad4cl_context* ctx;
    ad4cl_kernel* kernel;
    ad4cl_buffer *x_d, *y_d;
    CHECK(ad4cl_cuda_init(0, &ctx));
    CHECK(ad4cl_cuda_buffer_create(ctx, sizeof (double) * DATA_SIZE, &x_d));
    CHECK(ad4cl_cuda_buffer_create(ctx, sizeof (double) * DATA_SIZE, &y_d));
    CHECK(ad4cl_cuda_buffer_upload(x_d, 0, x.data(), sizeof (double) * DATA_SIZE));
    CHECK(ad4cl_cuda_buffer_upload(y_d, 0, y.data(), sizeof (double) * DATA_SIZE));
    CHECK(ad4cl_cuda_kernel_builtin(ctx, "linreg2_g", 0, &kernel));        // kernel.cl's AD, global family
    CHECK(ad4cl_cuda_kernel_set_arg_params(kernel, 2, params[0].id, 1));
    CHECK(ad4cl_cuda_kernel_set_arg_params(kernel, 3, params[1].id, 1));
    CHECK(ad4cl_cuda_kernel_set_arg_buffer(kernel, 4, x_d));
    CHECK(ad4cl_cuda_kernel_set_arg_buffer(kernel, 5, y_d));
    CHECK(ad4cl_cuda_kernel_set_arg_int(kernel, 7, DATA_SIZE));
    ad4cl_kernel_config cfg;
    CHECK(ad4cl_cuda_kernel_config_default(&cfg));
    cfg.global_size[0] = DATA_SIZE;
    cfg.max_ops = 4;
    cfg.epilogue = AD4CL_EPILOGUE_SUM;     // the (N/2) log epilogue is recorded on the host tape below
    CHECK(ad4cl_cuda_kernel_configure(kernel, &cfg));

    for (int iter = 0; iter < iterations; iter++) {
        struct ad_variable sum;
        CHECK(ad4cl_cuda_eval_to_tape(kernel, gs, params, 2, &sum));        // main.cpp:346-414 in one call
        struct ad_variable f = ad_times_dv(gs, static_cast<double> (DATA_SIZE) / 2.0, ad_log(gs, sum));  // :417
        int gsize = 0;
        double* g = compute_gradient(*gs, gsize);                           // :421
        std::printf("f  = %.10f\n", f.value);
        std::printf("%.10f, df/da = %.10f\n", params[0].value, g[params[0].id]);
        std::printf("%.10f, df/db = %.10f\n", params[1].value, g[params[1].id]);
        free(g);                                                            // :429
        params[0].value += .0000001;                                        // :448-449
        params[1].value += .0000001;
        gs->stack_current = 0;                                              // :451-452
        gs->current_variable_id = first_free_id;
    }
    ad4cl_cuda_kernel_free(kernel);
    ad4cl_cuda_buffer_free(x_d);
    ad4cl_cuda_buffer_free(y_d);
    ad4cl_cuda_shutdown(ctx);
    free(gs->gradient_stack);
    free(gs);
    return 0;
}
