/*
 * var_kernels.cl -- objectives written with ad4cl::var operators (CUDA side only; the
 * function-call forms of the same objectives, which also compile against the reference's ad.cl,
 * are in objectives.cl: linreg2_g and regression_logistic).  The parity test requires the two
 * forms to produce identical bits.
 */
__kernel void linreg2_var(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* a_,
        __global struct ad_variable* b_,
        __global double* x,
        __global double* y,
        __global struct ad_variable* out, int size) {
    const int id = get_global_id(0);
    if (id < size) {
        const ad4cl::var a(gs, *a_), b(gs, *b_);
        const ad4cl::var r = a * x[id] + b - y[id];
        out[id] = r * r;
    }
}

__kernel void regression_logistic_var(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    const int id = get_global_id(0);
    if (id < size) {
        ad4cl::var eta = ad4cl::var(gs, theta[0]) * d0[id];
        for (int j = 1; j < i0; j++) eta = eta + ad4cl::var(gs, theta[j]) * d0[(size_t) j * size + id];
        out[id] = log(exp(eta) + 1.0) - eta * d1[id];
    }
}
