/*
 * ad4cl_reduce.cuh -- second, fixed-order stage of the gradient reduction.
 * Included by exactly one translation unit (shim.cu).
 */
#pragma once

#include "ad4cl_tape.cuh"

/*
 * Second stage: result[c] = sum over blocks (in block order) of partials[b][c].
 * One block per column; a fixed strided pass then a fixed tree, so the value
 * does not depend on scheduling.  Optional scalar epilogue of the reference
 * (simple.cpp:365, main.cpp:417):  f = (n/2) log S, grad *= (n/2)/S.
 *   result layout: [grad[0..P-1], f, total ops, total slots]
 */
extern "C" __global__ void __launch_bounds__(256) ad4cl_reduce_partials(
        const double* __restrict__ partials, int nblocks, int ncols, int nparams,
        const double* __restrict__ global_grad, int epilogue, double n_obs,
        double* __restrict__ result) {
    __shared__ double scratch[2][8];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double v = 0.0, s = 0.0;
    for (int b = tid; b < nblocks; b += 256) {
        v += partials[(size_t) b * ncols + c];
        s += partials[(size_t) b * ncols + nparams];
    }
    v = ad4cl::warp_sum(v);
    s = ad4cl::warp_sum(s);
    if (lane == 0) {
        scratch[0][warp] = v;
        scratch[1][warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double tv = 0.0, ts = 0.0;
        for (int w = 0; w < 8; w++) {
            tv += scratch[0][w];
            ts += scratch[1][w];
        }
        if (c < nparams && global_grad != nullptr) tv = global_grad[c];
        if (epilogue == 1) {
            if (c < nparams) tv *= (n_obs / 2.0) * (1.0 / ts);
            else if (c == nparams) tv = (n_obs / 2.0) * log(ts);
        }
        result[c] = tv;
    }
}

