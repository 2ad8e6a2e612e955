/*
 * ad4cl_reduce.cuh -- second, fixed-order stage of the gradient reduction.
 * Included by exactly one translation unit (shim.cu).
 */
#pragma once

#include "ad4cl_tape.cuh"

/*
 * Second stage: result[first_col + c] = sum over blocks (in block order) of partials[b][c].
 * One block per column; a fixed strided pass then a fixed tree, so the value
 * does not depend on scheduling.  Optional scalar epilogue of the reference
 * (simple.cpp:365, main.cpp:417):  f = (n/2) log S, grad *= (n/2)/S.
 *   partials columns: [grad (ngrad of them), S, total ops, total slots]
 *   result layout   : [grad[0..P-1], f, total ops, total slots], written from first_col on
 */
extern "C" __global__ void __launch_bounds__(256) ad4cl_reduce_partials(
        const double* __restrict__ partials, int nblocks, int ncols, int ngrad, int first_col,
        int epilogue, double n_obs, double* __restrict__ result,
        const int* __restrict__ status = nullptr, double* __restrict__ status_out = nullptr) {
    __shared__ double scratch[2][8];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double v = 0.0, s = 0.0;
    for (int b = tid; b < nblocks; b += 256) {
        v += partials[(size_t) b * ncols + c];
        s += partials[(size_t) b * ncols + ngrad];
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
        if (epilogue == 1) {
            if (c < ngrad) tv *= (n_obs / 2.0) * (1.0 / ts);
            else if (c == ngrad) tv = (n_obs / 2.0) * log(ts);
        }
        result[first_col + c] = tv;
        /* the host path reads the sticky error bits with the result in one copy */
        if (c == 0 && status_out != nullptr) *status_out = (double) *status;
    }
}

/*
 * kAccGlobal: the gradient was accumulated by fp64 RED into global_grad[P]; copy it
 * into the result (scaled by the epilogue, whose S every block re-derives from the
 * partials in the same fixed order).
 */
extern "C" __global__ void __launch_bounds__(256) ad4cl_finish_global_grad(
        const double* __restrict__ global_grad, int nparams, const double* __restrict__ partials,
        int nblocks, int epilogue, double n_obs, double* __restrict__ result) {
    double scale = 1.0;
    if (epilogue == 1) {
        __shared__ double scratch[8];
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        double s = 0.0;
        for (int b = tid; b < nblocks; b += 256) s += partials[(size_t) b * 3];
        s = ad4cl::warp_sum(s);
        if (lane == 0) scratch[warp] = s;
        __syncthreads();
        double ts = 0.0;
        for (int w = 0; w < 8; w++) ts += scratch[w];
        scale = (n_obs / 2.0) * (1.0 / ts);
    }
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < nparams) result[i] = global_grad[i] * scale;
}
