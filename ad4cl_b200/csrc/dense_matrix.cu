/*
 * dense_matrix.cu -- the dense-contraction path for config 2 (test/matrix).
 *
 * The objective of test/matrix is L = sum_{j,i} G[j,i] * C[j,i] with C = A*B
 * (matrix_mul.cpp:39-57, 253-261; G = 1 there).  Recording it operator by operator
 * (matrixmul.cl:5-28: 2n AD operators per element of C) costs 96 n^3 bytes of tape; its
 * derivative is two more matrix products,
 *        dL/dA = G * B^T          dL/dB = A^T * G,
 * 4 n^3 flop in total.  This file computes C, dL/dA and dL/dB with a hand-written fp64 GEMM
 * for sm_100a, in two variants selected at run time:
 *   variant 0  CUDA-core DFMA: 64x64 block tile, 64 threads, 8x8 register tile per thread
 *   variant 1  fp64 tensor-core DMMA (mma.sync.m8n8k4.f64): 64x64 block tile, 4 warps of 32x32
 * so that ncu / CUDA events decide which one the library uses (profiles/README.md).
 * Row-major n x n matrices, n a multiple of 64 (the reference uses 256; the config 1024).
 */
#include "../../include/ad4cl_cuda.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <map>
#include <mutex>

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

/* element (r, c) of op(X) where X is row-major n x n and op is identity or transpose */
template <bool T>
__device__ __forceinline__ double at(const double* __restrict__ X, int n, int r, int c) {
    return T ? X[(size_t) c * n + r] : X[(size_t) r * n + c];
}

constexpr int PAD = 4; /* row stride 68: the DMMA fragment reads (4 k-rows x 4..8 columns per half-warp) hit distinct banks */

/* Global -> registers for the next k-tile (issued before the arithmetic on the current tile, so the
   HBM/L2 latency overlaps it), then registers -> shared after the arithmetic.  Tiles are kept k-major
   in shared memory: As[k][m], Bs[k][n].  The contiguous direction of the global read follows the
   operand's storage, so the loads are coalesced for op = N and op = T alike. */
template <bool TA, bool TB, int NT, bool ONES_A = false, bool ONES_B = false>
struct TileStage {
    static constexpr int EA = BM * BK / NT, EB = BK * BN / NT;
    double a[EA], b[EB];

    /* ONES_A / ONES_B: that operand is the all-ones matrix (G = 1, the reference's objective sum C):
       nothing is read for it */
    __device__ __forceinline__ void fetch(const double* __restrict__ A, const double* __restrict__ B, int n,
            int m0, int n0, int k0, int tid) {
#pragma unroll
        for (int i = 0; i < EA; i++) {
            const int e = tid + i * NT;
            const int m = TA ? e % BM : e / BK, k = TA ? e / BM : e % BK;
            a[i] = ONES_A ? 1.0 : at<TA>(A, n, m0 + m, k0 + k);
        }
#pragma unroll
        for (int i = 0; i < EB; i++) {
            const int e = tid + i * NT;
            const int k = TB ? e % BK : e / BN, c = TB ? e / BK : e % BN;
            b[i] = ONES_B ? 1.0 : at<TB>(B, n, k0 + k, n0 + c);
        }
    }

    __device__ __forceinline__ void store(double (*As)[BM + PAD], double (*Bs)[BN + PAD], int tid) const {
#pragma unroll
        for (int i = 0; i < EA; i++) {
            const int e = tid + i * NT;
            const int m = TA ? e % BM : e / BK, k = TA ? e / BM : e % BK;
            As[k][m] = a[i];
        }
#pragma unroll
        for (int i = 0; i < EB; i++) {
            const int e = tid + i * NT;
            const int k = TB ? e % BK : e / BN, c = TB ? e / BK : e % BN;
            Bs[k][c] = b[i];
        }
    }
};

/* ---------------------------------------------------------------- variant 0: DFMA */
template <bool TA, bool TB>
__global__ void __launch_bounds__(64) dgemm_fma(const double* __restrict__ A, const double* __restrict__ B,
        double* __restrict__ C, int n, double* __restrict__ block_sums) {
    __shared__ double As[2][BK][BM + PAD];
    __shared__ double Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0;

    TileStage<TA, TB, 64> stage;
    stage.fetch(A, B, n, m0, n0, 0, tid);
    stage.store(As[0], Bs[0], tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < n; k0 += BK) {
        const bool more = k0 + BK < n;
        if (more) stage.fetch(A, B, n, m0, n0, k0 + BK, tid);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = As[buf][kk][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 8; j++) b[j] = Bs[buf][kk][tx * 8 + j];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (more) stage.store(As[buf ^ 1], Bs[buf ^ 1], tid);
        __syncthreads();
        buf ^= 1;
    }
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double* row = C + (size_t) (m0 + ty * 8 + i) * n + n0 + tx * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            row[j] = acc[i][j];
            local += acc[i][j];
        }
    }
    if (block_sums) { /* fixed-order block sum of this tile of C (the objective is sum C) */
        __shared__ double red[64];
        red[tid] = local;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int t = 0; t < 64; t++) s += red[t];
            block_sums[blockIdx.y * gridDim.x + blockIdx.x] = s;
        }
    }
}

/* ---------------------------------------------------------------- variant 1: DMMA */
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(128) dgemm_mma(const double* __restrict__ A, const double* __restrict__ B,
        double* __restrict__ C, int n, double* __restrict__ block_sums) {
    __shared__ double As[2][BK][BM + PAD];
    __shared__ double Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     /* this warp's 32 x 32 sub-tile */
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int g = lane >> 2, q = lane & 3;                      /* fragment coordinates (PTX m8n8k4 layout) */
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    TileStage<TA, TB, 128> stage;
    stage.fetch(A, B, n, m0, n0, 0, tid);
    stage.store(As[0], Bs[0], tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < n; k0 += BK) {
        const bool more = k0 + BK < n;
        if (more) stage.fetch(A, B, n, m0, n0, k0 + BK, tid);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[buf][k4 + q][wm + i * 8 + g];   /* A[row g][col q] */
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[buf][k4 + q][wn + j * 8 + g];   /* B[row q][col g] */
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (more) stage.store(As[buf ^ 1], Bs[buf ^ 1], tid);
        __syncthreads();
        buf ^= 1;
    }
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {       /* C fragment: row g, columns 2q and 2q+1 */
            double* p = C + (size_t) (m0 + wm + i * 8 + g) * n + n0 + wn + j * 8 + 2 * q;
            p[0] = acc[i][j][0];
            p[1] = acc[i][j][1];
            local += acc[i][j][0] + acc[i][j][1];
        }
    if (block_sums) {
        __shared__ double red[128];
        red[tid] = local;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int t = 0; t < 128; t++) s += red[t];
            block_sums[blockIdx.y * gridDim.x + blockIdx.x] = s;
        }
    }
}

/* ---------------------------------------------------------------- variant 2: all three products, ONE launch
 * C = A B, dA = G B^T and dB = A^T G are 3 (n/64)^2 independent 64x64 tiles.  A persistent grid (as many CTAs
 * as are resident) draws tiles from one atomic counter, longest-first order being irrelevant (all tiles cost the
 * same): the last wave is at most one tile per CTA instead of the 1.7-wave grid of one launch per product, and
 * two launches, the all-ones fill of G and the separate sum kernel disappear -- the CTA that finishes the last
 * tile adds the per-tile sums of C in tile order.  The tile arithmetic is dgemm_mma's. */
struct Gemm3 {
    const double* A; const double* B; const double* G;   /* G == nullptr: all ones */
    double* C; double* dA; double* dB;                    /* each may be nullptr: product skipped */
    double* tile_sums;                                    /* (n/64)^2 doubles */
    double* sum_out;                                      /* may be nullptr */
    unsigned* counters;                                   /* [0] next tile, [1] tiles finished; both zero at launch, reset by the last CTA */
    int n;
};

template <bool TA, bool TB, bool ONES_A, bool ONES_B>
__device__ __forceinline__ double gemm_tile_mma(const double* __restrict__ A, const double* __restrict__ B,
        double* __restrict__ C, int n, int m0, int n0, double (*As)[BK][BM + PAD], double (*Bs)[BK][BN + PAD]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    TileStage<TA, TB, 128, ONES_A, ONES_B> stage;
    stage.fetch(A, B, n, m0, n0, 0, tid);
    __syncthreads();                       /* the previous tile's readers are done with the buffers */
    stage.store(As[0], Bs[0], tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < n; k0 += BK) {
        const bool more = k0 + BK < n;
        if (more) stage.fetch(A, B, n, m0, n0, k0 + BK, tid);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[buf][k4 + q][wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[buf][k4 + q][wn + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (more) stage.store(As[buf ^ 1], Bs[buf ^ 1], tid);
        __syncthreads();
        buf ^= 1;
    }
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double* p = C + (size_t) (m0 + wm + i * 8 + g) * n + n0 + wn + j * 8 + 2 * q;
            *reinterpret_cast<double2*> (p) = make_double2(acc[i][j][0], acc[i][j][1]);
            local += acc[i][j][0] + acc[i][j][1];
        }
    return local;
}

__global__ void __launch_bounds__(128, 4) dgemm3_persistent(const Gemm3 P) {
    __shared__ double As[2][BK][BM + PAD];
    __shared__ double Bs[2][BK][BN + PAD];
    __shared__ double red[128];
    __shared__ unsigned next_tile, finished;
    const int tid = threadIdx.x;
    const int tn = P.n / BN, per = tn * tn;
    /* products in a fixed order; a skipped product contributes no tiles */
    const int base1 = P.C ? per : 0, base2 = base1 + (P.dA ? per : 0), total = base2 + (P.dB ? per : 0);
    for (;;) {
        if (tid == 0) next_tile = atomicAdd(&P.counters[0], 1u);
        __syncthreads();
        const unsigned t = next_tile;
        if (t >= (unsigned) total) break;
        const int which = (int) t < base1 ? 0 : ((int) t < base2 ? 1 : 2);
        const int local_t = (int) t - (which == 0 ? 0 : (which == 1 ? base1 : base2));
        const int m0 = (local_t / tn) * BM, n0 = (local_t % tn) * BN;
        double s;
        if (which == 0) s = gemm_tile_mma<false, false, false, false>(P.A, P.B, P.C, P.n, m0, n0, As, Bs);
        else if (which == 1) s = P.G ? gemm_tile_mma<false, true, false, false>(P.G, P.B, P.dA, P.n, m0, n0, As, Bs)
                                     : gemm_tile_mma<false, true, true, false>(P.G, P.B, P.dA, P.n, m0, n0, As, Bs);
        else s = P.G ? gemm_tile_mma<true, false, false, false>(P.A, P.G, P.dB, P.n, m0, n0, As, Bs)
                     : gemm_tile_mma<true, false, false, true>(P.A, P.G, P.dB, P.n, m0, n0, As, Bs);
        if (which == 0 && P.sum_out) { /* fixed-order sum of this tile of C */
            red[tid] = s;
            __syncthreads();
            if (tid == 0) {
                double v = 0.0;
                for (int i = 0; i < 128; i++) v += red[i];
                P.tile_sums[local_t] = v;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) finished = atomicAdd(&P.counters[1], 1u);
        __syncthreads();
        if (finished == (unsigned) total - 1) { /* the last tile of the launch: the objective, and leave the counters ready */
            __threadfence();
            if (P.sum_out && P.C && tid == 0) {
                double v = 0.0;
                for (int i = 0; i < per; i++) v += reinterpret_cast<volatile double*> (P.tile_sums)[i];
                *P.sum_out = v;
            }
            if (tid == 0) P.counters[1] = 0u;
        }
    }
    /* every CTA leaves through one more draw (>= total); the one that sees the last of them resets the draw counter */
    if (tid == 0) {
        const unsigned draws = total + gridDim.x;
        if (next_tile == draws - 1) P.counters[0] = 0u;
    }
}

__global__ void fill_ones(double* p, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.0;
}

__global__ void sum_fixed_order(const double* __restrict__ v, int n, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += v[i];
        *out = s;
    }
}

/* fp64 FMA microbenchmark: 8 independent chains per thread, the denominator for the dense path */
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <bool TA, bool TB>
cudaError_t gemm(int variant, const double* A, const double* B, double* C, int n, double* block_sums, cudaStream_t st) {
    dim3 grid(n / BN, n / BM);
    if (variant == 1) dgemm_mma<TA, TB><<<grid, 128, 0, st>>>(A, B, C, n, block_sums);
    else dgemm_fma<TA, TB><<<grid, 64, 0, st>>>(A, B, C, n, block_sums);
    return cudaGetLastError();
}

} // namespace

extern "C" {

void* ad4cl_cuda_stream(ad4cl_context* ctx);
const char* ad4cl_cuda_last_error(void);
int ad4cl_cuda_set_error_(int code, const char* msg); /* shim.cu */

#pragma GCC visibility push(default)
/*
 * C = A*B, sum C, dA = G*B^T, dB = A^T*G on device memory (row-major n x n doubles).
 * G_dev may be NULL (G = 1: the reference's objective sum C); C_dev, dA_dev, dB_dev and sum_dev
 * may each be NULL to skip that output; workspace_dev: n*n + (n/64)^2 doubles of scratch.
 * variant 0: DFMA, 1: DMMA.  Asynchronous on the context's stream.
 */
int ad4cl_cuda_matrix_product_dense(ad4cl_context* ctx, int variant, int n, const double* A_dev, const double* B_dev,
        const double* G_dev, double* C_dev, double* dA_dev, double* dB_dev, double* sum_dev, double* workspace_dev) {
    if (!ctx || !A_dev || !B_dev || !workspace_dev) return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "NULL argument");
    if (n <= 0 || n % 64 != 0) return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "dense path needs n to be a multiple of 64");
    if (variant < 0 || variant > 2)
        return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "variant must be 0 (DFMA), 1 (DMMA, one launch per product) or 2 (DMMA, all products in one persistent launch)");
    cudaStream_t st = (cudaStream_t) ad4cl_cuda_stream(ctx);
    const size_t nn = (size_t) n * n;
    double* scratch = workspace_dev;          /* C when the caller does not want it, or G = 1 */
    double* sums = workspace_dev + nn;
    const int nblocks = (n / 64) * (n / 64);
    cudaError_t e = cudaSuccess;
    if (variant == 2) {
        /* the two tile counters live with the CONTEXT (its launches are ordered by its stream; two contexts on
           one device must not share them); zero between launches -- the kernel resets them */
        static std::mutex mu;
        static std::map<ad4cl_context*, unsigned*> counters_of;
        static std::map<int, int> resident_of;
        int dev = 0;
        cudaGetDevice(&dev);
        unsigned* counters = nullptr;
        int resident = 0;
        {
            std::lock_guard<std::mutex> lock(mu);
            unsigned*& slot = counters_of[ctx];
            if (!slot) {
                if (cudaMalloc(&slot, 2 * sizeof (unsigned)) != cudaSuccess || cudaMemset(slot, 0, 2 * sizeof (unsigned)) != cudaSuccess) {
                    slot = nullptr;
                    return ad4cl_cuda_set_error_(AD4CL_ERR_NOMEM, "dense path: counter allocation failed");
                }
            }
            counters = slot;
            int& res = resident_of[dev];
            if (!res) {
                int per_sm = 0, sms = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dgemm3_persistent, 128, 0);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                res = per_sm * sms;
            }
            resident = res;
        }
        if (resident < 1) return ad4cl_cuda_set_error_(AD4CL_ERR_CUDA, "dense path: the persistent kernel does not fit on this device");
        Gemm3 P;
        P.A = A_dev; P.B = B_dev; P.G = G_dev;
        P.C = (C_dev || sum_dev) ? (C_dev ? C_dev : scratch) : nullptr;
        P.dA = dA_dev; P.dB = dB_dev;
        P.tile_sums = sums; P.sum_out = sum_dev; P.counters = counters; P.n = n;
        const int total = ((P.C ? 1 : 0) + (P.dA ? 1 : 0) + (P.dB ? 1 : 0)) * nblocks;
        if (total > 0) dgemm3_persistent<<<resident < total ? resident : total, 128, 0, st>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            char msg[256];
            snprintf(msg, sizeof msg, "dense matrix product failed: %s", cudaGetErrorString(e));
            return ad4cl_cuda_set_error_(AD4CL_ERR_CUDA, msg);
        }
        return AD4CL_OK;
    }
    if (C_dev || sum_dev) {
        double* Cout = C_dev ? C_dev : scratch;
        e = gemm<false, false>(variant, A_dev, B_dev, Cout, n, sum_dev ? sums : nullptr, st);
        if (e == cudaSuccess && sum_dev) sum_fixed_order<<<1, 32, 0, st>>>(sums, nblocks, sum_dev);
    }
    const double* G = G_dev;
    if (e == cudaSuccess && !G && (dA_dev || dB_dev)) {
        fill_ones<<<(unsigned) ((nn + 255) / 256), 256, 0, st>>>(scratch, nn);
        G = scratch;
    }
    if (e == cudaSuccess && dA_dev) e = gemm<false, true>(variant, G, B_dev, dA_dev, n, nullptr, st);
    if (e == cudaSuccess && dB_dev) e = gemm<true, false>(variant, A_dev, G, dB_dev, n, nullptr, st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof msg, "dense matrix product failed: %s", cudaGetErrorString(e));
        return ad4cl_cuda_set_error_(AD4CL_ERR_CUDA, msg);
    }
    return AD4CL_OK;
}
/* Measured fp64 FMA rate of the device in TFLOP/s (2 flop per DFMA), the roofline denominator
   SURVEY.md 8(d) asks to measure rather than quote.  Synchronous; ~20 ms. */
int ad4cl_cuda_measure_fp64_peak(ad4cl_context* ctx, double* tflops) {
    if (!ctx || !tflops) return ad4cl_cuda_set_error_(AD4CL_ERR_INVALID, "NULL argument");
    cudaStream_t st = (cudaStream_t) ad4cl_cuda_stream(ctx);
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaMalloc(&out, sizeof (double) * blocks * threads);
    if (e != cudaSuccess) return ad4cl_cuda_set_error_(AD4CL_ERR_NOMEM, cudaGetErrorString(e));
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_peak<<<blocks, threads, 0, st>>>(out, 1000);
    cudaEventRecord(e0, st);
    dfma_peak<<<blocks, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return ad4cl_cuda_set_error_(AD4CL_ERR_CUDA, cudaGetErrorString(e));
    *tflops = 2.0 * 8.0 * iters * (double) blocks * threads / (ms * 1e-3) / 1e12;
    return AD4CL_OK;
}
#pragma GCC visibility pop

} // extern "C"
