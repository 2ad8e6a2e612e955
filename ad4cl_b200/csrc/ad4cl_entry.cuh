/*
 * ad4cl_entry.cuh -- the fused evaluation kernel around a user objective.
 *
 * Replaces, in one launch, the per-evaluation protocol of the reference's host
 * (test/admb/simple.cpp:283-384): NDRange record kernel -> read back the whole
 * stack -> host sum of out[] -> serial compute_gradient -> read g[param.id].
 *
 * A persistent grid (a multiple of the SM count) walks the work-groups of the
 * NDRange.  For every work-item a thread
 *   1. runs the user's kernel body (inlined; it records onto the thread tape),
 *   2. adds out.value to its running sum,
 *   3. sweeps its tape backwards seeded with d(out) = 1 (ad4cl_tape.cuh),
 *   4. rewinds the tape (the arena column is reused by its next work-item).
 * At the end the block reduces [d/d(theta_0..P-1), sum, #ops, #slots] with warp
 * shuffles + a fixed-order pass over the warps and writes one row of partials;
 * ad4cl_reduce_partials (second launch) sums the rows of all blocks in block
 * order, so the result is bit-reproducible for a given launch shape.
 *
 * The user's `out` argument is virtual: the body's `out[get_global_id(0)] = v`
 * lands in a register of the calling thread (no 16 B/work-item store to HBM as
 * in simple.cl:137).
 */
#pragma once

#include "ad.cuh"

namespace ad4cl {

struct LaunchInfo {
    long long n_items;      /* global work size (product over dimensions) */
    int g0;                 /* global size of dimension 0 (dimension 1 = n_items / g0) */
    int l0;                 /* local size of dimension 0 (local size 1 = blockDim.x / l0) */
    int nparams;            /* independents: ids 0..nparams-1 */
    int recording;          /* 1: objective + gradient, 0: objective only */
    int acc_mode;           /* AccMode */
    int window;             /* adjoint ring entries per thread (power of two) */
    int use_far;            /* far-adjoint plane present in the arena */
    int tape_in_smem;       /* 1: tape rows live in dynamic shared memory */
    int sweep_prefetch;     /* 1: the sweep prefetches two batches ahead into L2 (ad4cl_tape.cuh) */
    unsigned cap_slots;     /* per-thread tape capacity (operand slots) */
    char* arena;            /* HBM arena: one warp region per resident warp */
    unsigned long long warp_region;  /* bytes per warp region */
    double* partials;       /* [gridDim.x][ncols]; ncols = nparams + 3, or 3 with kAccGlobal: [grad..., sum, ops, slots] */
    double* global_grad;    /* kAccGlobal target, else null */
    int* status;            /* OR of Status bits */
    double* retain_meta;    /* non-null: RECORD ONLY.  One work-group per CTA (grid = number of groups), no
                               sweep, every thread keeps its tape in its own arena column and writes
                               {out.value, out.id, slots} here, 3 doubles per work-item (tape export path) */
};

/* Per-thread NDRange coordinates, published for get_global_id() and friends. */
struct ItemIds {
    int gid0, gid1;
};

#ifndef AD4CL_MAX_BLOCK
#define AD4CL_MAX_BLOCK 256
#endif

__shared__ ItemIds ad4cl_item_ids[AD4CL_MAX_BLOCK];
__shared__ int ad4cl_local_size0;

AD4CL_DEV int global_id(int d) {
    return d == 0 ? ad4cl_item_ids[threadIdx.x].gid0 : (d == 1 ? ad4cl_item_ids[threadIdx.x].gid1 : 0);
}
AD4CL_DEV int local_id(int d) {
    const int l0 = ad4cl_local_size0;
    return d == 0 ? (int) threadIdx.x % l0 : (d == 1 ? (int) threadIdx.x / l0 : 0);
}
AD4CL_DEV int local_size(int d) {
    const int l0 = ad4cl_local_size0;
    return d == 0 ? l0 : (d == 1 ? (int) blockDim.x / l0 : 1);
}
AD4CL_DEV int group_id(int d) {
    return global_id(d) / local_size(d);
}

extern __shared__ __align__(16) unsigned char ad4cl_dyn_smem[];

/*
 * Dynamic shared memory layout (doubles unless noted), B = blockDim.x, NW = B/32:
 *   ring      [B][window + 1]       thread-major, padded by one entry so that the lanes of a warp fall on
 *                                   distinct banks; re-used as the block reduction scratch red[NW][ncols]
 *                                   after the last work-item (the region is max((window+1)*B, NW*ncols) doubles)
 *   thread_acc[nparams][B]          kAccThread only
 *   warp_acc  [NW][nparams]         kAccWarp only
 *   tape rows [NW][smem_region]     tape_in_smem only (bytes)
 */
__host__ __device__ inline size_t entry_smem_bytes(int block, int nparams, int window, int acc_mode,
        bool tape_in_smem, unsigned cap_slots, bool far) {
    const int nw = block / 32;
    const size_t red = (size_t) nw * ((acc_mode == kAccGlobal ? 0 : nparams) + 3);
    size_t d = (size_t) (window + 1) * block;
    if (d < red) d = red;
    if (acc_mode == kAccThread) d += (size_t) nparams * block;
    if (acc_mode == kAccWarp) d += (size_t) nw * nparams;
    size_t bytes = d * sizeof (double);
    bytes = (bytes + 127) & ~(size_t) 127;
    if (tape_in_smem) bytes += (size_t) nw * warp_region_bytes(cap_slots, far);
    return bytes;
}

template <int MODE, typename Body>
__device__ __forceinline__ void run_objective(const LaunchInfo& L, Body&& body) {
    const int B = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = B >> 5;
    const int P = L.nparams;
    const int PR = MODE == kAccGlobal ? 0 : P;   /* gradient columns that go through the block reduction */
    const int ncols = PR + 3;

    double* smem = reinterpret_cast<double*> (ad4cl_dyn_smem);
    double* ring = smem;
    double* red = smem; /* the ring is dead (all zero) once the last work-item has been swept */
    size_t shared_region = (size_t) (L.window + 1) * B;
    if (shared_region < (size_t) nw * ncols) shared_region = (size_t) nw * ncols;
    double* thread_acc = ring + shared_region;
    double* warp_acc = thread_acc + (MODE == kAccThread ? (size_t) P * B : 0);
    size_t dbl_bytes = (size_t) ((warp_acc + (MODE == kAccWarp ? (size_t) nw * P : 0)) - smem) * sizeof (double);
    dbl_bytes = (dbl_bytes + 127) & ~(size_t) 127;

    for (int i = tid; i < (L.window + 1) * B; i += B) ring[i] = 0.0;
    if (MODE == kAccThread)
        for (int i = tid; i < P * B; i += B) thread_acc[i] = 0.0;
    if (MODE == kAccWarp)
        for (int i = tid; i < nw * P; i += B) warp_acc[i] = 0.0;
    if (tid == 0) ad4cl_local_size0 = L.l0;

    char* region;
    if (L.tape_in_smem) {
        region = reinterpret_cast<char*> (ad4cl_dyn_smem) + dbl_bytes + (size_t) warp * L.warp_region;
        if (L.use_far) { /* the far plane must start clear */
            double* fa = reinterpret_cast<double*> (region + (size_t) (L.cap_slots + kGuardRows) * kRowBytes);
            for (int i = lane; i < (int) (L.cap_slots + kGuardRows) * 32; i += 32) fa[i] = 0.0;
        }
    } else {
        region = L.arena + ((size_t) blockIdx.x * nw + warp) * L.warp_region;
    }
    __syncthreads();

    ThreadTape tape;
    tape_bind(tape, region, L.cap_slots, P, L.window, L.use_far != 0);

    SweepTarget T;
    T.ring = smem_u32(ring + (size_t) tid * (L.window + 1));
    T.stride = (unsigned) B * 8u;
    T.thread_acc = smem_u32(thread_acc + tid);
    T.warp_acc = smem_u32(warp_acc + (size_t) warp * P);
    T.global_acc = L.global_grad;

    double sum = 0.0;
    const long long ngroups = (L.n_items + B - 1) / B;
    const int groups0 = L.g0 / L.l0;           /* work-groups along dimension 0 */
    const int l1 = B / L.l0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        /* NDRange coordinates of this thread's work-item in work-group `grp` */
        const int grp0 = (int) (grp % groups0), grp1 = (int) (grp / groups0);
        const int gid0 = grp0 * L.l0 + tid % L.l0;
        const int gid1 = grp1 * l1 + tid / L.l0;
        ad4cl_item_ids[tid].gid0 = gid0;
        ad4cl_item_ids[tid].gid1 = gid1;
        const long long item = (long long) gid1 * L.g0 + gid0;

        struct ad_gradient_structure gs;
        gs.gradient_stack = nullptr;
        gs.current_ad_variable_id = P;
        gs.stack_current = 0;
        gs.recording = L.recording;
        gs.counter = 0;
        gs.index = nullptr;
        gs.tape = &tape;

        struct ad_variable outv = {0.0, kNoId};
        body(&gs, &outv - item);
        sum += outv.value;
        if (L.retain_meta != nullptr) { /* tape export: leave the rows in the arena for the host */
            double* m = L.retain_meta + 3 * item;
            m[0] = outv.value;
            m[1] = (double) outv.id;
            m[2] = (double) (tape.next_id - tape.first_temp);
            continue;
        }
        /* a thread whose tape overflowed skips the sweep (the host reports the error) */
        if (L.recording == 1 && __all_sync(0xffffffffu, tape.status == kOk)) tape_sweep<MODE>(tape, outv.id, T, L.sweep_prefetch != 0);
        tape_rewind(tape);
    }

    /* ---- block reduction: warp shuffles, then warps in fixed order ---- */
    __syncthreads();
    for (int p = 0; p < PR; p++) {
        double v;
        if (MODE == kAccThread) {
            v = warp_sum(thread_acc[(size_t) p * B + tid]);
        } else {
            v = warp_acc[(size_t) warp * P + p];
        }
        if (lane == 0) red[(size_t) warp * ncols + p] = v;
    }
    {
        const double s = warp_sum(sum);
        const double o = warp_sum((double) tape.total_ops);
        const double k = warp_sum((double) tape.total_slots);
        if (lane == 0) {
            red[(size_t) warp * ncols + PR] = s;
            red[(size_t) warp * ncols + PR + 1] = o;
            red[(size_t) warp * ncols + PR + 2] = k;
        }
    }
    const int st = __reduce_or_sync(0xffffffffu, tape.status);
    if (lane == 0 && st != 0) atomicOr(L.status, st);
    __syncthreads();
    for (int c = tid; c < ncols; c += B) {
        double v = 0.0;
        for (int wi = 0; wi < nw; wi++) v += red[(size_t) wi * ncols + c];
        L.partials[(size_t) blockIdx.x * ncols + c] = v;
    }
}

} // namespace ad4cl

/*
 * AD4CL_OBJECTIVE_ENTRY(NAME, (typed user parameters), (call arguments))
 * defines `NAME__entry`, the __global__ function the shim launches for the
 * user kernel NAME.  Inside the call-argument list, `gs`, `gradient_stack`
 * and `out` name the runtime-provided handles; everything else is forwarded
 * from the entry's own parameters.
 */
#define AD4CL_UNPAREN(...) __VA_ARGS__
#ifndef AD4CL_MIN_BLOCKS
#define AD4CL_MIN_BLOCKS 3 /* 3 x 256 threads per SM = 80 registers: measured faster than 4 (64 registers spill and rematerialise) */
#endif
#ifndef AD4CL_ENTRY_SUFFIX
#define AD4CL_ENTRY_SUFFIX /* the two ahead-of-time builds of one kernel need distinct symbols */
#endif
#define AD4CL_CAT4_(a, b, c, d) a##b##c##d
#define AD4CL_CAT4(a, b, c, d) AD4CL_CAT4_(a, b, c, d)
/* one __global__ function per accumulator mode, so a launch only brings the code of
   the sweep it uses into the instruction cache: NAME__entry<suffix>_m0 / _m1 / _m2 */
#define AD4CL_ENTRY_NAME(NAME, MODE) AD4CL_CAT4(NAME, __entry, AD4CL_ENTRY_SUFFIX, _m##MODE)
#define AD4CL_OBJECTIVE_ENTRY_MODE(NAME, MODE, PARAMS, ARGS)                               \
    extern "C" __global__ void __launch_bounds__(AD4CL_MAX_BLOCK, AD4CL_MIN_BLOCKS)        \
    AD4CL_ENTRY_NAME(NAME, MODE)(const ad4cl::LaunchInfo L, AD4CL_UNPAREN PARAMS) {        \
        ad4cl::run_objective<MODE>(L, [&](struct ad_gradient_structure* gs, struct ad_variable* out) { \
            struct ad_entry* gradient_stack = nullptr;                                     \
            (void) gradient_stack;                                                         \
            NAME(AD4CL_UNPAREN ARGS);                                                      \
        });                                                                                \
    }
#define AD4CL_OBJECTIVE_ENTRY(NAME, PARAMS, ARGS)                                         \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 0, PARAMS, ARGS)                                      \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 1, PARAMS, ARGS)                                      \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 2, PARAMS, ARGS)
