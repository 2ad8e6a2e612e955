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

constexpr int kMaxWorld = 16;          /* ranks of a one-shot all-reduce group (one box) */
constexpr int kInKernelPackMax = 4096; /* independents a CTA converts to ad_variables itself */

/* One-shot all-reduce group as seen by this rank (filled from an ad4cl_comm; world <= 1: none). */
struct CommInfo {
    double* inbox[kMaxWorld];       /* rank r's inbox [2 parities][world][stride] as mapped into this process */
    unsigned long long* epoch;      /* this rank's evaluation counter (device memory: graph replays advance it) */
    int rank, world, stride;        /* stride = doubles per slot, the last one is the arrival flag */
    unsigned long long timeout_ns;  /* a peer that has not arrived after this long fails the evaluation */
};

struct LaunchInfo {
    long long n_items;      /* global work size (product over dimensions, padded to whole work-groups) */
    int g0;                 /* global size of dimension 0, padded (dimension 1 = n_items / g0) */
    int g0_user;            /* global size of dimension 0 as configured: out[] is indexed with it */
    int l0;                 /* local size of dimension 0 (local size 1 = blockDim.x / l0) */
    int nparams;            /* independents: ids 0..nparams-1 */
    int recording;          /* 1: objective + gradient, 0: objective only */
    int acc_mode;           /* AccMode */
    int window;             /* adjoint ring entries per thread (power of two >= 4) */
    int use_far;            /* far-adjoint plane present in the arena */
    int tape_in_smem;       /* 1: tape planes live in dynamic shared memory */
    unsigned cap_slots;     /* per-thread tape capacity (operand slots) */
    char* arena;            /* HBM arena: one warp region per resident warp */
    unsigned long long warp_region;  /* bytes per warp region */
    double* partials;       /* [gridDim.x][ncols]; ncols = nparams + 3, or 3 with kAccGlobal: [grad..., sum, ops, slots] */
    double* global_grad;    /* kAccGlobal target, else null */
    int* status;            /* OR of Status bits */
    double* retain_meta;    /* non-null: RECORD ONLY.  One work-group per CTA (grid = number of groups), no
                               sweep, every thread keeps its tape in its own arena column and writes
                               {out.value, out.id, slots} here, 3 doubles per work-item (tape export path) */
    /* ---- single-launch evaluation (round 2): parameter packing and the second reduction stage in this kernel */
    const double* theta;    /* non-null: nparams plain doubles; every CTA converts them into `params` first */
    struct ad_variable* params; /* the {value, id = i} array the user kernel's V arguments point into */
    unsigned* ticket;       /* non-null: the CTA that finishes last reduces `partials` in fixed order into `result` */
    double* result;         /* [grad (nparams), f or S, ops, slots] */
    double* status_out;     /* non-null: receives (double) *status next to the result (one D2H copy for both) */
    int epilogue;           /* 1: f = (n/2) log S, grad *= (n/2) / S  (simple.cpp:365) */
    double n_obs;
    CommInfo comm;          /* world > 1: the last CTA also all-reduces the raw sums with the peers */
    /* ---- general epilogues: f = h(out_1 .. out_n) in two passes, nothing retained.
       Pass 1 (out_values set, objective only): every work-item stores out_i.value.  The caller turns them
       into seeds s_i = dh/d out_i.  Pass 2 (seeds set): the sweep of work-item i starts from s_i instead of 1,
       so the reduced gradient is sum_i s_i * d out_i / d theta.  Both indexed like the kernel's own out[]. */
    int sweep_prefetch;     /* general sweep: L2 prefetch two batches ahead */
    int lockstep;           /* the warps of a work-group wait for each other before every work-item */
    const double* seeds;
    double* out_values;
    long long n_user;       /* work-items the caller configured (the NDRange is padded to whole work-groups) */
    const int* skip;        /* non-null: the evaluation is skipped when *skip != 0 at execution time (a device-side
                               predicate: evaluations captured in a CUDA graph after an optimiser has converged) */
};

/* Per-thread NDRange coordinates, published for get_global_id() and friends. */
struct ItemIds {
    int gid0, gid1;
};

#ifndef AD4CL_STATIC_MAX_BLOCK
#define AD4CL_STATIC_MAX_BLOCK 256   /* register tier: the scalarised tape wants the registers of a 256-thread block */
#endif
#ifndef AD4CL_MAX_BLOCK
#define AD4CL_MAX_BLOCK 768
#endif
/* 1: the user kernel synchronises its work-group (barrier()): it cannot be re-run by a single warp, so
   the entry is built with the general recorder only (the shim defines this for JIT kernels whose text
   contains a barrier; none of the built-in kernels has one) */
#ifndef AD4CL_BODY_HAS_BARRIER
#define AD4CL_BODY_HAS_BARRIER 0
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
 *   tape      [NW][warp region]     tape_in_smem only (bytes)
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

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/*
 * One-shot all-reduce of the short raw-sum vector over NVLink peer memory, by ONE CTA (the last one
 * of the evaluation kernel, or the stand-alone kernel of ad4cl_cuda_comm_allreduce).  Every rank
 * stores its vector into slot [parity][rank] of EVERY rank's inbox (peer stores), publishes a flag
 * carrying the epoch, waits until all `world` flags of its own inbox show the epoch and sums the
 * slots in rank order -- all ranks obtain bit-identical results, the order does not depend on who
 * arrives first, and the latency is one NVLink store + one flag flight.  Two parities: a rank can
 * run at most one epoch ahead of the slowest.  `v` (shared or global memory, n doubles) holds this
 * rank's vector on entry and the sum over ranks on return; returns false after a timeout, with
 * NaN in v (a communicator that timed out must be rebuilt: its epochs are no longer aligned).
 */
__device__ __forceinline__ bool oneshot_allreduce_cta(const CommInfo& C, double* v, int n, unsigned long long epoch) {
    const int tid = threadIdx.x, B = blockDim.x;
    const size_t slot_of_me = ((size_t) (epoch & 1) * C.world + C.rank) * C.stride;
    for (int idx = tid; idx < C.world * n; idx += B) {
        const int r = idx / n, i = idx - r * n;
        C.inbox[r][slot_of_me + i] = v[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < C.world) {
        volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*> (C.inbox[tid] + slot_of_me + C.stride - 1);
        *flag = epoch;
    }
    double* mine = C.inbox[C.rank] + (size_t) (epoch & 1) * C.world * C.stride;
    __shared__ int failed;
    if (tid == 0) failed = 0;
    __syncthreads();
    if (tid < C.world) {
        const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*> (mine + (size_t) tid * C.stride + C.stride - 1);
        const unsigned long long t0 = global_timer_ns();
        while (*flag != epoch) {
            if (global_timer_ns() - t0 > C.timeout_ns) { /* a peer never arrived: report instead of hanging */
                failed = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    const bool ok = failed == 0;
    for (int i = tid; i < n; i += B) {
        double s = 0.0;
        for (int r = 0; r < C.world; r++) s += *reinterpret_cast<const volatile double*> (mine + (size_t) r * C.stride + i);
        v[i] = ok ? s : __longlong_as_double(0x7ff8000000000000LL);
    }
    __syncthreads();
    return ok;
}

/* the reference's scalar epilogue on the reduced vector [grad (P), S, ...]  (simple.cpp:365, main.cpp:417) */
__device__ __forceinline__ double apply_epilogue(double v, int c, int nparams, double S, double n_obs) {
    if (c < nparams) return v * ((n_obs / 2.0) * (1.0 / S));
    if (c == nparams) return (n_obs / 2.0) * log(S);
    return v;
}

/*
 * Second reduction stage, run by the CTA that finishes LAST (ticket): sums the per-block partial
 * rows in a fixed order (column c by warp c mod NW; lane l adds blocks l, l+32, ... ascending, then
 * the xor butterfly), so the result is bit-reproducible for a given launch shape regardless of
 * which CTA happens to be last.  Then the all-reduce over ranks (if any), the epilogue, the result.
 * `scratch` holds ncols doubles of shared memory.
 */
__device__ __forceinline__ void finish_evaluation(const LaunchInfo& L, int ncols, int ngrad, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int nblocks = gridDim.x;
    /* one CTA reads grid x ncols doubles: what it costs is load LATENCY, so every lane keeps 4 columns x 8 rows
       of independent loads in flight (the additions stay in ascending block order per column) */
    for (int c0 = warp; c0 < ncols; c0 += 4 * nw) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int b = lane; b < nblocks; b += 8 * 32) {
            double t[4][8];
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int bb = b + 32 * j, c = c0 + q * nw;
                    t[q][j] = (bb < nblocks && c < ncols) ? __ldcg(L.partials + (size_t) bb * ncols + c) : 0.0;
                }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int j = 0; j < 8; j++) v[q] += t[q][j];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double s = warp_sum(v[q]);
            if (lane == 0 && c0 + q * nw < ncols) scratch[c0 + q * nw] = s;
        }
    }
    __syncthreads();
    bool ok = true;
    if (L.comm.world > 1) {
        __shared__ unsigned long long epoch_s;
        if (tid == 0) epoch_s = ++(*L.comm.epoch);
        __syncthreads();
        ok = oneshot_allreduce_cta(L.comm, scratch, ncols, epoch_s);
        if (!ok && tid == 0) atomicOr(L.status, 4);
    }
    const double S = scratch[ngrad];
    for (int c = tid; c < ncols; c += blockDim.x) {
        double v = scratch[c];
        if (L.epilogue == 1 && ngrad == L.nparams) v = apply_epilogue(v, c, ngrad, S, L.n_obs);
        L.result[(ngrad == L.nparams ? 0 : L.nparams) + c] = v;
    }
    if (tid == 0) {
        if (L.status_out != nullptr) *L.status_out = (double) *reinterpret_cast<volatile int*> (L.status);
        *L.ticket = 0u; /* ready for the next evaluation */
    }
}

#if AD4CL_TAPE_STATIC
/*
 * The evaluation kernel of the REGISTER tier (ad4cl_tape.cuh): same persistent walk over the
 * work-groups, but the tape, the adjoints and the gradient accumulators of a thread are registers.
 * The user kernel receives LOCAL copies of the independents, th[p] = {theta_p, id = p} with literal
 * ids, so that every index of the unrolled forward and reverse code is a compile-time constant.
 * No warp-level operation inside the work-item loop: divergent kernels need no special care here.
 * Dynamic shared memory: NW * (AD4CL_P + 3) doubles (block reduction scratch).
 */
template <int RECORDING, typename Body>
__device__ __forceinline__ void run_objective_static(const LaunchInfo& L, Body&& body) {
    if (L.skip != nullptr && *reinterpret_cast<const volatile int*> (L.skip) != 0) return;
    constexpr int P = AD4CL_P;
    constexpr int ncols = P + 3;
    const int B = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = B >> 5;
    double* red = reinterpret_cast<double*> (ad4cl_dyn_smem);
    if (tid == 0) ad4cl_local_size0 = L.l0;
    __syncthreads();

    struct ad_variable th[P > 0 ? P : 1];
#pragma unroll
    for (int p = 0; p < P; p++) {
        th[p].value = L.theta[p];
        th[p].id = p;
    }
    double g[P > 0 ? P : 1];
#pragma unroll
    for (int p = 0; p < P; p++) g[p] = 0.0;

    double sum = 0.0;
    long long total_ops = 0, total_slots = 0;
    int status = kOk;
    const long long ngroups = (L.n_items + B - 1) / B;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        /* a FRESH tape per work-item: nothing loop-carried stands between the stores of the forward
           pass and the loads of the sweep, so the compiler sees the id words as the constants they are */
        ThreadTape tape;
        tape_bind_static(tape);
        /* the register-tier instances are 1-D kernels (the shim checks): no 64-bit division per work-item,
           out[get_global_id(0)] */
        const int gid0 = (int) grp * B + tid;
        ad4cl_item_ids[tid].gid0 = gid0;
        ad4cl_item_ids[tid].gid1 = 0;
        const long long out_index = gid0;

        struct ad_gradient_structure gs;
        gs.gradient_stack = nullptr;
        gs.current_ad_variable_id = P;
        gs.stack_current = 0;
        gs.recording = RECORDING; /* a literal: every operator's `if (gs->recording == 1)` folds */
        gs.counter = 0;
        gs.index = nullptr;
        gs.tape = &tape;

        struct ad_variable outv = {0.0, kNoId};
        body(&gs, &outv - out_index, th);
        sum += outv.value;
        const bool mine = gid0 < L.g0_user && out_index < L.n_user;
        if (L.out_values != nullptr && mine) L.out_values[out_index] = outv.value;
        const double seed = L.seeds == nullptr ? 1.0 : (mine ? L.seeds[out_index] : 0.0);
        if (RECORDING == 1 && tape.status == kOk) tape_sweep_static(tape, outv.id, g, seed);
        total_ops += tape.nops;
        total_slots += tape.next_id - P;
        status |= tape.status;
    }

    /* ---- block reduction: warp shuffles, then warps in fixed order ---- */
#pragma unroll
    for (int p = 0; p < P; p++) {
        const double v = warp_sum(g[p]);
        if (lane == 0) red[(size_t) warp * ncols + p] = v;
    }
    {
        const double s = warp_sum(sum);
        const double o = warp_sum((double) total_ops);
        const double k = warp_sum((double) total_slots);
        if (lane == 0) {
            red[(size_t) warp * ncols + P] = s;
            red[(size_t) warp * ncols + P + 1] = o;
            red[(size_t) warp * ncols + P + 2] = k;
        }
    }
    const int st = __reduce_or_sync(0xffffffffu, status);
    if (lane == 0 && st != 0) atomicOr(L.status, st);
    __syncthreads();
    double mine = 0.0;
    if (tid < ncols) {
        for (int wi = 0; wi < nw; wi++) mine += red[(size_t) wi * ncols + tid];
        L.partials[(size_t) blockIdx.x * ncols + tid] = mine;
    }
    if (L.ticket != nullptr) {
        __shared__ unsigned last;
        __threadfence();
        __syncthreads();
        if (tid == 0) last = atomicAdd(L.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (last) {
            __threadfence();
            finish_evaluation(L, ncols, P, red);
        }
    }
}
#else
/* POLICY = kRecordGeneral: the LANE kernel -- one copy of the user kernel, the general recorder only (the
   round-1 code path, with its instruction-cache footprint).  POLICY = kRecordUniform: the UNIFORM kernel --
   the optimistic uniform recorder first, a second copy of the user kernel with the general recorder for the
   work-items (and, after three failures in a row, the warps) it cannot take.  Which kernel a launch uses is
   the shim's choice (AD4CL_RECORDER_*). */
template <int MODE, int POLICY, typename Body>
__device__ __forceinline__ void run_objective(const LaunchInfo& L, Body&& body) {
    if (L.skip != nullptr && *reinterpret_cast<const volatile int*> (L.skip) != 0) return;
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
    /* ad_init_var (ad4cl.h:90-93) for the independents: {theta_i, id = i}.  Every CTA writes the same
       values, and reads them only after its own barrier below. */
    if (L.theta != nullptr) {
        for (int i = tid; i < P; i += B) {
            L.params[i].value = L.theta[i];
            L.params[i].id = i;
        }
    }

    char* region;
    if (L.tape_in_smem) {
        region = reinterpret_cast<char*> (ad4cl_dyn_smem) + dbl_bytes + (size_t) warp * L.warp_region;
        if (L.use_far) { /* the far plane must start clear */
            double* fa = reinterpret_cast<double*> (region + region_layout(L.cap_slots, true).off_far);
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
    PendingParam pend = {0.0, -1};
    /* Which recorder a work-item starts with is decided per warp: the optimistic uniform recorder
       until it failed for 3 work-items in a row (a kernel with lane-dependent parameter ids or divergent
       control flow), then the general one for good.  Warp-uniform by construction (set from votes). */
    int opt_fails = 0;
    const long long ngroups = (L.n_items + B - 1) / B;
    const int groups0 = L.g0 / L.l0;           /* work-groups along dimension 0 */
    const int l1 = B / L.l0;
    const bool one_d = L.l0 == B && (long long) groups0 == ngroups;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        /* lock step: keep the warps of a CTA inside the same work-item's code -- 3 (or 1) fetch streams
           per SM instead of 24; the hot code of a long objective is larger than the 32 KB L1.5 instruction
           cache and the warps otherwise starve on instruction fetch (profiles/README.md) */
        if (L.lockstep) __syncthreads();
        /* NDRange coordinates of this thread's work-item in work-group `grp` */
        int gid0, gid1;
        if (one_d) { /* no 64-bit division per work-item */
            gid0 = (int) grp * B + tid;
            gid1 = 0;
        } else {
            const int grp0 = (int) (grp % groups0), grp1 = (int) (grp / groups0);
            gid0 = grp0 * L.l0 + tid % L.l0;
            gid1 = grp1 * l1 + tid / L.l0;
        }
        ad4cl_item_ids[tid].gid0 = gid0;
        ad4cl_item_ids[tid].gid1 = gid1;
        const long long item = (long long) gid1 * L.g0 + gid0;
        /* The user's `out` is virtual: out[gid0 + g0_user * gid1] is a register of this thread.  The
           index convention is the reference's own (simple.cl:137 out[id], matrixmul.cl:27
           C[i + widthA * j] with widthA = global size 0); a kernel must write its own element only. */
        const long long out_index = (long long) gid1 * L.g0_user + gid0;

        struct ad_gradient_structure gs;
        gs.gradient_stack = nullptr;
        gs.current_ad_variable_id = P;
        gs.stack_current = 0;
        gs.recording = L.recording;
        gs.counter = 0;
        gs.index = nullptr;
        gs.tape = &tape;

        struct ad_variable outv = {0.0, kNoId};
        const bool mine = gid0 < L.g0_user && out_index < L.n_user;
        const double seed = L.seeds == nullptr ? 1.0 : (mine ? L.seeds[out_index] : 0.0);
        bool general = true;
#if !AD4CL_BODY_HAS_BARRIER
        if (POLICY == kRecordUniform) general = opt_fails >= 3 || L.recording != 1 || L.retain_meta != nullptr;
        if (POLICY == kRecordUniform && !general) {
            /* ---- first copy of the user kernel: the uniform recorder ---- */
            tape_begin(tape, kRecordUniform);
            body(&gs, &outv - out_index);
            /* ONE vote: no lane saw anything the uniform recorder cannot take, and all lanes seed the same id */
            int same = 0;
            __match_all_sync(0xffffffffu, (tape.bad || tape_overflowed(tape)) ? 0xffffffff00000000ull + (unsigned) lane
                                                                              : tape_uniform_key(tape, outv.id), &same);
            /* classify the rows (once per row for the whole warp); a far operand without a far plane is the
               one thing it can still refuse */
            if (same && tape_classify_uniform(tape_hdr(tape), tape.far_adj, tape.next_id - tape.first_temp, tape.first_temp, tape.wm1)) same = 0;
            if (same) {
                opt_fails = 0;
                if (L.sweep_prefetch) tape_sweep_uniform<MODE, true>(tape, outv.id, seed, T, pend);
                else tape_sweep_uniform<MODE, false>(tape, outv.id, seed, T, pend);
            } else {
                opt_fails++;
                general = true;
                outv.value = 0.0;
                outv.id = kNoId;
                gs.counter = 0;
            }
        }
#endif
        if (general) {
            /* ---- second copy: the general recorder (also: objective only, tape export) ---- */
            tape_begin(tape, kRecordGeneral);
            body(&gs, &outv - out_index);
            if (tape_overflowed(tape)) tape.status |= kErrTapeOverflow;
            if (L.retain_meta != nullptr) { /* tape export: leave the rows in the arena for the host */
                double* m = L.retain_meta + 3 * item;
                m[0] = outv.value;
                m[1] = (double) outv.id;
                m[2] = (double) (tape.next_id - tape.first_temp);
            } else if (L.recording == 1 && __all_sync(0xffffffffu, tape.status == kOk)) {
                /* a thread whose tape overflowed skips the sweep (the host reports the error) */
                if (MODE == kAccWarp) param_flush(T.warp_acc, pend);   /* keep the order of the additions fixed */
                tape.status |= tape_sweep_general<MODE>(tape_rows(tape), tape_id_off(), tape.far_adj,
                        tape.next_id - tape.first_temp, tape.first_temp, tape.wm1, outv.id, seed, T,
                        (int) L.cap_slots + kGuardRows, L.sweep_prefetch);
            }
        }
        sum += outv.value;
        if (L.out_values != nullptr && mine) L.out_values[out_index] = outv.value;
        tape_rewind(tape);
    }
    if (MODE == kAccWarp) param_flush(T.warp_acc, pend);

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

    /* ---- second stage in the same launch: the CTA that finishes last sums the rows of all CTAs ---- */
    if (L.ticket != nullptr) {
        __shared__ unsigned last;
        __threadfence();
        __syncthreads();
        if (tid == 0) last = atomicAdd(L.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (last) {
            __threadfence();
            finish_evaluation(L, ncols, PR, red);
        }
    }
}

#endif /* AD4CL_TAPE_STATIC */

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
#define AD4CL_MIN_BLOCKS 1 /* x AD4CL_MAX_BLOCK = 768 threads per SM = 80 registers, whether launched as 3 x 256 or 1 x 768:
                              measured faster than 1024 threads (64 registers spill and rematerialise) */
#endif
#ifndef AD4CL_ENTRY_SUFFIX
#define AD4CL_ENTRY_SUFFIX /* the two ahead-of-time builds of one kernel need distinct symbols */
#endif
#define AD4CL_CAT4_(a, b, c, d) a##b##c##d
#define AD4CL_CAT4(a, b, c, d) AD4CL_CAT4_(a, b, c, d)
/* one __global__ function per accumulator mode, so a launch only brings the code of
   the sweep it uses into the instruction cache: NAME__entry<suffix>_m0 / _m1 / _m2 */
#define AD4CL_ENTRY_NAME(NAME, MODE) AD4CL_CAT4(NAME, __entry, AD4CL_ENTRY_SUFFIX, _m##MODE)
/*
 * AD4CL_STATIC_ENTRY(NAME, (typed user parameters), (call arguments)): the register-tier entry
 * `NAME__entry<suffix>_s` (needs -DAD4CL_TAPE_STATIC=1 -DAD4CL_P=<independents>).  Same parameter list as
 * the dynamic entry, so the shim marshals both alike; inside the call-argument list `th` is the local array
 * of the independents (pass `th + first_id` where the dynamic entry forwards an ad_variable pointer) and
 * integer arguments that bound loops should be written as literals.
 */
#if AD4CL_TAPE_STATIC
#define AD4CL_STATIC_ENTRY_NAME(NAME) AD4CL_CAT4(NAME, __entry, AD4CL_ENTRY_SUFFIX, _s)
#define AD4CL_STATIC_ENTRY(NAME, PARAMS, ARGS)                                             \
    extern "C" __global__ void __launch_bounds__(AD4CL_STATIC_MAX_BLOCK)                    \
    AD4CL_STATIC_ENTRY_NAME(NAME)(const ad4cl::LaunchInfo L, AD4CL_UNPAREN PARAMS) {        \
        auto body = [&](struct ad_gradient_structure* gs, struct ad_variable* out,         \
                struct ad_variable* th) __attribute__((always_inline)) {                   \
            struct ad_entry* gradient_stack = nullptr;                                     \
            (void) gradient_stack; (void) th;                                              \
            NAME(AD4CL_UNPAREN ARGS);                                                      \
        };                                                                                 \
        if (L.recording == 1) ad4cl::run_objective_static<1>(L, body);                     \
        else ad4cl::run_objective_static<0>(L, body);                                      \
    }
#endif
#if !AD4CL_TAPE_STATIC
/* one __global__ function per (recorder policy, accumulator mode), so a launch only brings the code it uses
   into the instruction cache: NAME__entry<suffix>_m0/_m1/_m2 (LANE kernel) and _u0/_u1/_u2 (UNIFORM kernel;
   not built for kernels with work-group barriers, which cannot be re-run by a single warp) */
#define AD4CL_ENTRY_NAME2(NAME, TAG, MODE) AD4CL_CAT4(NAME, __entry, AD4CL_ENTRY_SUFFIX, _##TAG##MODE)
#define AD4CL_OBJECTIVE_ENTRY_MODE(NAME, MODE, POLICY, TAG, PARAMS, ARGS)                  \
    extern "C" __global__ void __launch_bounds__(AD4CL_MAX_BLOCK, AD4CL_MIN_BLOCKS)        \
    AD4CL_ENTRY_NAME2(NAME, TAG, MODE)(const ad4cl::LaunchInfo L, AD4CL_UNPAREN PARAMS) {  \
        /* always_inline: the body may be instantiated twice (uniform and general recorder), and only     \
           inlined copies let the compiler keep the tape cursor in registers and fold the recorder mode */ \
        ad4cl::run_objective<MODE, POLICY>(L, [&](struct ad_gradient_structure* gs, struct ad_variable* out) \
                __attribute__((always_inline)) {                                           \
            struct ad_entry* gradient_stack = nullptr;                                     \
            (void) gradient_stack;                                                         \
            NAME(AD4CL_UNPAREN ARGS);                                                      \
        });                                                                                \
    }
#define AD4CL_OBJECTIVE_ENTRY(NAME, PARAMS, ARGS)                                         \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 0, ad4cl::kRecordGeneral, m, PARAMS, ARGS)            \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 1, ad4cl::kRecordGeneral, m, PARAMS, ARGS)            \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 2, ad4cl::kRecordGeneral, m, PARAMS, ARGS)            \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 0, ad4cl::kRecordUniform, u, PARAMS, ARGS)            \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 1, ad4cl::kRecordUniform, u, PARAMS, ARGS)            \
    AD4CL_OBJECTIVE_ENTRY_MODE(NAME, 2, ad4cl::kRecordUniform, u, PARAMS, ARGS)
#endif /* !AD4CL_TAPE_STATIC */
