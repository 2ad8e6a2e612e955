/*
 * ad4cl_tape.cuh -- the per-thread Jacobian tape of the B200 AD runtime.
 *
 * What the reference does (ad.cl:174-198 and every other operator): reserve a
 * slot on ONE global stack with atomic_inc, write a 40-byte ad_entry
 * {dx,id}[2],id,size (ad.cl:75-79) at a 40-byte stride per work-item, and
 * leave the adjoint sweep to a serial host loop (ad4cl.h:775-802).
 *
 * What this file does instead (sm_100a):
 *   - every thread owns a private tape; a warp's 32 tapes are interleaved into
 *     ONE contiguous LIFO stream of 384-byte rows
 *         row s = [ dx of slot s, lanes 0..31 : 256 B ][ id word, lanes 0..31 : 128 B ]
 *     so a warp-wide push is two fully coalesced stores (2 + 1 cache lines) and
 *     the forward pass writes / the reverse pass reads strictly sequential
 *     memory per warp.  The stream lives in an HBM arena sized by RESIDENT
 *     warps (not by work-items) or in shared memory.
 *   - a slot is one (partial, operand id) pair: 12 bytes.  An operator with k
 *     variable operands pushes k slots; the result id is implicit (= number of
 *     independents + the index of the operator's LAST slot on this thread's
 *     tape), and that last slot carries END in the top bit of its id word.
 *     That is the 12k-bytes-per-op record of SURVEY.md 8(d).
 *   - the reverse sweep runs in the same kernel, right after the work-item's
 *     forward pass.  Adjoints of temporaries live in a shared-memory ring of W
 *     entries per thread indexed by slot index (operands are almost always
 *     recent); an operand W or more slots old is detected at record time,
 *     flagged FAR in its own slot and HAS_FAR in the END slot of the operator
 *     that produced it, and its adjoint goes through a "far" plane in the arena.
 *   - adjoints of the independents (ids below first_temp) never touch the
 *     tape: they go to the accumulator the launch selected (ad4cl_entry.cuh).
 */
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

namespace ad4cl {

/* id word of a slot:
 *   [31] END      last slot of an operator (its index is the result's position)
 *   [30] NOP      slot carries no partial (leaf variable)
 *   [29] PARAM    operand is an independent; payload = its id
 *   [28] FAR      operand is a temporary W or more slots old; payload = its position
 *   [27] HAS_FAR  (END slots) the result of this operator received FAR contributions
 *   [26:0] payload
 */
constexpr int kEndFlag = (int) 0x80000000u;
constexpr int kNopFlag = 0x40000000;
constexpr int kParamFlag = 0x20000000;
constexpr int kFarFlag = 0x10000000;
constexpr int kHasFarFlag = 0x08000000;
constexpr int kIdMask = 0x07FFFFFF;
constexpr int kNoId = -1;                   /* "no output written" */

constexpr int kRowBytes = 384;              /* 32 lanes x (8 B partial + 4 B id word) */
constexpr int kRowIdOffset = 256;
constexpr int kFarRowBytes = 256;           /* far adjoints: 32 lanes x 8 B per slot index */
constexpr int kGuardRows = 2;               /* an operator has at most 2 slots */

enum Status : int {
    kOk = 0,
    kErrTapeOverflow = 1,   /* more slots than the arena column holds */
    kErrWindow = 2,         /* far operand met while the far plane is disabled */
    kErrBounds = 8,         /* AD4CL_DEBUG_BOUNDS build: an arena access left this thread's column */
};

/* Self-checking build (make EXTRA=-DAD4CL_DEBUG_BOUNDS=1): every arena access is range-checked
   against the thread's own column and skipped + reported instead of executed.  compute-sanitizer
   is not available on the GPU pool this was developed on; the GPU test-suite is run against this
   build instead. */
#ifndef AD4CL_DEBUG_BOUNDS
#define AD4CL_DEBUG_BOUNDS 0
#endif

struct ThreadTape {
    char* cur;            /* address of this lane's partial in the next free row */
    char* begin;          /* row 0 */
    int cap_id;           /* first_temp + capacity in slots; kGuardRows guard rows follow the column */
    int id_off;           /* byte offset from a row's partial to its id word, this lane */
    int next_id;          /* first_temp + index of the next free row */
    int first_temp;       /* number of independents: ids >= first_temp are tape positions */
    int wm1;              /* W - 1, W a power of two */
    int nops;             /* operators recorded by the current work-item (statistics) */
    char* far_adj;        /* this lane's far-adjoint column (row stride kFarRowBytes), or null */
#if AD4CL_DEBUG_BOUNDS
    int cap_rows;         /* rows of this column including the guard rows */
#endif
    int status;           /* sticky Status bits */
    long long total_ops;  /* statistics over all work-items of this thread */
    long long total_slots;
};

/* Bytes one warp needs in the arena for cap_slots slots. */
__host__ __device__ inline size_t warp_region_bytes(unsigned cap_slots, bool far) {
    size_t b = (size_t) (cap_slots + kGuardRows) * kRowBytes;
    if (far) b += (size_t) (cap_slots + kGuardRows) * kFarRowBytes;
    return (b + 127) & ~(size_t) 127;
}

__device__ __forceinline__ void tape_bind(ThreadTape& t, char* warp_region, unsigned cap_slots,
        int first_temp, int window, bool far) {
    const int lane = threadIdx.x & 31;
    t.begin = warp_region + lane * 8;
    t.cur = t.begin;
    t.cap_id = first_temp + (int) cap_slots;
    t.id_off = kRowIdOffset - lane * 4;
    t.next_id = first_temp;
    t.first_temp = first_temp;
    t.wm1 = window - 1;
    t.nops = 0;
    t.far_adj = far ? warp_region + (size_t) (cap_slots + kGuardRows) * kRowBytes + lane * 8 : nullptr;
#if AD4CL_DEBUG_BOUNDS
    t.cap_rows = (int) cap_slots + kGuardRows;
#endif
    t.status = kOk;
    t.total_ops = 0;
    t.total_slots = 0;
}

/* An operand W or more slots old cannot use the ring: mark the END slot of the
   operator that produced it.  The far plane is all zeros between work-items
   (the sweep restores it), so nothing has to be cleared here, and the OR needs
   no return value: it is issued as a fire-and-forget reduction.  Rare, so it is
   kept OUT of line and takes its inputs by value (the tape cursor stays in
   registers): ~50 push sites per objective would otherwise each carry a copy.
   (A predicated `red` in inline PTX instead of the call was measured slower:
   ptxas wraps it in a branch anyway.) */
static __device__ __noinline__ int tape_note_far(char* begin, int id_off, const char* far_adj, int pos, int nslots) {
    if (far_adj == nullptr) return kErrWindow;
#if AD4CL_DEBUG_BOUNDS
    if (pos < 0 || pos >= nslots) return kErrBounds;
#endif
    (void) nslots;
    atomicOr(reinterpret_cast<int*> (begin + (size_t) pos * kRowBytes + id_off), kHasFarFlag);
    return kOk;
}

/* Push one (partial, operand) slot.  `flags` is 0, kEndFlag or kEndFlag|kNopFlag.
   The operand is classified here (independent / recent temporary / far
   temporary) so the sweep only tests flag bits.  The store is unconditional
   (guard rows absorb the last operator of an overflowing tape); capacity is
   checked once per operator in tape_close_op. */
__device__ __forceinline__ void tape_push(ThreadTape* t, double dx, int id, int flags) {
    const int pos = id - t->first_temp;
    /* pos < 0: an independent.  Otherwise a temporary; the operator being recorded ends at
       this row or the next one, so the operand is out of the ring's reach when
       id + (W - 1) <= next_id. */
    int word = (pos < 0 ? (id | kParamFlag) : pos) | flags;
    if (pos >= 0 && id + t->wm1 <= t->next_id && !(flags & kNopFlag)) {
        word |= kFarFlag;
        t->status |= tape_note_far(t->begin, t->id_off, t->far_adj, pos, t->next_id - t->first_temp);
    }
#if AD4CL_DEBUG_BOUNDS
    if (t->cur < t->begin || t->cur >= t->begin + (size_t) t->cap_rows * kRowBytes
            || (size_t) (t->cur - t->begin) != (size_t) (t->next_id - t->first_temp) * kRowBytes) {
        t->status |= kErrBounds;
        return;
    }
#endif
    *reinterpret_cast<double*> (t->cur) = dx;
    *reinterpret_cast<int*> (t->cur + t->id_off) = word;
    t->cur += kRowBytes;
    t->next_id++;
}

/* Close the operator whose `k` slots were just pushed; returns its result id
   (the position of its END slot).  An operator that does not fit sets the sticky
   overflow bit and restarts the column at row 0: every later access stays inside
   this thread's own column, the sweep of the work-item is skipped
   (ad4cl_entry.cuh) and the host reports the error. */
__device__ __forceinline__ int tape_close_op(ThreadTape* t, int k) {
    (void) k;
    if (t->next_id > t->cap_id) {
        t->status |= kErrTapeOverflow;
        t->cur = t->begin;
        t->next_id = t->first_temp;
    }
    t->nops++;
    return t->next_id - 1;
}

/* Forget the current work-item's tape (the arena column is reused). */
__device__ __forceinline__ void tape_rewind(ThreadTape& t) {
    t.total_ops += t.nops;
    t.total_slots += t.next_id - t.first_temp;
    t.cur = t.begin;
    t.next_id = t.first_temp;
    t.nops = 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, off);
    }
    return v;
}

/* Where adjoints of the independents go during the sweep. */
enum AccMode : int {
    kAccThread = 0, /* per-thread accumulators in shared memory, tree-reduced at the end */
    kAccWarp = 1,   /* warp-shuffle reduce per contribution, per-warp accumulators */
    kAccGlobal = 2, /* fp64 RED into a global gradient array (many independents) */
};

/* shared-memory accesses by 32-bit shared-window address: no generic-pointer
   conversion in the sweep's inner loop */
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    /* volatile: the compiler must keep the result in a register instead of re-deriving
       the shared window base (5 instructions) at every use inside the sweep loop */
    unsigned a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" : : "r"(addr), "d"(v) : "memory");
}

struct SweepTarget {
    unsigned ring;        /* shared address of this thread's ring: W + 1 doubles per thread (the pad keeps
                             the lanes of a warp on distinct banks), entry r at ring + 8 r */
    unsigned stride;      /* kAccThread: bytes between consecutive parameters of a column (block size * 8) */
    unsigned thread_acc;  /* kAccThread: shared address of this thread's column, parameter p at + p * stride */
    unsigned warp_acc;    /* kAccWarp: shared address of this warp's row of P accumulators */
    double* global_acc;   /* kAccGlobal: gradient array indexed by id */
};

/* kAccWarp: all 32 lanes arrive here together; lanes are grouped by parameter id in
   lowest-lane-first order so the summation order is fixed.  Every lane holds the group's sum
   after the butterfly and every lane performs the same update of the warp's accumulator (same
   address, same value: one shared-memory wavefront).  Inlined at each use: an out-of-line copy
   (call + return per contribution) was measured 5 % slower on catch-at-age. */
__device__ __forceinline__
void warp_param_reduce(unsigned warp_acc, bool is_param, int id, double c) {
    /* precondition (checked by the caller with one vote): at least one lane has is_param */
    unsigned pending = __ballot_sync(0xffffffffu, is_param);
    do {
        const int leader = __ffs(pending) - 1;
        const int lid = __shfl_sync(0xffffffffu, id, leader);
        const bool mine = is_param && id == lid;
        const unsigned members = __ballot_sync(0xffffffffu, mine);
        const double v = warp_sum(mine ? c : 0.0);
        const unsigned a = warp_acc + (unsigned) lid * 8u;
        sts_f64(a, lds_f64(a) + v);
        pending &= ~members;
    } while (pending);
}

template <int MODE>
__device__ __forceinline__ void sweep_param(const SweepTarget& T, bool is_param, int id, double c) {
    if (MODE == kAccThread) {
        if (is_param) {
            const unsigned a = T.thread_acc + (unsigned) id * T.stride;
            sts_f64(a, lds_f64(a) + c);
        }
    } else if (MODE == kAccWarp) {
        if (__any_sync(0xffffffffu, is_param)) warp_param_reduce(T.warp_acc, is_param, id, c);
    } else {
        if (is_param) atomicAdd(&T.global_acc[id], c);
    }
}

/* The sweep loop exists in two bodies selected per batch by a warp vote -- with and without the
   far-plane code -- when that pays: measured +4 % with per-thread accumulators (regression,
   tape_stress: short bodies), -4 % with the warp-shuffle reduction inlined four times
   (catch-at-age), where a single body is kept. */
#ifndef AD4CL_SWEEP_TWO_BODIES
#define AD4CL_SWEEP_TWO_BODIES(MODE) ((MODE) == kAccThread)
#endif
#ifndef AD4CL_PREFETCH_BATCHES
#define AD4CL_PREFETCH_BATCHES 2
#endif
constexpr int kPrefetchBatches = AD4CL_PREFETCH_BATCHES; /* the prefetch runs this many batches below the one being fetched */
#ifndef AD4CL_SWEEP_UNROLL
#define AD4CL_SWEEP_UNROLL 4
#endif
constexpr int kSweepUnroll = AD4CL_SWEEP_UNROLL; /* rows fetched per batch; two batches are in flight */

/*
 * One slot of the sweep.  FAR = false is the body for a batch in which no lane of the warp holds
 * a FAR operand or a HAS_FAR result: it carries no far-plane code at all, so the common path has
 * no branches to skip over it (a taken branch costs the sweep more than the instructions it
 * skips).  Returns false on a failed bounds check (AD4CL_DEBUG_BOUNDS builds only).
 */
template <int MODE, bool FAR>
__device__ __forceinline__ bool sweep_slot(ThreadTape& t, const SweepTarget& T, double& w, int idw, double dx,
        int me, int wmask, unsigned pad_off, char* far_adj) {
    if (idw < 0) { /* END: pop the adjoint of the operator that ends at this slot */
        const unsigned r = T.ring + ((unsigned) (me & wmask) << 3);
        w = lds_f64(r);
        sts_f64(r, 0.0);
    }
    if (FAR && (idw & (kFarFlag | kHasFarFlag))) { /* rare: ONE test for both far cases, so the common path skips one block */
        if (idw & kHasFarFlag) { /* (END slots only) contributions that came through the far plane */
#if AD4CL_DEBUG_BOUNDS
            if (far_adj == nullptr || me < 0 || me >= t.cap_rows || idw >= 0) {
                t.status |= kErrBounds;
                return false;
            }
#endif
            double* fa = reinterpret_cast<double*> (far_adj + (size_t) me * kFarRowBytes);
            w += *fa;
            *fa = 0.0;
        }
        if ((idw & kFarFlag) && far_adj != nullptr) {
#if AD4CL_DEBUG_BOUNDS
            if ((idw & kIdMask) >= me) {
                t.status |= kErrBounds;
                return false;
            }
#endif
            *reinterpret_cast<double*> (far_adj + (size_t) (idw & kIdMask) * kFarRowBytes) += w * dx;
        }
    }
    const int payload = idw & kIdMask;
    const double c = w * dx;
    const int cls = idw & (kNopFlag | kParamFlag | kFarFlag);
#if AD4CL_DEBUG_BOUNDS
    if ((cls == 0 || cls == kFarFlag) && (payload < 0 || payload >= me)) { /* operands precede their use */
        t.status |= kErrBounds;
        return false;
    }
    if (cls == kParamFlag && payload >= t.first_temp) {
        t.status |= kErrBounds;
        return false;
    }
    if (!FAR && (cls == kFarFlag || (idw < 0 && (idw & kHasFarFlag)))) { /* the batch vote let a far slot through */
        t.status |= kErrBounds;
        return false;
    }
#endif
    {   /* Branch-free scatter: a recent temporary adds to its ring entry, with per-thread
           accumulators (kAccThread) an independent adds to its accumulator, and a slot of any
           other class adds to the pad entry of this thread's ring (index W), which nothing reads. */
        unsigned a = T.ring + (cls == 0 ? ((unsigned) (payload & wmask) << 3) : pad_off);
        if (MODE == kAccThread) a = cls == kParamFlag ? T.thread_acc + (unsigned) payload * T.stride : a;
        sts_f64(a, lds_f64(a) + c);
    }
    if (MODE != kAccThread) sweep_param<MODE>(T, cls == kParamFlag, payload, c);
    return true;
}

/*
 * Reverse sweep of the current work-item's tape, seeded with 1 at `out_id`.
 * Restates the loop of ad4cl.h:786-797 (w = g[id]; g[id] = 0; g[operand] +=
 * w * dx) on the slot stream: the adjoint of an operator is popped from the
 * ring when its END slot is met, and each slot scatters w * dx to the ring
 * (recent temporaries), the far plane (old temporaries) or the accumulator
 * (independents).  All lanes of the warp iterate together (tapes of
 * neighbouring work-items normally have identical shape).  The stream is read
 * kSweepUnroll rows at a time, one batch ahead of the batch being processed,
 * so the L2/HBM latency of a row is overlapped with the adjoint arithmetic of
 * the previous ones; `prefetch` additionally pulls the batch after that into L2.
 */
template <int MODE>
__device__ __forceinline__ void tape_sweep(ThreadTape& t, int out_id, const SweepTarget& T, bool prefetch) {
    constexpr int U = kSweepUnroll;
    const int wmask = t.wm1;
    const unsigned pad_off = (unsigned) (wmask + 1) << 3;
    const int id_off = t.id_off;
    char* const far_adj = t.far_adj;
    int rows = t.next_id - t.first_temp;

    /* seed d(out) = 1 */
    const int seed_pos = out_id - t.first_temp;
    sweep_param<MODE>(T, out_id >= 0 && out_id < t.first_temp, out_id, 1.0);
    if (seed_pos >= 0 && seed_pos < rows) {
        /* rows recorded after `out` are dead; drop those that could alias the seed's ring entry */
        if (rows - seed_pos > wmask) rows = seed_pos + wmask;
        sts_f64(T.ring + ((unsigned) (seed_pos & wmask) << 3), 1.0);
    }
    const char* cur = t.begin + (size_t) rows * kRowBytes;

#if AD4CL_DEBUG_BOUNDS
    if (rows < 0 || rows > t.cap_rows) {
        t.status |= kErrBounds;
        return;
    }
#endif
    double w = 0.0;
    /* the first batch takes the rows that do not fill a whole batch */
    double dx_now[U], dx_next[U];
    int id_now[U], id_next[U];
    const int head = rows & (U - 1);
#pragma unroll
    for (int j = 0; j < U; j++) {
        if (j < head) {
            const char* row = cur - (j + 1) * kRowBytes;
            dx_now[j] = *reinterpret_cast<const double*> (row);
            id_now[j] = *reinterpret_cast<const int*> (row + id_off);
        } else {
            dx_now[j] = 0.0;
            id_now[j] = kNopFlag;
        }
    }
    cur -= (size_t) head * kRowBytes;
    int idx = rows;            /* slot j of the batch in flight has index idx - 1 - j */
    int next_idx = rows - head;
    int batches = rows / U;    /* whole batches still in memory */
    bool more = true;
    while (__any_sync(0xffffffffu, more)) {
        const bool fetch = batches > 0;
        if (fetch) {
#pragma unroll
            for (int j = 0; j < U; j++) {
                const char* row = cur - (j + 1) * kRowBytes;
                dx_next[j] = *reinterpret_cast<const double*> (row);
                id_next[j] = *reinterpret_cast<const int*> (row + id_off);
            }
            cur -= U * kRowBytes;
            --batches;
            /* The batch just requested is consumed one iteration from now; when the kernel is
               latency-bound that is less than an HBM round trip (ncu, catch-at-age: 20 % of all stall
               samples sat on the first use of a fetched row).  With `prefetch` the batch AFTER the
               next one is pulled into L2 now, without registers: the warp's stream is contiguous,
               a batch is U * 384 B = 3 U cache lines, and lanes 0 .. 3U-1 each touch one of them
               (generic prefetch: a no-op when the tape lives in shared memory).  It is a launch
               option because it costs a bandwidth-bound sweep up to 5 % (tape_stress): the shim
               times both settings on the first evaluation (LaunchInfo.sweep_prefetch). */
            if (prefetch && batches >= kPrefetchBatches && id_off > kRowIdOffset - 4 * 3 * U) {
                const int lane = (kRowIdOffset - id_off) >> 2;
                const char* line = cur - lane * 8 - kPrefetchBatches * U * kRowBytes + lane * 128;
                asm volatile("prefetch.L2 [%0];" : : "l"(line));
            }
        } else {
#pragma unroll
            for (int j = 0; j < U; j++) {
                dx_next[j] = 0.0;
                id_next[j] = kNopFlag;
            }
        }
#if AD4CL_DEBUG_BOUNDS
        if (fetch && cur < t.begin) {
            t.status |= kErrBounds;
            return;
        }
#endif
        int flags = 0;
#pragma unroll
        for (int j = 0; j < U; j++) flags |= id_now[j];
        if (!AD4CL_SWEEP_TWO_BODIES(MODE) || __any_sync(0xffffffffu, (flags & (kFarFlag | kHasFarFlag)) != 0)) {
#pragma unroll
            for (int j = 0; j < U; j++)
                if (!sweep_slot<MODE, true>(t, T, w, id_now[j], dx_now[j], idx - 1 - j, wmask, pad_off, far_adj)) return;
        } else {
#pragma unroll
            for (int j = 0; j < U; j++)
                if (!sweep_slot<MODE, false>(t, T, w, id_now[j], dx_now[j], idx - 1 - j, wmask, pad_off, far_adj)) return;
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            dx_now[j] = dx_next[j];
            id_now[j] = id_next[j];
        }
        idx = next_idx;
        next_idx -= U;
        more = fetch;
    }
}
} // namespace ad4cl
