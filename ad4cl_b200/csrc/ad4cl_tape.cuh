/*
 * ad4cl_tape.cuh -- the per-thread Jacobian tape of the B200 AD runtime.
 *
 * What the reference does (ad.cl:174-198 and every other operator): reserve a
 * slot on ONE global stack with atomic_inc, write a 40-byte ad_entry
 * {dx,id}[2],id,size (ad.cl:75-79) at a 40-byte stride per work-item, and
 * leave the adjoint sweep to a serial host loop (ad4cl.h:775-802).
 *
 * What this file does instead (sm_100a):
 *   - every thread owns a private tape; a slot is one (partial, operand id)
 *     pair.  An operator with k variable operands pushes k slots; the result id
 *     is implicit (= number of independents + the index of the operator's LAST
 *     slot on this thread's tape), and that last slot carries END.
 *   - the 32 tapes of a warp are stored as ONE stream per warp, split by what
 *     is the same in all lanes and what is not (round 2; round 1 stored
 *     384-byte rows of 32 partials + 32 id words):
 *         header plane   8 B per row, ONE copy per warp: {id word, cell offset}.
 *                        SIMT lanes record the same operator sequence, so the
 *                        structure of the tape is warp-uniform; a push checks
 *                        that with one MATCH.ALL over (row, id word).
 *         cell plane     the partials: cell k of lane l at k * 256 + 8 l.  Only
 *                        slots whose partial is not a compile-time +1 / -1 own a
 *                        cell (plus / minus / += : ~46 % of catch-at-age's slots);
 *                        unit slots point at two constant cells (+1.0, -1.0).
 *         lane-word plane  128 B per row, touched only by rows that are NOT
 *                        warp-uniform (lane-dependent parameter ids, divergent
 *                        control flow): the header then says so and every lane
 *                        keeps its own id word.
 *     A uniform row costs 8 B + 256 B (or 8 B) instead of 384 B, its id word is
 *     classified once per warp on uniform branches, and an independent's
 *     adjoint contribution is ONE warp_sum (no ballot / leader loop).
 *   - the reverse sweep runs in the same kernel, right after the work-item's
 *     forward pass.  Adjoints of temporaries live in a shared-memory ring of W
 *     entries per thread indexed by slot index; an operand that is the result
 *     of the immediately preceding operator is flagged CHAIN at record time and
 *     its contribution is carried in a register to that operator's pop instead
 *     of going through the ring; an operand W or more slots old is flagged FAR
 *     (and HAS_FAR on the END slot of its producer) and goes through a "far"
 *     plane in the arena.
 *   - two sweep loops: tape_sweep_uniform (whole work-item recorded uniformly:
 *     warp-synchronous, headers broadcast, software-pipelined cell loads) and
 *     tape_sweep_general (anything else: every lane walks its own rows).
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
 *   [30] NOP      slot carries no partial (leaf variable / constant operand)
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
/* "no output written" / not a variable.  INT_MIN, not -1: as an operand it fails the push's one unsigned
   reach test for every row index (ad_plus_eq_g on the virtual out before it was assigned) and is then
   recorded as a constant by tape_classify_rare. */
constexpr int kNoId = (int) 0x80000000u;

/* second header word (uniform recorder): byte offset of the slot's cell in the lane's cell column */

constexpr int kHdrBytes = 8;
constexpr int kRowBytes = 384;              /* general recorder: 32 lanes x (8 B partial + 4 B id word) */
constexpr int kRowIdOffset = 256;
constexpr int kCellRowBytes = 256;          /* 32 lanes x 8 B */
constexpr int kLaneWordRowBytes = 128;      /* 32 lanes x 4 B */
constexpr int kFarRowBytes = 256;           /* far adjoints: 32 lanes x 8 B per slot index */
constexpr int kUnitCells = 2;               /* cell 0 holds +1.0, cell 1 holds -1.0 */
constexpr int kGuardRows = 4;               /* an operator has at most 2 slots; the sweep pads to a multiple of 4 rows */
constexpr int kSweepUnroll = 4;             /* rows per batch of both sweep loops */

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
/* A/B switch for the general sweep: 0 removes the per-batch vote and the warp-uniform batch body (every
   lane always classifies its own id word, as in round 1) */
#ifndef AD4CL_LANE_UNIFORM_BATCHES
#define AD4CL_LANE_UNIFORM_BATCHES 0
#endif

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

/* Planes of one warp's region for cap_slots slots. */
struct RegionLayout {
    unsigned long long off_cells, off_lane_words, off_far, bytes;
};

__host__ __device__ inline RegionLayout region_layout(unsigned cap_slots, bool far) {
    const unsigned long long rows = (unsigned long long) cap_slots + kGuardRows;
    RegionLayout r;
    r.off_cells = (rows * kHdrBytes + 127) & ~127ull;
    r.off_lane_words = r.off_cells + (rows + kUnitCells) * kCellRowBytes;
    r.off_far = r.off_lane_words + rows * kLaneWordRowBytes;
    r.bytes = r.off_far + (far ? rows * kFarRowBytes : 0);
    r.bytes = (r.bytes + 127) & ~127ull;
    return r;
}

/* Bytes one warp needs in the arena for cap_slots slots. */
__host__ __device__ inline size_t warp_region_bytes(unsigned cap_slots, bool far) {
    return (size_t) region_layout(cap_slots, far).bytes;
}


#ifndef AD4CL_TAPE_STATIC
#define AD4CL_TAPE_STATIC 0
#endif
/* 1: the uniform recorder checks warp-uniformity with per-lane hashes compared once per work-item instead
   of a MATCH.ALL vote per push (exact vs 2^-64; measured variant, see profiles/README.md).  Default 0. */
#ifndef AD4CL_PIN_TAPE
#define AD4CL_PIN_TAPE 1
#endif
#ifndef AD4CL_UNIFORM_HASH
#define AD4CL_UNIFORM_HASH 1
#endif


#if AD4CL_TAPE_STATIC
/*
 * ---------------------------------------------------------------------------------------------
 * REGISTER TIER (-DAD4CL_TAPE_STATIC=1 -DAD4CL_P=<independents>): the tape of a short straight-line
 * objective as two small per-thread arrays indexed by the slot counter.  The entry (ad4cl_entry.cuh,
 * run_objective_static) hands the user kernel LOCAL copies of the independents whose ids are
 * literals and passes the kernel's integer arguments as literals, so after inlining and full
 * unrolling every slot index, id word and adjoint index is a compile-time constant: the arrays are
 * scalarised into registers (no LDL/STL, no arena, no shared-memory ring), the sweep is straight-line
 * code, unit partials fold away, and the independents' adjoints are per-thread registers.  This is
 * what the reference's per-work-item private stacks were meant for (ad.cl:51-57, 91-97) and what a
 * 4-operator kernel like test/admb/simple.cl:123-137 wants.  A kernel whose indices do NOT fold
 * still computes correctly (dynamic indexing lands in local memory); the shim only selects this
 * tier when the compiled entry uses no local memory (shim.cu, prepare()).
 * ---------------------------------------------------------------------------------------------
 */
#ifndef AD4CL_STATIC_SLOTS
#define AD4CL_STATIC_SLOTS 64
#endif
#ifndef AD4CL_P
#error "AD4CL_TAPE_STATIC needs -DAD4CL_P=<number of independents>"
#endif

struct ThreadTape {
    double dx[AD4CL_STATIC_SLOTS];
    int word[AD4CL_STATIC_SLOTS];
    int next_id;          /* first_temp + slots recorded */
    int first_temp;       /* = AD4CL_P */
    int nops;
    int status;
    long long total_ops, total_slots;
};

__device__ __forceinline__ void tape_bind_static(ThreadTape& t) {
    t.next_id = AD4CL_P;
    t.first_temp = AD4CL_P;
    t.nops = 0;
    t.status = kOk;
    t.total_ops = 0;
    t.total_slots = 0;
}

template <int UNIT>
__device__ __forceinline__ void tape_push(ThreadTape* t, double dx, int id, int flags) {
    const int n = t->next_id - AD4CL_P;
    int word;
    if (flags & kNopFlag) word = flags;
    else if ((unsigned) id < (unsigned) AD4CL_P) word = id | kParamFlag | flags;
    else if ((unsigned) (id - AD4CL_P) < (unsigned) n) word = (id - AD4CL_P) | flags;
    else word = flags | kNopFlag; /* not a variable of this tape: a constant operand */
    if (n < AD4CL_STATIC_SLOTS) {
        t->dx[n] = UNIT == 0 ? dx : (UNIT > 0 ? 1.0 : -1.0);
        t->word[n] = word;
    } else {
        t->status |= kErrTapeOverflow;
    }
    t->next_id++;
}

template <int UA, int UB>
__device__ __forceinline__ void tape_push2(ThreadTape* t, double dx0, int id0, double dx1, int id1) {
    tape_push<UA>(t, dx0, id0, 0);
    tape_push<UB>(t, dx1, id1, kEndFlag);
}

__device__ __forceinline__ int tape_close_op(ThreadTape* t) {
    t->nops++;
    return t->next_id - 1;
}

__device__ __forceinline__ void tape_rewind(ThreadTape& t) {
    t.total_ops += t.nops;
    t.total_slots += t.next_id - AD4CL_P;
    t.next_id = AD4CL_P;
    t.nops = 0;
}

__device__ __forceinline__ void tape_read_row(const ThreadTape* t, int row, double* dx, int* word) {
    *dx = t->dx[row];
    *word = t->word[row];
}

/* Compile-time-indexed access to the small per-thread arrays: template recursion instead of loops, so
   every index is a constant as soon as the functions are inlined (no dependence on when the compiler
   unrolls), and a run-time index costs a chain of predicated operations instead of local memory. */
template <int N>
struct StaticIndex {
    /* a[i] += c for the one i in [0, N) that equals idx */
    static __device__ __forceinline__ void add(double* a, int idx, double c) {
        if (idx == N - 1) a[N - 1] += c;
        StaticIndex<N - 1>::add(a, idx, c);
    }
    static __device__ __forceinline__ void set(double* a, int idx, double c) {
        if (idx == N - 1) a[N - 1] = c;
        StaticIndex<N - 1>::set(a, idx, c);
    }
    static __device__ __forceinline__ void zero(double* a) {
        a[N - 1] = 0.0;
        StaticIndex<N - 1>::zero(a);
    }
};
template <>
struct StaticIndex<0> {
    static __device__ __forceinline__ void add(double*, int, double) {}
    static __device__ __forceinline__ void set(double*, int, double) {}
    static __device__ __forceinline__ void zero(double*) {}
};

/* rows S-1 .. 0 of the reverse sweep */
template <int S>
struct StaticSweep {
    static __device__ __forceinline__ void run(const ThreadTape& t, int n, double& w, double* adj, double* g) {
        constexpr int s = S - 1;
        if (s < n) {
            const int wd = t.word[s];
            if (wd < 0) w = adj[s];
            const double c = w * t.dx[s];
            if (!(wd & kNopFlag)) {
                const int payload = wd & kIdMask;
                if (wd & kParamFlag) StaticIndex<(AD4CL_P > 0 ? AD4CL_P : 1)>::add(g, payload, c);
                else StaticIndex<s>::add(adj, payload, c); /* operands precede their use */
            }
        }
        StaticSweep<S - 1>::run(t, n, w, adj, g);
    }
};
template <>
struct StaticSweep<0> {
    static __device__ __forceinline__ void run(const ThreadTape&, int, double&, double*, double*) {}
};

/* Reverse sweep of the register tape into per-thread accumulators g[AD4CL_P] (ad4cl.h:786-797 on the
   slot stream).  With compile-time id words every adj[] / g[] index is a constant. */
__device__ __forceinline__ void tape_sweep_static(const ThreadTape& t, int out_id, double (&g)[AD4CL_P > 0 ? AD4CL_P : 1], double seed) {
    const int n = t.next_id - AD4CL_P;
    double adj[AD4CL_STATIC_SLOTS];
    StaticIndex<AD4CL_STATIC_SLOTS>::zero(adj);
    if ((unsigned) out_id < (unsigned) AD4CL_P) StaticIndex<(AD4CL_P > 0 ? AD4CL_P : 1)>::add(g, out_id, seed);
    else if (out_id - AD4CL_P < n) StaticIndex<AD4CL_STATIC_SLOTS>::set(adj, out_id - AD4CL_P, seed);
    double w = 0.0;
    StaticSweep<AD4CL_STATIC_SLOTS>::run(t, n, w, adj, g);
}

#else /* !AD4CL_TAPE_STATIC: the arena / shared-memory tape */

/*
 * Two recorders share the planes of a warp's region; which one a push uses is `mode`, a literal in each
 * of the two inlined copies of the user kernel (ad4cl_entry.cuh), so a push contains only its own code.
 *
 *   kRecordUniform  OPTIMISTIC: assumes all 32 lanes push the same id word at the same row and the
 *                   operand is an independent or a recent temporary.  One header per warp, a cell only
 *                   for a non-unit partial.  Nothing is branched on: whatever breaks the assumption
 *                   (partial mask, lane-dependent id word, far or foreign operand, full column) only
 *                   sets the sticky `bad` flag, and the work-item is then recorded AGAIN by the general
 *                   recorder.  Re-running the user kernel is safe because its only effects are the tape
 *                   and the virtual out (kernels with work-group barriers are built general-only).
 *   kRecordGeneral  the round-1 stream: every lane keeps its own id word and its own partial, one
 *                   384-byte row per slot [partials of the 32 lanes : 256 B][id words : 128 B], so a push
 *                   is two coalesced stores at one running pointer: any control flow, any ids.  Its sweep
 *                   (tape_sweep_general) still notices, batch by batch, when the 32 id words agree, and then
 *                   classifies on warp-uniform branches and reduces independents' contributions in pairs.
 */
constexpr int kRecordUniform = 0;
constexpr int kRecordGeneral = 1;

struct ThreadTape {
    int mode;             /* kRecordUniform / kRecordGeneral */
    /* ONE cursor, next_id, addresses every plane: row r = next_id - first_temp has its header at
       hdr + 8 r, its cell at cells + 256 (2 + r) and its general-recorder row at rows + 384 r.  The bases
       kept here are biased by -first_temp rows, so an address is one IMAD.WIDE of next_id (and an
       immediate offset for the later pushes of the same basic block). */
    char* hdr_b;          /* uniform recorder: header plane of this warp (the same address in all lanes), biased */
    char* cells_b;        /* uniform recorder: this lane's cell column, biased (row r's cell, not cell 0) */
    char* rows_b;         /* general recorder: this lane's partials (row stride kRowBytes), biased */
    char* ids_b;          /* general recorder: this lane's id words (row stride kRowBytes), biased */
    char* far_adj;        /* this lane's far-adjoint column (row stride kFarRowBytes), or null */
    int next_id;          /* first_temp + index of the next free row */
    int cap_id;           /* first_temp + capacity in slots + 1: where next_id is clamped (tape_close_op); a
                             work-item that ends with next_id == cap_id overflowed.  kGuardRows rows follow
                             the capacity, so the operator that hits the clamp still writes inside the column */
    int first_temp;       /* number of independents: ids >= first_temp are tape positions */
    int wm1;              /* W - 1, W a power of two >= 4 */
    int nops;             /* operators recorded by the current work-item (statistics) */
    int bad;              /* uniform recorder: sticky "this work-item needs the general recorder" */
    unsigned h1, h2;      /* AD4CL_UNIFORM_HASH: running hashes of this lane's id words (see tape_push) */
    int status;           /* sticky Status bits */
    long long total_ops;  /* statistics over all work-items of this thread */
    long long total_slots;
};

/* the unbiased bases (sweeps, classification, export: once per work-item, not per push) */
__device__ __forceinline__ char* tape_hdr(const ThreadTape& t) { return t.hdr_b + (long long) t.first_temp * kHdrBytes; }
__device__ __forceinline__ char* tape_cells(const ThreadTape& t) {
    return t.cells_b + ((long long) t.first_temp - kUnitCells) * kCellRowBytes;
}
__device__ __forceinline__ char* tape_rows(const ThreadTape& t) { return t.rows_b + (long long) t.first_temp * kRowBytes; }
__device__ __forceinline__ int tape_id_off() { return kRowIdOffset - (int) (threadIdx.x & 31) * 4; }

__device__ __forceinline__ void tape_bind(ThreadTape& t, char* warp_region, unsigned cap_slots,
        int first_temp, int window, bool far) {
    const int lane = threadIdx.x & 31;
    const RegionLayout lay = region_layout(cap_slots, far);
    t.mode = kRecordGeneral;
    char* cells = warp_region + lay.off_cells + lane * 8;
    /* the general recorder's rows share the space of the cell plane and the (former) lane-word plane:
       (rows + 2) * 256 + rows * 128 = 512 + rows * 384 bytes */
    char* rows = cells + kUnitCells * kCellRowBytes;
    t.hdr_b = warp_region - (long long) first_temp * kHdrBytes;
    t.cells_b = cells + ((long long) kUnitCells - first_temp) * kCellRowBytes;
    t.rows_b = rows - (long long) first_temp * kRowBytes;
    t.ids_b = t.rows_b + (kRowIdOffset - lane * 4);
    t.far_adj = far ? warp_region + lay.off_far + lane * 8 : nullptr;
    t.next_id = first_temp;
    t.cap_id = first_temp + (int) cap_slots + 1;
    t.first_temp = first_temp;
    t.wm1 = window - 1;
    t.nops = 0;
    t.bad = 0;
    t.h1 = t.h2 = 0u;
    t.status = kOk;
    t.total_ops = 0;
    t.total_slots = 0;
    /* the two constant cells every unit slot of the uniform recorder points at */
    *reinterpret_cast<double*> (cells) = 1.0;
    *reinterpret_cast<double*> (cells + kCellRowBytes) = -1.0;
#if AD4CL_PIN_TAPE
    /* Everything above derives from kernel parameters, and under register pressure the compiler prefers to
       re-derive it from the constant bank at every push (2-4 extra instructions per push, measured).  Opaque
       to the optimiser from here on: these stay in registers. */
    asm volatile("" : "+l"(t.hdr_b), "+l"(t.cells_b), "+l"(t.rows_b), "+l"(t.ids_b));
    asm volatile("" : "+r"(t.cap_id), "+r"(t.first_temp), "+r"(t.wm1));
#endif
}

/* Start recording a work-item (or start over, after a failed optimistic attempt). */
__device__ __forceinline__ void tape_begin(ThreadTape& t, int mode) {
    t.mode = mode;
    t.next_id = t.first_temp;
    t.nops = 0;
    t.bad = 0;
    t.h1 = t.h2 = 0u;
}

/* Did the work-item just recorded run out of capacity?  (next_id stopped at the clamp) */
__device__ __forceinline__ bool tape_overflowed(const ThreadTape& t) { return t.next_id >= t.cap_id; }

/* What the warp votes on after a work-item recorded by the uniform recorder: equal in all 32 lanes (and
   no lane `bad`) <=> the work-item is warp-uniform.  With per-push votes (default) only the seed id is left
   to compare; with AD4CL_UNIFORM_HASH the lanes' running hashes of everything they pushed are compared
   here instead of their id words at every push. */
__device__ __forceinline__ unsigned long long tape_uniform_key(const ThreadTape& t, int out_id) {
#if AD4CL_UNIFORM_HASH
    return ((unsigned long long) (t.h1 ^ (unsigned) out_id) << 32) | t.h2;
#else
    return (unsigned) out_id;
#endif
}

/*
 * General recorder, rare operand classes, out of line and by value (the tape cursor stays in registers):
 *   - a temporary W or more slots old cannot use the ring: FAR, and HAS_FAR on the END slot of the
 *     operator that produced it.  The far plane is all zeros between work-items (the sweep restores it).
 *   - an id that is neither an independent nor a position below the current row is not a variable of
 *     this tape (the virtual out before it is written, an uninitialised ad_variable): recorded as a
 *     constant operand instead of an id the sweep would follow out of the arena.
 * Returns the id word in the low half and Status bits in the high half.
 */
static __device__ __noinline__ unsigned long long tape_classify_rare(char* ids, const char* far_adj,
        int id, int first_temp, int row, int flags) {
    const int pos = id - first_temp;
    if (pos < 0 || pos >= row) return (unsigned) (flags | kNopFlag);
    if (far_adj == nullptr) return ((unsigned long long) kErrWindow << 32) | (unsigned) (flags | kNopFlag);
    atomicOr(reinterpret_cast<int*> (ids + (size_t) pos * kRowBytes), kHasFarFlag);
    return (unsigned) (pos | flags | kFarFlag);
}

/*
 * Push one (partial, operand) slot.  `flags` is 0, kEndFlag or kEndFlag|kNopFlag; UNIT = +1 / -1 says the
 * partial is the compile-time constant +1.0 / -1.0 (dx is then ignored), 0 a general partial.
 *
 * Uniform recorder: the header gets the RAW word id | flags.  Everything a lane does here is the vote
 * (same raw word in all 32 lanes, full mask), two cheap sanity tests (the id names a variable that exists:
 * id < next_id; the column has room) folded into the sticky `bad`, two stores and the cursor bumps.
 * Classifying the operand (independent / recent / far temporary) is NOT done per lane and per push: the
 * structure is warp-uniform, so tape_classify_uniform does it once per ROW, the rows dealt out over the lanes.
 *
 * General recorder: the operand is classified here, per lane.
 */
template <int UNIT>
__device__ __forceinline__ void tape_push(ThreadTape* t, double dx, int id, int flags) {
    /* Both recorders advance their cursors UNCONDITIONALLY by compile-time constants: inside a basic block
       the compiler then addresses consecutive pushes as [cursor + immediate] and updates the cursor once.
       Capacity is checked once per operator (tape_close_op); the guard rows absorb the last operator. */
    if (t->mode == kRecordUniform) {
        const int word = (flags & kNopFlag) ? flags : (id | flags);
        const bool valid = (flags & kNopFlag) || (unsigned) id < (unsigned) t->next_id;
#if AD4CL_UNIFORM_HASH
        /* no vote per push: every lane folds its raw word into two independent multiplicative hashes and the
           warp compares them ONCE per work-item (tape_uniform_key).  A lane that skipped a push, pushed in
           another order or pushed another word ends with different hashes, up to a 2^-64 collision. */
        t->h1 = t->h1 * 0x9E3779B1u + (unsigned) word;
        t->h2 = (t->h2 ^ (unsigned) word) * 0x85EBCA6Bu + 0xC2B2AE35u;
        t->bad |= valid ? 0 : 1;
#else
        const unsigned mask = __activemask();
        int same = 0;
        __match_all_sync(mask, word, &same);
        t->bad |= (!valid || !same || mask != 0xffffffffu) ? 1 : 0;
#endif
#if AD4CL_DEBUG_BOUNDS
        if ((unsigned) (t->next_id - t->first_temp) >= (unsigned) (t->cap_id - t->first_temp - 1 + kGuardRows)) {
            t->status |= kErrBounds;   /* the clamp of tape_close_op keeps the cursor inside the guard rows: never here */
            return;
        }
#endif
        const int aux = UNIT == 0 ? (t->next_id - t->first_temp + kUnitCells) * kCellRowBytes
                                  : (UNIT > 0 ? 0 : kCellRowBytes);
        *reinterpret_cast<int2*> (t->hdr_b + (long long) t->next_id * kHdrBytes) = make_int2(word, aux);
        if (UNIT == 0) *reinterpret_cast<double*> (t->cells_b + (long long) t->next_id * kCellRowBytes) = dx;
        t->next_id++;
    } else {
        const bool is_param = (unsigned) id < (unsigned) t->first_temp;
        /* A temporary is within the ring's reach when it is fewer than W - 1 rows below this one (the
           operator being recorded ends at this row or the next).  One unsigned compare also rejects ids
           at or above the current row and kNoId. */
        const bool reach = (unsigned) (t->next_id - 1 - id) < (unsigned) (t->wm1 - 1);
        int word = (flags & kNopFlag) ? flags : (is_param ? (id | kParamFlag) : (id - t->first_temp)) | flags;
        if (!(flags & kNopFlag) && !is_param && !reach) {
            const unsigned long long r = tape_classify_rare(t->ids_b + (long long) t->first_temp * kRowBytes, t->far_adj,
                    id, t->first_temp, t->next_id - t->first_temp, flags);
            word = (int) (unsigned) r;
            t->status |= (int) (r >> 32);
        }
#if AD4CL_DEBUG_BOUNDS
        if ((unsigned) (t->next_id - t->first_temp) >= (unsigned) (t->cap_id - t->first_temp - 1 + kGuardRows)) {
            t->status |= kErrBounds;
            return;
        }
#endif
        *reinterpret_cast<double*> (t->rows_b + (long long) t->next_id * kRowBytes) = UNIT == 0 ? dx : (UNIT > 0 ? 1.0 : -1.0);
        *reinterpret_cast<int*> (t->ids_b + (long long) t->next_id * kRowBytes) = word;
        t->next_id++;
    }
}

/*
 * Uniform recorder, after the work-item: turn the raw words id | flags of the `rows` headers into the
 * classified id words the sweep reads (PARAM + id; position of a recent temporary; FAR + position, with
 * HAS_FAR on the row that produced the operand).  ONE classification per row for the whole warp, row r by
 * lane r mod 32.
 * Returns non-zero (warp-uniform) when the tape needs the general recorder after all: a far operand with the
 * far plane disabled.  Out of line: runs once per work-item.
 */
static __device__ __noinline__ int tape_classify_uniform(char* hdr, const char* far_adj, int rows, int first_temp, int wm1) {
    int bad = 0;
    for (int r = threadIdx.x & 31; r < rows; r += 32) {
        int* hw = reinterpret_cast<int*> (hdr + (size_t) r * kHdrBytes);
        const int raw = *hw;
        const int flags = raw & (kEndFlag | kNopFlag);
        const int id = raw & kIdMask;
        int word = flags;
        if (!(flags & kNopFlag)) {
            if (id < first_temp) {
                word = id | kParamFlag | flags;
            } else {
                word = (id - first_temp) | flags;
                if (!((unsigned) (first_temp + r - 1 - id) < (unsigned) (wm1 - 1))) { /* beyond the ring's reach */
                    word |= kFarFlag;
                    if (far_adj == nullptr) bad = 1;
                }
            }
        }
        *hw = word;
    }
    __syncwarp();
    /* second pass, after every row has its final word: HAS_FAR on the rows that produced the far operands
       (several far rows, handled by different lanes, may name the same producer: atomic OR) */
    for (int r = threadIdx.x & 31; r < rows; r += 32) {
        const int word = *reinterpret_cast<const int*> (hdr + (size_t) r * kHdrBytes);
        if ((word & (kFarFlag | kNopFlag | kParamFlag)) == kFarFlag)
            atomicOr(reinterpret_cast<int*> (hdr + (size_t) (word & kIdMask) * kHdrBytes), kHasFarFlag);
    }
    __syncwarp();
    return __any_sync(0xffffffffu, bad);
}

/* Both slots of a two-operand operator. */
template <int UA, int UB>
__device__ __forceinline__ void tape_push2(ThreadTape* t, double dx0, int id0, double dx1, int id1) {
    tape_push<UA>(t, dx0, id0, 0);
    tape_push<UB>(t, dx1, id1, kEndFlag);
}

/* Close the operator whose slots were just pushed; returns its result id (the position of its END slot).
   Capacity costs ONE instruction per operator: next_id is clamped at cap_id, so a tape that does not fit
   keeps overwriting the guard rows behind its column -- every access stays inside this thread's own
   column -- and the work-item ends with next_id == cap_id (tape_overflowed): the uniform recorder then
   hands the work-item to the general one, the general one sets the sticky overflow bit (the sweep is
   skipped, the host reports the error). */
__device__ __forceinline__ int tape_close_op(ThreadTape* t) {
    t->next_id = min(t->next_id, t->cap_id);
    t->nops++;
    return t->next_id - 1;
}

/* The work-item is done: fold its counts into the statistics (the arena column is reused). */
__device__ __forceinline__ void tape_rewind(ThreadTape& t) {
    t.total_ops += t.nops;
    t.total_slots += t.next_id - t.first_temp;
    t.next_id = t.first_temp;
    t.nops = 0;
}

/* Debug / export accessor: id word and partial of row `row` of this lane's tape. */
__device__ __forceinline__ void tape_read_row(const ThreadTape* t, int row, double* dx, int* word) {
    if (t->mode == kRecordUniform) {
        const int2 h = *reinterpret_cast<const int2*> (tape_hdr(*t) + (size_t) row * kHdrBytes);
        *word = h.x;
        *dx = *reinterpret_cast<const double*> (tape_cells(*t) + (unsigned) h.y);
    } else {
        *word = *reinterpret_cast<const int*> (tape_rows(*t) + (size_t) row * kRowBytes + tape_id_off());
        *dx = *reinterpret_cast<const double*> (tape_rows(*t) + (size_t) row * kRowBytes);
    }
}

struct SweepTarget {
    unsigned ring;        /* shared address of this thread's ring: W + 1 doubles per thread (the pad keeps
                             the lanes of a warp on distinct banks), entry r at ring + 8 r */
    unsigned stride;      /* kAccThread: bytes between consecutive parameters of a column (block size * 8) */
    unsigned thread_acc;  /* kAccThread: shared address of this thread's column, parameter p at + p * stride */
    unsigned warp_acc;    /* kAccWarp: shared address of this warp's row of P accumulators */
    double* global_acc;   /* kAccGlobal: gradient array indexed by id */
};

/* kAccWarp, lane-dependent ids: all 32 lanes arrive here together; lanes are grouped by parameter id
   in lowest-lane-first order so the summation order is fixed.  Every lane holds the group's sum
   after the butterfly and every lane performs the same update of the warp's accumulator (same
   address, same value: one shared-memory wavefront). */
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

/* per-lane form (general sweep, seed): every lane may hold a different id */
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

/* Seeds d(out) = seed (1 unless the caller supplied per-work-item seeds) and returns the number of rows
   the sweep has to visit (rows recorded after `out` are dead; those that could alias the seed's ring
   entry are dropped). */
template <int MODE>
__device__ __forceinline__ int sweep_seed(int rows, int first_temp, int wmask, int out_id, double seed, const SweepTarget& T) {
    const int seed_pos = out_id - first_temp;
    sweep_param<MODE>(T, out_id >= 0 && out_id < first_temp, out_id, seed);
    if (seed_pos >= 0 && seed_pos < rows) {
        if (rows - seed_pos > wmask) rows = seed_pos + wmask;
        sts_f64(T.ring + ((unsigned) (seed_pos & wmask) << 3), seed);
    }
    return rows;
}

/*
 * kAccWarp in the uniform sweep: contributions to the independents are reduced over the warp TWO AT A
 * TIME.  A contribution waits in `pend`; when the next one (to another independent) arrives, the lower
 * half-warp takes over the pending value and the upper half the new one -- one exchange over lane
 * distance 16 -- and the four remaining butterfly steps reduce both at once: 5 shuffle steps for two
 * sums instead of 10.  A contribution to the SAME independent as the pending one is just added to it.
 * The pending contribution survives across work-items and is flushed when the thread's work is done.
 * The order of the additions is fixed by the tape, so results stay bit-reproducible.
 */
struct PendingParam {
    double c;
    int id;      /* < 0: nothing pending */
};

__device__ __forceinline__ void param_flush(unsigned warp_acc, PendingParam& pend) {
    if (pend.id >= 0) {
        const double v = warp_sum(pend.c);
        const unsigned a = warp_acc + (unsigned) pend.id * 8u;
        sts_f64(a, lds_f64(a) + v);
        pend.id = -1;
    }
}

__device__ __forceinline__ void param_add_paired(unsigned warp_acc, PendingParam& pend, int id, double c) {
    if (pend.id < 0) {
        pend.c = c;
        pend.id = id;
    } else if (pend.id == id) {
        pend.c += c;
    } else {
        const bool lower = (threadIdx.x & 16u) == 0u;
        const double mine = lower ? pend.c : c;      /* the sum this half-warp finishes */
        const double theirs = lower ? c : pend.c;    /* what the other half needs from this lane */
        double v = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        /* lanes 0-15 hold the pending independent's sum, lanes 16-31 the new one's: two addresses, one
           read-modify-write instruction (the ids differ, so the two updates cannot collide) */
        const unsigned a = warp_acc + (unsigned) (lower ? pend.id : id) * 8u;
        sts_f64(a, lds_f64(a) + v);
        pend.id = -1;
    }
}

/* The far cases of one row of the uniform sweep, out of line and by value: an END row whose result received
   contributions through the far plane (HAS_FAR) collects them into w; a FAR operand adds w * dx to the far
   plane entry of its producer.  Returns w. */
static __device__ __noinline__ double sweep_far_row(char* far_adj, int word, int me, double w, double dx) {
    if (word < 0 && (word & kHasFarFlag)) {
        double* fa = reinterpret_cast<double*> (far_adj + (size_t) me * kFarRowBytes);
        w = __dadd_rn(w, *fa);
        *fa = 0.0;
    }
    if (word & kFarFlag) {
        double* fa = reinterpret_cast<double*> (far_adj + (size_t) (word & kIdMask) * kFarRowBytes);
        *fa = __dadd_rn(*fa, __dmul_rn(w, dx));
    }
    return w;
}

/* 16-byte generic load of two headers */
__device__ __forceinline__ int4 ld_hdr2(const void* p) {
    int4 v;
    asm volatile("ld.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

/*
 * Reverse sweep of a work-item recorded by the uniform recorder: all 32 lanes hold the same rows with
 * the same id words, only the partials differ.  Restates the loop of ad4cl.h:786-797 (w = g[id];
 * g[id] = 0; g[operand] += w * dx) on the slot stream.
 *
 * Rows are visited in aligned batches of 4 (row 4b+3 down to 4b), software-pipelined two deep and
 * unrolled twice so that no register is ever copied:
 *     headers of batch b-2 (cell offsets)  ->  cells of batch b-1  ->  arithmetic of batch b
 * Header loads are broadcasts (the same address in all lanes); a batch's headers are read twice (once
 * for the offsets, once for the id words) instead of being held in registers.  All class tests are
 * on warp-uniform values: no divergence, and an independent's contribution is reduced over the warp
 * without any ballot / leader loop.
 */
template <int MODE>
struct UniformSweep {
    const char* hdr;
    const char* cells;
    char* far_adj;
    const SweepTarget& T;
    unsigned wmask8;      /* (W - 1) << 3 */
    double w;
    PendingParam& pend;

    /* batch indices below 0 are clamped to batch 0: what is loaded for them is never used (the loop ends
       before it would be), and the loads need no predicates */
    __device__ __forceinline__ void load_offsets(int b, unsigned (&a)[4]) const {
        const char* p = hdr + (size_t) (b < 0 ? 0 : b) * (4 * kHdrBytes);
        const int4 x = ld_hdr2(p), y = ld_hdr2(p + 16);
        a[0] = (unsigned) x.y; a[1] = (unsigned) x.w; a[2] = (unsigned) y.y; a[3] = (unsigned) y.w;
    }
    __device__ __forceinline__ void load_words(int b, int (&wd)[4]) const {
        const char* p = hdr + (size_t) (b < 0 ? 0 : b) * (4 * kHdrBytes);
        const int4 x = ld_hdr2(p), y = ld_hdr2(p + 16);
        wd[0] = x.x; wd[1] = x.z; wd[2] = y.x; wd[3] = y.z;
    }
    __device__ __forceinline__ void load_cells(const unsigned (&a)[4], double (&d)[4]) const {
#pragma unroll
        for (int j = 0; j < 4; j++) d[j] = *reinterpret_cast<const double*> (cells + a[j]);
    }
    /* One row.  FAR = false is the body of a batch none of whose four rows is FAR or HAS_FAR: no far code
       at all (one test per BATCH decides).  Every test is on a warp-uniform value. */
    template <bool FAR>
    __device__ __forceinline__ void row(int word, double dx, unsigned pop_addr, int me) {
        if (word < 0) { /* END: pop the adjoint of the operator that ends at this slot */
            w = lds_f64(pop_addr);
            sts_f64(pop_addr, 0.0);
        }
        if (FAR && (word & (kFarFlag | kHasFarFlag))) w = sweep_far_row(far_adj, word, me, w, dx);
        /* explicit rounding, the same in every sweep (and every build): a temporary's adjoint takes ONE fused
           multiply-add per slot, an independent's contribution is the rounded product (it is reduced over
           the warp before it is added) */
        if ((word & (kNopFlag | kParamFlag | kFarFlag)) == 0) { /* a recent temporary: its ring entry */
            const unsigned a = T.ring + (((unsigned) word << 3) & wmask8);
            sts_f64(a, __fma_rn(w, dx, lds_f64(a)));
        }
        const double c = __dmul_rn(w, dx);
        if (word & kParamFlag) {
            const int id = word & kIdMask;
            if (MODE == kAccThread) {
                const unsigned a = T.thread_acc + (unsigned) id * T.stride;
                sts_f64(a, __dadd_rn(lds_f64(a), c));
            } else if (MODE == kAccWarp) {
                param_add_paired(T.warp_acc, pend, id, c);
            } else {
                const double v = warp_sum(c);
                if ((threadIdx.x & 31) == 0) atomicAdd(&T.global_acc[id], v);
            }
        }
    }
    __device__ __forceinline__ void batch(int b, unsigned rb, const int (&wd)[4], const double (&d)[4]) {
        if ((wd[0] | wd[1] | wd[2] | wd[3]) & (kFarFlag | kHasFarFlag)) {
#pragma unroll
            for (int j = 3; j >= 0; j--) row<true>(wd[j], d[j], rb + 8u * j, 4 * b + j);
        } else {
#pragma unroll
            for (int j = 3; j >= 0; j--) row<false>(wd[j], d[j], rb + 8u * j, 4 * b + j);
        }
    }
};

template <int MODE, bool PREFETCH>
__device__ __forceinline__ void tape_sweep_uniform(ThreadTape& t, int out_id, double seed, const SweepTarget& T, PendingParam& pend) {
    int rows = sweep_seed<MODE>(t.next_id - t.first_temp, t.first_temp, t.wm1, out_id, seed, T);
    if (rows <= 0) return;
#if AD4CL_DEBUG_BOUNDS
    {   /* self-checking build: every header the sweep is about to follow is validated first, row r by lane
           r mod 32 -- the cell offset names the +1 / -1 cell or the row's own cell, a temporary operand precedes
           its use, an independent exists, far rows have a plane; anything else is reported, not swept */
        const int cap_rows = t.cap_id - t.first_temp - 1 + kGuardRows;
        int bad = rows > cap_rows ? 1 : 0;
        for (int r = threadIdx.x & 31; r < rows && !bad; r += 32) {
            const int2 h = *reinterpret_cast<const int2*> (tape_hdr(t) + (size_t) r * kHdrBytes);
            const int word = h.x, cl = word & (kNopFlag | kParamFlag), pl = word & kIdMask;
            const unsigned off = (unsigned) h.y;
            if (!(off == 0u || off == (unsigned) kCellRowBytes || off == (unsigned) (kUnitCells + r) * kCellRowBytes)) bad = 1;
            if (cl == 0 && pl >= r) bad = 1;
            if (cl == kParamFlag && pl >= t.first_temp) bad = 1;
            if ((word & (kFarFlag | kHasFarFlag)) && t.far_adj == nullptr) bad = 1;
        }
        if (__any_sync(0xffffffffu, bad)) {
            t.status |= kErrBounds;
            return;
        }
    }
#endif
    /* pad to whole batches: the rows above `rows` are dead or were never written */
    const int nb = (rows + 3) >> 2;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (rows + j < 4 * nb) *reinterpret_cast<int2*> (tape_hdr(t) + (size_t) (rows + j) * kHdrBytes) = make_int2(kNopFlag, 0);
    }
    /* (W - 1) << 3, pinned in a register (the compiler otherwise re-derives it from the launch
       parameters in the constant bank at every use) */
    unsigned wmask8 = (unsigned) t.wm1 << 3;
    asm volatile("" : "+r"(wmask8));
    /* likewise the two plane pointers (otherwise re-derived from the launch parameters every batch) */
    const char* hdr = tape_hdr(t);
    const char* cells = tape_cells(t);
    asm volatile("" : "+l"(hdr), "+l"(cells));
    UniformSweep<MODE> S{hdr, cells, t.far_adj, T, wmask8, 0.0, pend};

    int b = nb - 1;
    /* shared-memory offset of the ring entry of row 4b (W >= 4 and batches are aligned: the four pops of
       a batch are at +0, +8, +16, +24 from it); stepped down with the batch index */
    unsigned rb = ((unsigned) (4 * b) << 3) & wmask8;
    unsigned offA[4], offB[4];
    int wordA[4], wordB[4];
    double dxA[4], dxB[4];
    S.load_offsets(b, offA);
    S.load_words(b, wordA);
    S.load_cells(offA, dxA);
    S.load_offsets(b - 1, offB);
    /* `prefetch`: a cell's address does not depend on its header (cell 2 + s belongs to row s), so lanes 0..7 can
       pull the 1 KB of cells of the batch THREE below the one in use into L2 long before its headers are read --
       for tapes longer than L2, where the one batch of register look-ahead is shorter than an HBM round trip.
       Rows whose partial is +-1 have no cell: their lines are fetched for nothing, which a bandwidth-bound sweep
       would pay for and a latency-bound one does not (the tuner measures). */
    const char* pf_base = cells - (size_t) (threadIdx.x & 31) * 8 + (size_t) kUnitCells * kCellRowBytes + (size_t) (threadIdx.x & 31) * 128;
    const bool pf_lane = PREFETCH && (threadIdx.x & 31) < 8;   /* a template parameter: the sweep without prefetch carries none of this */
    for (;;) {
        if (PREFETCH && pf_lane && b >= 3) asm volatile("prefetch.global.L2 [%0];" : : "l"(pf_base + (size_t) (b - 3) * (4 * kCellRowBytes)));
        /* A holds batch b; B becomes batch b-1 */
        S.load_cells(offB, dxB);
        S.load_words(b - 1, wordB);
        S.load_offsets(b - 2, offA);
        S.batch(b, T.ring + rb, wordA, dxA);
        rb = (rb - 32u) & wmask8;
        if (--b < 0) break;
        if (PREFETCH && pf_lane && b >= 3) asm volatile("prefetch.global.L2 [%0];" : : "l"(pf_base + (size_t) (b - 3) * (4 * kCellRowBytes)));
        /* B holds batch b; A becomes batch b-1 */
        S.load_cells(offA, dxA);
        S.load_words(b - 1, wordA);
        S.load_offsets(b - 2, offB);
        S.batch(b, T.ring + rb, wordB, dxB);
        rb = (rb - 32u) & wmask8;
        if (--b < 0) break;
    }
}

/* The far cases of one row of the general sweep (per lane), out of line. */
static __device__ __noinline__ double sweep_far_lane(char* far_adj, int idw, int me, double w, double dx) {
    if ((idw & kHasFarFlag) && idw < 0) { /* contributions that came through the far plane */
        double* fa = reinterpret_cast<double*> (far_adj + (size_t) me * kFarRowBytes);
        w = __dadd_rn(w, *fa);
        *fa = 0.0;
    }
    if ((idw & kFarFlag) && !(idw & kNopFlag) && far_adj != nullptr) {
        double* fa = reinterpret_cast<double*> (far_adj + (size_t) (idw & kIdMask) * kFarRowBytes);
        *fa = __dadd_rn(*fa, __dmul_rn(w, dx));
    }
    return w;
}

#ifndef AD4CL_PREFETCH_BATCHES
#define AD4CL_PREFETCH_BATCHES 2
#endif
constexpr int kPrefetchBatches = AD4CL_PREFETCH_BATCHES; /* the L2 prefetch runs this many batches below the one being fetched */

/*
 * General sweep (rows of the general recorder): every lane walks its OWN rows -- lanes of a diverged warp
 * hold tapes of different lengths, id words may differ between lanes.  Restates ad4cl.h:786-797 on the slot
 * stream.  Four rows per batch, fetched one batch ahead of their use; `prefetch` additionally pulls the batch
 * after the next into L2 (lanes 0-11 touch its 12 cache lines).  All lanes of the warp iterate together.
 *
 * Per batch ONE pair of 64-bit votes asks whether the four id words are the same in all 32 lanes (they are,
 * whenever neighbouring work-items record the same operator sequence).  If so the rows are classified on
 * warp-uniform branches and an independent's contribution is reduced over the warp without the ballot /
 * leader loop, two contributions per butterfly (param_add_paired); if not, every lane classifies its own
 * word.  Per lane both bodies perform the same fp64 operations in the same order.
 * Out of line and by value: the hot entry stays small.  Returns Status bits (AD4CL_DEBUG_BOUNDS builds).
 */
template <int MODE>
static __device__ __noinline__ int tape_sweep_general(const char* rows_base, int id_off, char* far_adj, int rows,
        int first_temp, int wmask, int out_id, double seed, SweepTarget T, int cap_rows, int prefetch) {
    constexpr int U = kSweepUnroll;
    rows = sweep_seed<MODE>(rows, first_temp, wmask, out_id, seed, T);
    int status = kOk;
#if AD4CL_DEBUG_BOUNDS
    if (rows < 0 || rows > cap_rows) {
        status |= kErrBounds;
        rows = 0;
    }
#endif
    PendingParam pend = {0.0, -1};   /* flushed before returning: the caller's pending pair is its own */
    const unsigned pad_off = (unsigned) (wmask + 1) << 3;
    const char* cur = rows_base + (size_t) rows * kRowBytes;
    double w = 0.0;
    double dx_now[U], dx_next[U];
    int id_now[U], id_next[U];
    /* the first batch takes the rows that do not fill a whole batch */
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
#pragma unroll
        for (int j = 0; j < U; j++) {   /* operands precede their use; independents are below first_temp; far rows need the plane */
            const int idw = id_now[j], me = idx - 1 - j;
            const int cl = idw & (kNopFlag | kParamFlag), pl = idw & kIdMask;
            bool bad = (cl == 0 && (pl < 0 || pl >= me)) || (cl == kParamFlag && pl >= first_temp);
            if ((idw & (kFarFlag | kHasFarFlag)) && (far_adj == nullptr || me < 0 || me >= cap_rows)) bad = true;
            if (bad) {
                status |= kErrBounds;
                id_now[j] = kNopFlag;
            }
        }
#endif
        int same01 = 0, same23 = 0;
#if AD4CL_LANE_UNIFORM_BATCHES
        __match_all_sync(0xffffffffu, ((unsigned long long) (unsigned) id_now[1] << 32) | (unsigned) id_now[0], &same01);
        __match_all_sync(0xffffffffu, ((unsigned long long) (unsigned) id_now[3] << 32) | (unsigned) id_now[2], &same23);
#endif
        if (same01 && same23) {
            /* ---- warp-uniform batch: every test below is on a value all 32 lanes share ---- */
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int idw = id_now[j];
                const int me = idx - 1 - j;
                if (idw < 0) { /* END: pop the adjoint of the operator that ends at this slot */
                    const unsigned r = T.ring + ((unsigned) (me & wmask) << 3);
                    w = lds_f64(r);
                    sts_f64(r, 0.0);
                }
                if (idw & (kFarFlag | kHasFarFlag)) w = sweep_far_lane(far_adj, idw, me, w, dx_now[j]);
                if ((idw & (kNopFlag | kParamFlag | kFarFlag)) == 0) {
                    const unsigned a = T.ring + ((unsigned) (idw & wmask) << 3);
                    sts_f64(a, __fma_rn(w, dx_now[j], lds_f64(a)));
                }
                const double c = __dmul_rn(w, dx_now[j]);
                if ((idw & (kNopFlag | kParamFlag)) == kParamFlag) {
                    const int id = idw & kIdMask;
                    if (MODE == kAccThread) {
                        const unsigned a = T.thread_acc + (unsigned) id * T.stride;
                        sts_f64(a, __dadd_rn(lds_f64(a), c));
                    } else if (MODE == kAccWarp) {
                        param_add_paired(T.warp_acc, pend, id, c);
                    } else {
                        const double v = warp_sum(c);
                        if (id_off == kRowIdOffset) atomicAdd(&T.global_acc[id], v);   /* lane 0 */
                    }
                }
            }
        } else {
            /* ---- every lane classifies its own word ---- */
            if (MODE == kAccWarp) param_flush(T.warp_acc, pend);   /* keep the order of the additions fixed */
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int idw = id_now[j];
                const int me = idx - 1 - j;
                if (idw < 0) {
                    const unsigned r = T.ring + ((unsigned) (me & wmask) << 3);
                    w = lds_f64(r);
                    sts_f64(r, 0.0);
                }
                if (idw & (kFarFlag | kHasFarFlag)) w = sweep_far_lane(far_adj, idw, me, w, dx_now[j]);
                const int payload = idw & kIdMask;
                const double c = __dmul_rn(w, dx_now[j]);
                const int cls = idw & (kNopFlag | kParamFlag | kFarFlag);
                {   /* Branch-free scatter: a recent temporary adds to its ring entry, with per-thread
                       accumulators (kAccThread) an independent adds to its accumulator, and a slot of any
                       other class adds to the pad entry of this thread's ring (index W), which nothing reads. */
                    unsigned a = T.ring + (cls == 0 ? ((unsigned) (payload & wmask) << 3) : pad_off);
                    if (MODE == kAccThread) a = cls == kParamFlag ? T.thread_acc + (unsigned) payload * T.stride : a;
                    sts_f64(a, (MODE == kAccThread && cls == kParamFlag) ? __dadd_rn(lds_f64(a), c)
                                                                          : __fma_rn(w, dx_now[j], lds_f64(a)));
                }
                if (MODE != kAccThread) sweep_param<MODE>(T, cls == kParamFlag, payload, c);
            }
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
    if (MODE == kAccWarp) param_flush(T.warp_acc, pend);
    (void) cap_rows;
    return status;
}

#endif /* AD4CL_TAPE_STATIC */

} // namespace ad4cl
