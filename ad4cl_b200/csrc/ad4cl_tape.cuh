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
 *   [26] CHAIN    operand is the result of the immediately preceding operator
 *   [25:0] payload
 */
constexpr int kEndFlag = (int) 0x80000000u;
constexpr int kNopFlag = 0x40000000;
constexpr int kParamFlag = 0x20000000;
constexpr int kFarFlag = 0x10000000;
constexpr int kHasFarFlag = 0x08000000;
constexpr int kChainFlag = 0x04000000;
constexpr int kIdMask = 0x03FFFFFF;
/* "no output written" / not a variable.  INT_MIN, not -1: as an operand it fails the push's one unsigned
   reach test for every row index (ad_plus_eq_g on the virtual out before it was assigned) and is then
   recorded as a constant by tape_classify_rare. */
constexpr int kNoId = (int) 0x80000000u;

/* second header word: byte offset of the slot's cell in the lane's cell column (a multiple of 256) | flags */
constexpr unsigned kAuxNonUniform = 1u;     /* the row is not warp-uniform: id word in the lane-word plane, cell always own */
constexpr unsigned kAuxOffsetMask = ~255u;

constexpr int kHdrBytes = 8;
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
/* A/B switch (make general -> libad4cl_cuda_general.so, tests/test_gpu_format_ab.py): 1 records every
   row as non-uniform, i.e. the per-lane format and the general sweep everywhere.  Every lane then
   performs the same floating-point operations in the same order as in the uniform format, so the
   results are bit-identical by construction. */
#ifndef AD4CL_FORCE_GENERAL
#define AD4CL_FORCE_GENERAL 0
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
__device__ __forceinline__ void tape_push(ThreadTape* t, double dx, int id, int flags, int prev_id) {
    (void) prev_id;
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
__device__ __forceinline__ void tape_sweep_static(const ThreadTape& t, int out_id, double (&g)[AD4CL_P > 0 ? AD4CL_P : 1]) {
    const int n = t.next_id - AD4CL_P;
    double adj[AD4CL_STATIC_SLOTS];
    StaticIndex<AD4CL_STATIC_SLOTS>::zero(adj);
    if ((unsigned) out_id < (unsigned) AD4CL_P) StaticIndex<(AD4CL_P > 0 ? AD4CL_P : 1)>::add(g, out_id, 1.0);
    else if (out_id - AD4CL_P < n) StaticIndex<AD4CL_STATIC_SLOTS>::set(adj, out_id - AD4CL_P, 1.0);
    double w = 0.0;
    StaticSweep<AD4CL_STATIC_SLOTS>::run(t, n, w, adj, g);
}

#else /* !AD4CL_TAPE_STATIC: the arena / shared-memory tape */

struct ThreadTape {
    char* hdr;            /* header plane of this warp, row 0 (the same address in all lanes) */
    char* hdr_cur;        /* header of the next free row */
    char* cells;          /* this lane's cell column: cell k at cells + 256 k */
    char* lane_words;     /* this lane's id-word column for non-uniform rows: row r at lane_words + 128 r */
    char* far_adj;        /* this lane's far-adjoint column (row stride kFarRowBytes), or null */
    unsigned cell_off;    /* byte offset of this lane's next free cell */
    int next_id;          /* first_temp + index of the next free row */
    int cap_id;           /* first_temp + capacity in slots; kGuardRows guard rows follow */
    int first_temp;       /* number of independents: ids >= first_temp are tape positions */
    int wm1;              /* W - 1, W a power of two >= 4 */
    int nops;             /* operators recorded by the current work-item (statistics) */
    /* Uniformity vote key of a push = (id word & key_and) | key_or.  While every push of the current
       work-item was warp-uniform: key_and = ~0, key_or = 0 (the key is the id word; lanes that took
       part in the same full-mask uniform pushes are at the same row, so the row is not part of the
       key).  After the first push that was not: key_and = 0, key_or = HAS_FAR | lane -- a value no
       clean key can have (HAS_FAR is only ever set later, by tape_note_far) and no other lane has, so
       a lane that diverged never again looks uniform to anybody until the next work-item. */
    unsigned key_and, key_or;
    int status;           /* sticky Status bits */
    long long total_ops;  /* statistics over all work-items of this thread */
    long long total_slots;
};

__device__ __forceinline__ bool tape_is_uniform(const ThreadTape& t) { return t.key_or == 0u; }

__device__ __forceinline__ void tape_bind(ThreadTape& t, char* warp_region, unsigned cap_slots,
        int first_temp, int window, bool far) {
    const int lane = threadIdx.x & 31;
    const RegionLayout lay = region_layout(cap_slots, far);
    t.hdr = warp_region;
    t.hdr_cur = warp_region;
    t.cells = warp_region + lay.off_cells + lane * 8;
    t.lane_words = warp_region + lay.off_lane_words + lane * 4;
    t.far_adj = far ? warp_region + lay.off_far + lane * 8 : nullptr;
    t.cell_off = kUnitCells * kCellRowBytes;
    t.next_id = first_temp;
    t.cap_id = first_temp + (int) cap_slots;
    t.first_temp = first_temp;
    t.wm1 = window - 1;
    t.nops = 0;
    t.key_and = ~0u;
    t.key_or = 0u;
    t.status = kOk;
    t.total_ops = 0;
    t.total_slots = 0;
    /* the two constant cells every unit slot points at */
    *reinterpret_cast<double*> (t.cells) = 1.0;
    *reinterpret_cast<double*> (t.cells + kCellRowBytes) = -1.0;
}

/*
 * Rare operand classes, out of line and by value (the tape cursor stays in registers; ~50 push sites
 * per objective would otherwise each carry a copy):
 *   - a temporary W or more slots old cannot use the ring: FAR, and HAS_FAR on the END slot of the
 *     operator that produced it -- in the warp's header (read when that row is uniform) and in this
 *     lane's own word (read when it is not).  The far plane is all zeros between work-items (the
 *     sweep restores it), so lanes that see HAS_FAR without having noted anything read a zero.
 *   - an id that is neither an independent nor a position below the current row is not a variable of
 *     this tape (the virtual out before it is written, an uninitialised ad_variable): recorded as
 *     a constant operand instead of an id the sweep would follow out of the arena.
 * Returns the id word in the low half and Status bits in the high half.
 */
static __device__ __noinline__ unsigned long long tape_classify_rare(char* hdr, char* lane_words, const char* far_adj,
        int id, int first_temp, int row, int flags) {
    const int pos = id - first_temp;
    if (pos < 0 || pos >= row) return (unsigned) (flags | kNopFlag);
    if (far_adj == nullptr) return ((unsigned long long) kErrWindow << 32) | (unsigned) (flags | kNopFlag);
    atomicOr(reinterpret_cast<int*> (hdr + (size_t) pos * kHdrBytes), kHasFarFlag);
    atomicOr(reinterpret_cast<int*> (lane_words + (size_t) pos * kLaneWordRowBytes), kHasFarFlag);
    return (unsigned) (pos | flags | kFarFlag);
}

/* The push is not warp-uniform (partial mask, lane-dependent id word, lanes at different rows) or the
   column is full.  Not uniform: the header only says so, the lane keeps its own id word and always
   its own cell.  Every lane that passes row r stores the same header (the cell offset is a function
   of the row index: all rows below r were either uniform for the whole warp or non-uniform for
   everybody who passed them).  Full: nothing is stored and the sticky overflow bit is returned (the
   sweep of the work-item is skipped, the host reports AD4CL_ERR_TAPE_OVERFLOW).
   Returns the new cell offset in the low half and Status bits in the high half. */
static __device__ __noinline__ unsigned long long tape_push_diverged(char* hdr_cur, char* cells, char* lane_words,
        int row, int cap_rows, int word, double dx, unsigned cell_off) {
    if (row >= cap_rows) return ((unsigned long long) kErrTapeOverflow << 32) | cell_off;
    *reinterpret_cast<int2*> (hdr_cur) = make_int2(0, (int) (cell_off | kAuxNonUniform));
    *reinterpret_cast<int*> (lane_words + (size_t) row * kLaneWordRowBytes) = word;
    *reinterpret_cast<double*> (cells + cell_off) = dx;
    return cell_off + kCellRowBytes;
}

/* Push one (partial, operand) slot.  `flags` is 0, kEndFlag or kEndFlag|kNopFlag; UNIT = +1 / -1
   says the partial is the compile-time constant +1.0 / -1.0 (dx is then ignored), 0 a general
   partial; prev_id is the id of the previous operator's result (CHAIN detection).  The operand is
   classified here (independent / chained / recent / far temporary) so the sweep only tests flag
   bits.  The common path -- all 32 lanes push the same id word and the column has room -- is two
   stores; everything else is one out-of-line call. */
template <int UNIT>
__device__ __forceinline__ void tape_push(ThreadTape* t, double dx, int id, int flags, int prev_id) {
    int word;
    if (flags & kNopFlag) {
        word = flags;  /* a leaf: no operand */
    } else {
        const bool is_param = (unsigned) id < (unsigned) t->first_temp;
        word = (is_param ? (id | kParamFlag) : (id - t->first_temp)) | flags;
        if (id == prev_id && !is_param) word |= kChainFlag;
        /* A temporary is within the ring's reach when it is fewer than W - 1 rows below this one (the
           operator being recorded ends at this row or the next).  One unsigned compare also rejects
           ids at or above the current row. */
        const bool near = (unsigned) (t->next_id - 1 - id) < (unsigned) (t->wm1 - 1);
        if (!is_param && !near) {
            const unsigned long long r = tape_classify_rare(t->hdr, t->lane_words, t->far_adj, id, t->first_temp,
                    t->next_id - t->first_temp, flags);
            word = (int) (unsigned) r;
            t->status |= (int) (r >> 32);
        }
    }
    const unsigned mask = __activemask();
    int same = 0;
    __match_all_sync(mask, ((unsigned) word & t->key_and) | t->key_or, &same);
    if (!AD4CL_FORCE_GENERAL && same && mask == 0xffffffffu && t->next_id < t->cap_id) {
        const unsigned aux = UNIT == 0 ? t->cell_off : (UNIT > 0 ? 0u : (unsigned) kCellRowBytes);
        *reinterpret_cast<int2*> (t->hdr_cur) = make_int2(word, (int) aux);
        if (UNIT == 0) {
            *reinterpret_cast<double*> (t->cells + t->cell_off) = dx;
            t->cell_off += kCellRowBytes;
        }
    } else {
        const unsigned long long r = tape_push_diverged(t->hdr_cur, t->cells, t->lane_words, t->next_id - t->first_temp,
                t->cap_id - t->first_temp, word, UNIT == 0 ? dx : (UNIT > 0 ? 1.0 : -1.0), t->cell_off);
        t->cell_off = (unsigned) r;
        t->status |= (int) (r >> 32);
        t->key_and = 0u;
        t->key_or = (unsigned) kHasFarFlag | (threadIdx.x & 31u);
    }
    t->hdr_cur += kHdrBytes;
    t->next_id++;
}

/* Close the operator whose slots were just pushed; returns its result id (the position of its END
   slot).  Capacity is checked by the pushes themselves. */
__device__ __forceinline__ int tape_close_op(ThreadTape* t) {
    t->nops++;
    return t->next_id - 1;
}

/* Forget the current work-item's tape (the arena column is reused). */
__device__ __forceinline__ void tape_rewind(ThreadTape& t) {
    t.total_ops += t.nops;
    t.total_slots += t.next_id - t.first_temp;
    t.hdr_cur = t.hdr;
    t.cell_off = kUnitCells * kCellRowBytes;
    t.next_id = t.first_temp;
    t.nops = 0;
    t.key_and = ~0u;
    t.key_or = 0u;
}

/* Debug / export accessor: id word (CHAIN stripped) and partial of row `row` of this lane's tape. */
__device__ __forceinline__ void tape_read_row(const ThreadTape* t, int row, double* dx, int* word) {
    const int2 h = *reinterpret_cast<const int2*> (t->hdr + (size_t) row * kHdrBytes);
    const unsigned aux = (unsigned) h.y;
    int w = h.x;
    if (aux & kAuxNonUniform) w = *reinterpret_cast<const int*> (t->lane_words + (size_t) row * kLaneWordRowBytes);
    *word = w & ~kChainFlag;
    *dx = *reinterpret_cast<const double*> (t->cells + (aux & kAuxOffsetMask));
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

/* warp-uniform form (uniform sweep): the same independent in all 32 lanes */
template <int MODE>
__device__ __forceinline__ void sweep_param_uniform(const SweepTarget& T, int id, double c) {
    if (MODE == kAccThread) {
        const unsigned a = T.thread_acc + (unsigned) id * T.stride;
        sts_f64(a, lds_f64(a) + c);
    } else if (MODE == kAccWarp) {
        /* the butterfly of warp_param_reduce with all 32 lanes in one group: same bits */
        const double v = warp_sum(c);
        const unsigned a = T.warp_acc + (unsigned) id * 8u;
        sts_f64(a, lds_f64(a) + v);
    } else {
        const double v = warp_sum(c);
        if ((threadIdx.x & 31) == 0) atomicAdd(&T.global_acc[id], v);
    }
}

/* Seeds d(out) = 1 and returns the number of rows the sweep has to visit (rows recorded after `out`
   are dead; those that could alias the seed's ring entry are dropped). */
template <int MODE>
__device__ __forceinline__ int sweep_seed(int rows, int first_temp, int wmask, int out_id, const SweepTarget& T) {
    const int seed_pos = out_id - first_temp;
    sweep_param<MODE>(T, out_id >= 0 && out_id < first_temp, out_id, 1.0);
    if (seed_pos >= 0 && seed_pos < rows) {
        if (rows - seed_pos > wmask) rows = seed_pos + wmask;
        sts_f64(T.ring + ((unsigned) (seed_pos & wmask) << 3), 1.0);
    }
    return rows;
}

/*
 * Reverse sweep of a work-item whose whole tape was recorded warp-uniformly (the common case): all
 * 32 lanes hold the same rows with the same id words, only the partials differ.  Restates the loop
 * of ad4cl.h:786-797 (w = g[id]; g[id] = 0; g[operand] += w * dx) on the slot stream.
 *
 * Rows are visited in aligned batches of 4 (row 4b+3 down to 4b), software-pipelined two deep and
 * unrolled twice so that no register is ever copied:
 *     headers of batch b-2 (cell offsets)  ->  cells of batch b-1  ->  arithmetic of batch b
 * Header loads are 16-byte broadcasts (the same address in all lanes); a batch's headers are read
 * twice (once for the offsets, once for the id words) instead of being held in registers.
 * All class tests are on warp-uniform values: real branches, no divergence.
 */
/* 16-byte generic load of two headers; `volatile` keeps it ONE instruction (the compiler otherwise splits
   it into the 32-bit loads of the components it needs) and in program order with the other loads */
__device__ __forceinline__ int4 ld_hdr2(const void* p) {
    int4 v;
    asm volatile("ld.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int MODE>
struct UniformSweep {
    const char* hdr;
    const char* cells;
    char* far_adj;
    const SweepTarget& T;
    int wmask;
    double w, carry;

    __device__ __forceinline__ void load_offsets(int b, unsigned (&a)[4]) const {
        if (b >= 0) {
            const char* p = hdr + (size_t) b * (4 * kHdrBytes);
            const int4 x = ld_hdr2(p), y = ld_hdr2(p + 16);
            a[0] = (unsigned) x.y; a[1] = (unsigned) x.w; a[2] = (unsigned) y.y; a[3] = (unsigned) y.w;
        } else {
            a[0] = a[1] = a[2] = a[3] = 0u;
        }
    }
    __device__ __forceinline__ void load_words(int b, int (&wd)[4]) const {
        if (b >= 0) {
            const char* p = hdr + (size_t) b * (4 * kHdrBytes);
            const int4 x = ld_hdr2(p), y = ld_hdr2(p + 16);
            wd[0] = x.x; wd[1] = x.z; wd[2] = y.x; wd[3] = y.z;
        } else {
            wd[0] = wd[1] = wd[2] = wd[3] = kNopFlag;
        }
    }
    __device__ __forceinline__ void load_cells(const unsigned (&a)[4], double (&d)[4]) const {
#pragma unroll
        for (int j = 0; j < 4; j++) d[j] = *reinterpret_cast<const double*> (cells + a[j]);
    }
    /* One row.  Every test is on a warp-uniform value.  The common classes (pop, chained operand,
       recent temporary) are short predicated sequences; an independent is one real branch around the
       warp reduction; the two far cases share one rare test. */
    __device__ __forceinline__ void row(int word, double dx, unsigned pop_addr, int me) {
        if (word < 0) { /* END: pop the adjoint of the operator that ends at this slot */
            w = lds_f64(pop_addr) + carry;
            sts_f64(pop_addr, 0.0);
            carry = 0.0;
        }
        if (word & (kFarFlag | kHasFarFlag)) { /* rare */
            if (word < 0 && (word & kHasFarFlag)) { /* contributions that came through the far plane */
                double* fa = reinterpret_cast<double*> (far_adj + (size_t) me * kFarRowBytes);
                w += *fa;
                *fa = 0.0;
            }
            if (word & kFarFlag) *reinterpret_cast<double*> (far_adj + (size_t) (word & kIdMask) * kFarRowBytes) += w * dx;
        }
        const double c = w * dx;
        if (word & kChainFlag) carry += c;
        if ((word & (kNopFlag | kParamFlag | kFarFlag | kChainFlag)) == 0) {
            const unsigned a = T.ring + ((unsigned) (word & wmask) << 3);
            sts_f64(a, lds_f64(a) + c);
        }
        if (word & kParamFlag) sweep_param_uniform<MODE>(T, word & kIdMask, c);
    }
    __device__ __forceinline__ void batch(int b, unsigned rb, const int (&wd)[4], const double (&d)[4]) {
#pragma unroll
        for (int j = 3; j >= 0; j--) row(wd[j], d[j], rb + 8u * j, 4 * b + j);
    }
};

template <int MODE>
__device__ __forceinline__ void tape_sweep_uniform(ThreadTape& t, int out_id, const SweepTarget& T) {
    int rows = sweep_seed<MODE>(t.next_id - t.first_temp, t.first_temp, t.wm1, out_id, T);
    if (rows <= 0) return;
    /* pad to whole batches: the rows above `rows` are dead or were never written */
    const int nb = (rows + 3) >> 2;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (rows + j < 4 * nb) *reinterpret_cast<int2*> (t.hdr + (size_t) (rows + j) * kHdrBytes) = make_int2(kNopFlag, 0);
    }
    UniformSweep<MODE> S{t.hdr, t.cells, t.far_adj, T, t.wm1, 0.0, 0.0};

    int b = nb - 1;
    /* shared address of the ring entry of row 4b (W >= 4 and batches are aligned: the four pops of a
       batch are at +0, +8, +16, +24 from it); stepped down with the batch index, kept in a register */
    const unsigned ring_span = (unsigned) (t.wm1 + 1) << 3;
    unsigned rb = (unsigned) ((4 * b) & t.wm1) << 3;
    unsigned offA[4], offB[4];
    int wordA[4], wordB[4];
    double dxA[4], dxB[4];
    S.load_offsets(b, offA);
    S.load_words(b, wordA);
    S.load_cells(offA, dxA);
    S.load_offsets(b - 1, offB);
    for (;;) {
        /* A holds batch b; B becomes batch b-1 */
        S.load_cells(offB, dxB);
        S.load_words(b - 1, wordB);
        S.load_offsets(b - 2, offA);
        S.batch(b, T.ring + rb, wordA, dxA);
        rb = (rb - 32u) & (ring_span - 1u);
        if (--b < 0) break;
        /* B holds batch b; A becomes batch b-1 */
        S.load_cells(offA, dxA);
        S.load_words(b - 1, wordA);
        S.load_offsets(b - 2, offB);
        S.batch(b, T.ring + rb, wordB, dxB);
        rb = (rb - 32u) & (ring_span - 1u);
        if (--b < 0) break;
    }
}

/*
 * General sweep: every lane walks its OWN rows (lanes of a diverged warp hold tapes of different
 * lengths; rows whose id words differ between lanes keep them in the lane-word plane).  Out of
 * line and by value: it runs once per work-item of the (rare) warps that need it, and the hot
 * entry stays small.  Four rows per batch; the headers of the next batch are requested before the
 * cells of the current one are used.  Returns Status bits (AD4CL_DEBUG_BOUNDS builds).
 */
template <int MODE>
static __device__ __noinline__ int tape_sweep_general(const char* hdr, const char* cells, const char* lane_words,
        char* far_adj, int rows, int first_temp, int wmask, int out_id, SweepTarget T, int cap_rows) {
    constexpr int U = kSweepUnroll;
    rows = sweep_seed<MODE>(rows, first_temp, wmask, out_id, T);
    int status = kOk;
    double w = 0.0, carry = 0.0;
    int2 h_now[U], h_next[U];
    /* the first batch takes the rows that do not fill a whole batch */
    const int head = rows & (U - 1);
    int top = rows;                 /* slot j of the current batch has index top - 1 - j */
#pragma unroll
    for (int j = 0; j < U; j++) {
        h_now[j] = j < head ? *reinterpret_cast<const int2*> (hdr + (size_t) (top - 1 - j) * kHdrBytes)
                            : make_int2(kNopFlag, 0);
    }
    int next_top = rows - head;
    bool more = true;
    while (__any_sync(0xffffffffu, more)) {
        const bool fetch = next_top > 0;
#pragma unroll
        for (int j = 0; j < U; j++) {
            h_next[j] = fetch ? *reinterpret_cast<const int2*> (hdr + (size_t) (next_top - 1 - j) * kHdrBytes)
                              : make_int2(kNopFlag, 0);
        }
        double dx[U];
        int word[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const unsigned aux = (unsigned) h_now[j].y;
            const int me = top - 1 - j;
#if AD4CL_DEBUG_BOUNDS
            if ((aux & kAuxOffsetMask) > (unsigned) (cap_rows + kUnitCells - 1) * kCellRowBytes || me >= cap_rows) {
                status |= kErrBounds; /* report and neutralise the row; every lane keeps reaching the votes below */
                dx[j] = 0.0;
                word[j] = kNopFlag;
                continue;
            }
#endif
            dx[j] = *reinterpret_cast<const double*> (cells + (aux & kAuxOffsetMask));
            word[j] = (aux & kAuxNonUniform) ? *reinterpret_cast<const int*> (lane_words + (size_t) me * kLaneWordRowBytes)
                                             : h_now[j].x;
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            int idw = word[j];
            const int me = top - 1 - j;
#if AD4CL_DEBUG_BOUNDS
            {   /* operands precede their use; independents are below first_temp; far rows need the plane */
                const int cl = idw & (kNopFlag | kParamFlag), pl = idw & kIdMask;
                bool bad = (cl == 0 && (pl < 0 || pl >= me)) || (cl == kParamFlag && pl >= first_temp);
                if ((idw & (kFarFlag | kHasFarFlag)) && (far_adj == nullptr || me < 0 || me >= cap_rows)) bad = true;
                if (bad) {
                    status |= kErrBounds;
                    idw = kNopFlag;
                }
            }
#endif
            if (idw < 0) { /* END */
                const unsigned r = T.ring + ((unsigned) (me & wmask) << 3);
                w = lds_f64(r) + carry;
                sts_f64(r, 0.0);
                carry = 0.0;
            }
            if (idw & (kFarFlag | kHasFarFlag)) { /* rare: ONE test for both far cases */
                if ((idw & kHasFarFlag) && idw < 0) {
                    double* fa = reinterpret_cast<double*> (far_adj + (size_t) me * kFarRowBytes);
                    w += *fa;
                    *fa = 0.0;
                }
                if ((idw & kFarFlag) && !(idw & kNopFlag) && far_adj != nullptr) {
                    *reinterpret_cast<double*> (far_adj + (size_t) (idw & kIdMask) * kFarRowBytes) += w * dx[j];
                }
            }
            const int payload = idw & kIdMask;
            const double c = w * dx[j];
            const int cls = idw & (kNopFlag | kParamFlag | kFarFlag | kChainFlag);
            if (cls == kChainFlag) carry += c;
            if (cls == 0) {
                const unsigned a = T.ring + ((unsigned) (payload & wmask) << 3);
                sts_f64(a, lds_f64(a) + c);
            }
            if (MODE == kAccThread) {
                if (cls == kParamFlag) {
                    const unsigned a = T.thread_acc + (unsigned) payload * T.stride;
                    sts_f64(a, lds_f64(a) + c);
                }
            } else {
                sweep_param<MODE>(T, cls == kParamFlag, payload, c);
            }
        }
#pragma unroll
        for (int j = 0; j < U; j++) h_now[j] = h_next[j];
        top = next_top;
        next_top -= U;
        more = fetch;
    }
    (void) cap_rows;
    return status;
}

#endif /* AD4CL_TAPE_STATIC */

} // namespace ad4cl
