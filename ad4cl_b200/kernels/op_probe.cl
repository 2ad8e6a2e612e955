/*
 * op_probe.cl -- parity-test kernel (CUDA side only): work-item i executes ONE operator
 * of ONE family on data-supplied operands and writes what it recorded, so that every
 * operator of ad.cuh can be compared with what the reference's ad.cl records for the same
 * operands (tests/golden/op_probes.json).
 *   d0  : operand values, (a, b) per point
 *   d1  : output, 8 doubles per work-item:
 *         value, k (slots), dx0, operand0, dx1, operand1, result-is-temporary, lhs value
 *         operand = independent id (0 = a's, 1 = b's) or -1 for a temporary
 *   i0  : family  0 global ad_*   1 pad_*   2 private *_p
 *   work-item id = point * AD4CL_PROBE_OPS + op
 */
#define AD4CL_PROBE_OPS 30

#define PROBE_BIN(BASE, NM)                                         \
    case BASE:     r = F(NM)(G, va, vb); break;                     \
    case BASE + 1: r = F(NM##_vd)(G, va, b); break;                 \
    case BASE + 2: r = F(NM##_dv)(G, a, vb); break;
#define PROBE_ALL                                                   \
    PROBE_BIN(0, plus) PROBE_BIN(3, minus) PROBE_BIN(6, times) PROBE_BIN(9, divide) \
    case 12: lhs = va; F(plus_eq)(G, &lhs, vb); r = lhs; break;     \
    case 13: lhs = va; F(plus_eq_d)(G, &lhs, b); r = lhs; break;    \
    case 14: r = F(cos)(G, va); break;                              \
    case 15: r = F(sin)(G, va); break;                              \
    case 16: r = F(tan)(G, va); break;                              \
    case 17: r = F(acos)(G, va); break;                             \
    case 18: r = F(asin)(G, va); break;                             \
    case 19: r = F(atan)(G, va); break;                             \
    case 20: r = F(cosh)(G, va); break;                             \
    case 21: r = F(sinh)(G, va); break;                             \
    case 22: r = F(tanh)(G, va); break;                             \
    case 23: r = F(exp)(G, va); break;                              \
    case 24: r = F(log)(G, va); break;                              \
    case 25: r = F(log10)(G, va); break;                            \
    case 26: r = F(sqrt)(G, va); break;                             \
    case 27: r = F(pow)(G, va, vb); break;                          \
    case 28: r = F(pow_vd)(G, va, b); break;                        \
    case 29: r = F(pow_dv)(G, a, vb); break;

__kernel void op_probe(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        const int op = id % AD4CL_PROBE_OPS, point = id / AD4CL_PROBE_OPS;
        const double a = d0[2 * point], b = d0[2 * point + 1];
        struct ad_variable va = theta[0], vb = theta[1], r, lhs;
        va.value = a;
        vb.value = b;
        lhs = va;
        r = va;
        const int nparams = gs->current_ad_variable_id;
        double dx[2] = {0.0, 0.0};
        int word[2] = {0, 0};
        int k = 0;
        if (i0 == 2) {
            struct ad_private_gradient_structure p;
            ad_init_p(&p);
            p.current_ad_variable_id = nparams;
#define F(nm) ad_##nm##_p
#define G (&p)
            switch (op) { PROBE_ALL }
#undef F
#undef G
            k = p.nslots;
            for (int s = 0; s < k; s++) {
                dx[s] = p.slot_dx[s];
                const int w = p.slot_id[s] & ~ad4cl::kEndFlag;
                word[s] = w < nparams ? w : -1;
            }
            out[id] = ad_commit_p(gs, &p, r);
        } else {
            struct ad_gradient_structure pgs;
            pad_init(1, &pgs, gs, gradient_stack);
            const int before = gs->tape->next_id;
            if (i0 == 0) {
#define F(nm) ad_##nm
#define G gs
                switch (op) { PROBE_ALL }
#undef F
#undef G
            } else {
#define F(nm) pad_##nm
#define G (&pgs)
                switch (op) { PROBE_ALL }
#undef F
#undef G
            }
            k = gs->tape->next_id - before;
            for (int s = 0; s < k; s++) {
                int w;
                ad4cl_tape_peek(gs, k - 1 - s, &dx[s], &w);
                word[s] = (w & ad4cl::kParamFlag) ? (w & ad4cl::kIdMask) : -1;
            }
            out[id] = r;
        }
        double* o = d1 + 8 * (size_t) id;
        o[0] = r.value;
        o[1] = (double) k;
        o[2] = dx[0];
        o[3] = (double) word[0];
        o[4] = dx[1];
        o[5] = (double) word[1];
        o[6] = r.id >= nparams ? 1.0 : 0.0;
        o[7] = lhs.value;
    }
}
