/*
 * objectives.cl -- user objective kernels written in the ad.cl dialect.
 *
 * These are the likelihood terms of BASELINE.json configs 1, 3, 4, 5 expressed
 * exactly the way test/admb/simple.cl:106-141 expresses its own: a __kernel
 * that receives (gs, gradient_stack, parameters, data, out, size), calls the
 * ad_* / pad_* operators of the device AD runtime, and stores one ad_variable
 * per work-item into out[].  The same text is compiled
 *   - by nvcc / NVRTC against ad4cl_b200/csrc/ad.cuh  (the product), and
 *   - by gcc against the reference's own ad.cl, or against oracle/ad_port.h
 *     (the CPU oracle; tests and the CPU baseline only).
 * Nothing here is specific to either side: plain C99 that is also valid C++.
 *
 * Common signature ("objective ABI") for every kernel except linreg2:
 *   theta : P independent ad_variables, ids 0..P-1
 *   d0,d1 : two fp64 data arrays
 *   out   : one ad_variable per work-item (the term that is summed)
 *   size  : number of work-items that carry data (global size may be larger)
 *   i0,i1 : two integer shape arguments
 */

/* ------------------------------------------------------------------------
 * config 1: sum of squares term of the shipped example (kernel.cl:17-25,
 * test/admb/simple.cl:123-137): out = ((a*x + b) - y)^2, 4 AD ops.
 * linreg2_g uses the global-stack family, linreg2_p the block-reserved one.
 * --------------------------------------------------------------------- */
__kernel void linreg2_g(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* a,
        __global struct ad_variable* b,
        __global double* x,
        __global double* y,
        __global struct ad_variable* out, int size) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_variable pred = ad_plus(gs, ad_times_vd(gs, *a, x[id]), *b);
        struct ad_variable res = ad_minus_vd(gs, pred, y[id]);
        out[id] = ad_times(gs, res, res);
    }
}

__kernel void linreg2_p(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __constant struct ad_variable* a,
        __constant struct ad_variable* b,
        __global double* x,
        __global double* y,
        __global struct ad_variable* out, int size) {
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_gradient_structure pgs;
        ad_init(gs, gradient_stack);
        pad_init(4, &pgs, gs, gradient_stack);
        struct ad_variable pred = pad_plus(&pgs, pad_times_vd(&pgs, *a, x[id]), *b);
        struct ad_variable res = pad_minus_vd(&pgs, pred, y[id]);
        out[id] = pad_times(&pgs, res, res);
    }
}

/* ------------------------------------------------------------------------
 * config 3: regression with P = i0 parameters, X stored [P][size]
 * (observation-contiguous), y = d1.
 *   linear  : out = (sum_j theta_j X_ji - y_i)^2            21 ops at P=10
 *   logistic: out = log(1 + exp(eta)) - y_i * eta           24 ops at P=10
 * --------------------------------------------------------------------- */
__kernel void regression_linear(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_variable eta = ad_times_vd(gs, theta[0], d0[id]);
        for (int j = 1; j < i0; j++) {
            eta = ad_plus(gs, eta, ad_times_vd(gs, theta[j], d0[(size_t) j * size + id]));
        }
        struct ad_variable res = ad_minus_vd(gs, eta, d1[id]);
        out[id] = ad_times(gs, res, res);
    }
}

__kernel void regression_logistic(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_variable eta = ad_times_vd(gs, theta[0], d0[id]);
        for (int j = 1; j < i0; j++) {
            eta = ad_plus(gs, eta, ad_times_vd(gs, theta[j], d0[(size_t) j * size + id]));
        }
        struct ad_variable softplus = ad_log(gs, ad_plus_vd(gs, ad_exp(gs, eta), 1.0));
        out[id] = ad_minus(gs, softplus, ad_times_vd(gs, eta, d1[id]));
    }
}

/* ------------------------------------------------------------------------
 * config 4: catch-at-age negative log-likelihood term, P = 100.
 *   work-item i = rep_local + R*(year + CAA_YEARS*age), R = i0 replicates of every
 *   (year, age) cell held by this shard, so the replicate index is fastest and
 *   (year, age) -- hence the whole op sequence and every parameter id except the
 *   fleet deviation -- is uniform across neighbouring work-items.
 *   theta: [0,49) log recruitment of cohort c = year - age + 9
 *          [49,89) log F_year   89: a50   90: slope   91: log M
 *          92: log q   93: log sigma   [94,100) fleet deviations of log q
 *   d0   : observed log catch (this shard's observations)
 *   i1   : how the shard maps its work-items to (cell, replicate):
 *          i1 >= 0  replicate slices: the shard holds replicates [i1, i1 + i0) of EVERY cell
 *                   (i0 = replicates per cell in the shard), so all shards have the same mix
 *                   of tape lengths; one GPU: i1 = 0, i0 = R_global
 *          i1 <  0  whole cells dealt round-robin: with p = -i1 - 1, the shard holds all
 *                   i0 = R_global replicates of cells c0 + s*j, c0 = p % 1024, s = p / 1024
 *                   (rank + world*j): long runs of identical work-items survive the sharding
 *   d1   : d1[0] = R_global, the replicates per cell of the WHOLE problem
 *   fleet of a replicate = rep*6/R_global  (6 contiguous blocks of replicates: the fleet
 *          deviation's parameter id is warp-uniform except at the 5 block boundaries)
 * Baranov: Z = M + F_y s_a ; N = R_c exp(-sum_{k<a} Z_{y-a+k,k}) ;
 *          Chat = F s / Z (1 - exp(-Z)) N q ;
 *          out = (ln C - ln Chat)^2 / (2 sigma^2) + ln sigma
 * --------------------------------------------------------------------- */
#define CAA_YEARS 40
#define CAA_AGES 10
#define CAA_LOGF 49
#define CAA_A50 89
#define CAA_SLOPE 90
#define CAA_LOGM 91
#define CAA_LOGQ 92
#define CAA_LOGSIGMA 93
#define CAA_FLEETDEV 94
#define CAA_FLEETS 6

__kernel void catch_at_age(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        const int R = i0;
        const int rep = i1 >= 0 ? i1 + id % R : id % R;
        const int cell = i1 >= 0 ? id / R : (-i1 - 1) % 1024 + ((-i1 - 1) / 1024) * (id / R);
        const int year = cell % CAA_YEARS;
        const int age = cell / CAA_YEARS;
        const int cohort = year - age + (CAA_AGES - 1);
        const int fleet = (int) (((long) rep * CAA_FLEETS) / (long) d1[0]); /* `long` is 64-bit in OpenCL C, CUDA and LP64 C */

        struct ad_variable M = ad_exp(gs, theta[CAA_LOGM]);
        struct ad_variable Z = M;
        struct ad_variable Fs = M;
        struct ad_variable cumZ = M;
        int have_cum = 0;
        /* walk the cohort from age 0 up to this age; the last pass (k == age)
           leaves this cell's own Z and F*s */
        for (int k = 0; k <= age; k++) {
            const int yy = year - age + k;
            if (yy >= 0) {
                struct ad_variable arg = ad_times(gs, theta[CAA_SLOPE],
                        ad_minus_dv(gs, (double) k, theta[CAA_A50]));
                struct ad_variable sel = ad_divide_dv(gs, 1.0,
                        ad_plus_vd(gs, ad_exp(gs, ad_times_vd(gs, arg, -1.0)), 1.0));
                Fs = ad_times(gs, ad_exp(gs, theta[CAA_LOGF + yy]), sel);
                Z = ad_plus(gs, M, Fs);
            } else {
                Z = M; /* no fishery before the first year */
            }
            if (k < age) {
                if (have_cum) {
                    cumZ = ad_plus(gs, cumZ, Z);
                } else {
                    cumZ = Z;
                    have_cum = 1;
                }
            }
        }
        struct ad_variable logN = theta[cohort];
        if (have_cum) {
            logN = ad_minus(gs, logN, cumZ);
        }
        struct ad_variable N = ad_exp(gs, logN);
        struct ad_variable surv = ad_minus_dv(gs, 1.0, ad_exp(gs, ad_times_vd(gs, Z, -1.0)));
        struct ad_variable q = ad_exp(gs, ad_plus(gs, theta[CAA_LOGQ], theta[CAA_FLEETDEV + fleet]));
        struct ad_variable chat = ad_times(gs,
                ad_times(gs, ad_times(gs, ad_divide(gs, Fs, Z), surv), N), q);
        struct ad_variable res = ad_minus_dv(gs, d0[id], ad_log(gs, chat));
        struct ad_variable sigma = ad_exp(gs, theta[CAA_LOGSIGMA]);
        struct ad_variable two_s2 = ad_times_vd(gs, ad_times(gs, sigma, sigma), 2.0);
        out[id] = ad_plus(gs, ad_divide(gs, ad_times(gs, res, res), two_s2), theta[CAA_LOGSIGMA]);
    }
}

/* ------------------------------------------------------------------------
 * config 5: tape-length stress.  steps = i0, P = i1, x = d0.
 *   v_0 = theta_0 * x ;  v_{j+1} = 0.5 v_j + theta_{j mod P} x   (3 ops/step)
 *   out = v_steps                       L = 3*steps + 1 ops, 4*steps + 1 slots
 * --------------------------------------------------------------------- */
__kernel void tape_stress(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        const double x = d0[id];
        struct ad_variable v = ad_times_vd(gs, theta[0], x);
        int p = 0;
        for (int j = 0; j < i0; j++) {
            v = ad_plus(gs, ad_times_vd(gs, v, 0.5), ad_times_vd(gs, theta[p], x));
            p = (p + 1 == i1) ? 0 : p + 1;
        }
        out[id] = v;
    }
}

/* ------------------------------------------------------------------------
 * op coverage kernel (parity tests only): one work-item = one chain through
 * every operator of the global-stack family, on data-dependent operands, so
 * that a single launch compares every value and every partial against the
 * reference.  theta has 3 parameters; d0 in (0.1, 0.9).
 * --------------------------------------------------------------------- */
__kernel void op_zoo(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        const double x = d0[id];
        struct ad_variable a = ad_times_vd(gs, theta[0], x);          /* in (0,1) */
        struct ad_variable b = ad_plus_vd(gs, theta[1], x);           /* > 1 */
        struct ad_variable c = ad_plus_dv(gs, 0.25, theta[2]);
        struct ad_variable acc = ad_plus(gs, a, b);
        acc = ad_minus(gs, acc, c);
        acc = ad_minus_vd(gs, acc, 0.125);
        acc = ad_minus_dv(gs, 9.0, acc);
        acc = ad_times(gs, acc, b);
        acc = ad_times_dv(gs, 0.75, acc);
        acc = ad_divide(gs, acc, b);
        acc = ad_divide_vd(gs, acc, 1.5);
        acc = ad_plus(gs, acc, ad_divide_dv(gs, 2.0, b));
        ad_plus_eq(gs, &acc, a);
        ad_plus_eq_d(gs, &acc, 0.5);
        acc = ad_plus(gs, acc, ad_sin(gs, a));
        acc = ad_plus(gs, acc, ad_cos(gs, b));
        acc = ad_plus(gs, acc, ad_tan(gs, a));
        acc = ad_plus(gs, acc, ad_asin(gs, a));
        acc = ad_plus(gs, acc, ad_acos(gs, a));
        acc = ad_plus(gs, acc, ad_atan(gs, b));
        acc = ad_plus(gs, acc, ad_sinh(gs, a));
        acc = ad_plus(gs, acc, ad_cosh(gs, a));
        acc = ad_plus(gs, acc, ad_tanh(gs, b));
        acc = ad_plus(gs, acc, ad_exp(gs, a));
        acc = ad_plus(gs, acc, ad_log(gs, b));
        acc = ad_plus(gs, acc, ad_log10(gs, b));
        acc = ad_plus(gs, acc, ad_sqrt(gs, b));
        acc = ad_plus(gs, acc, ad_pow(gs, b, a));
        acc = ad_plus(gs, acc, ad_pow_vd(gs, b, 1.75));
        acc = ad_plus(gs, acc, ad_pow_dv(gs, 1.5, a));
        out[id] = acc;
    }
}

/* the same chain through the block-reserved family (incl. the pad_ math functions this
 * runtime adds; with the reference runtime or the port only the arithmetic part exists,
 * so the oracle compares against op_zoo of the global family: same values, and the same
 * partials except pad_times_dv, which is correct in the reference, ad.cl:840) */
#ifdef AD4CL_HAVE_PAD_MATH
__kernel void op_zoo_pad(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_gradient_structure pgs;
        ad_init(gs, gradient_stack);
        pad_init(47, &pgs, gs, gradient_stack);
        const double x = d0[id];
        struct ad_variable a = pad_times_vd(&pgs, theta[0], x);
        struct ad_variable b = pad_plus_vd(&pgs, theta[1], x);
        struct ad_variable c = pad_plus_dv(&pgs, 0.25, theta[2]);
        struct ad_variable acc = pad_plus(&pgs, a, b);
        acc = pad_minus(&pgs, acc, c);
        acc = pad_minus_vd(&pgs, acc, 0.125);
        acc = pad_minus_dv(&pgs, 9.0, acc);
        acc = pad_times(&pgs, acc, b);
        acc = pad_times_dv(&pgs, 0.75, acc);
        acc = pad_divide(&pgs, acc, b);
        acc = pad_divide_vd(&pgs, acc, 1.5);
        acc = pad_plus(&pgs, acc, pad_divide_dv(&pgs, 2.0, b));
        pad_plus_eq(&pgs, &acc, a);
        pad_plus_eq_d(&pgs, &acc, 0.5);
        acc = pad_plus(&pgs, acc, pad_sin(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_cos(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_tan(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_asin(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_acos(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_atan(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_sinh(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_cosh(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_tanh(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_exp(&pgs, a));
        acc = pad_plus(&pgs, acc, pad_log(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_log10(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_sqrt(&pgs, b));
        acc = pad_plus(&pgs, acc, pad_pow(&pgs, b, a));
        acc = pad_plus(&pgs, acc, pad_pow_vd(&pgs, b, 1.75));
        acc = pad_plus(&pgs, acc, pad_pow_dv(&pgs, 1.5, a));
        out[id] = acc;
    }
}
#endif

/* config 1 on a PRIVATE stack: the 4 operators stay in registers / local memory and only the two
 * pre-accumulated partials d(out)/da, d(out)/db reach the thread tape (3 slots instead of 6) */
#ifdef AD4CL_HAVE_COMMIT_P
__kernel void linreg2_private(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __constant struct ad_variable* a,
        __constant struct ad_variable* b,
        __global double* x,
        __global double* y,
        __global struct ad_variable* out, int size) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_private_gradient_structure p;
        ad_init_p(&p);
        p.current_ad_variable_id = gs->current_ad_variable_id;
        struct ad_variable res = ad_minus_vd_p(&p, ad_plus_p(&p, ad_times_vd_p(&p, *a, x[id]), *b), y[id]);
        out[id] = ad_commit_p(gs, &p, ad_times_p(&p, res, res));
    }
}
#endif

/* the same chain on a PRIVATE stack (ad.cl:1216-1829), pre-accumulated into the thread
 * tape with ad_commit_p: only 3 slots per work-item leave the registers / local memory */
#ifdef AD4CL_HAVE_COMMIT_P
__kernel void op_zoo_private(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* theta,
        __global double* d0, __global double* d1,
        __global struct ad_variable* out, int size, int i0, int i1) {
    ad_init(gs, gradient_stack);
    const int id = get_global_id(0);
    if (id < size) {
        struct ad_private_gradient_structure p;
        ad_init_p(&p);
        p.current_ad_variable_id = gs->current_ad_variable_id; /* private results follow the independents */
        const double x = d0[id];
        struct ad_variable a = ad_times_vd_p(&p, theta[0], x);
        struct ad_variable b = ad_plus_vd_p(&p, theta[1], x);
        struct ad_variable c = ad_plus_dv_p(&p, 0.25, theta[2]);
        struct ad_variable acc = ad_plus_p(&p, a, b);
        acc = ad_minus_p(&p, acc, c);
        acc = ad_minus_vd_p(&p, acc, 0.125);
        acc = ad_minus_dv_p(&p, 9.0, acc);
        acc = ad_times_p(&p, acc, b);
        acc = ad_times_dv_p(&p, 0.75, acc);
        acc = ad_divide_p(&p, acc, b);
        acc = ad_divide_vd_p(&p, acc, 1.5);
        acc = ad_plus_p(&p, acc, ad_divide_dv_p(&p, 2.0, b));
        ad_plus_eq_p(&p, &acc, a);
        ad_plus_eq_d_p(&p, &acc, 0.5);
        acc = ad_plus_p(&p, acc, ad_sin_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_cos_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_tan_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_asin_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_acos_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_atan_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_sinh_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_cosh_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_tanh_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_exp_p(&p, a));
        acc = ad_plus_p(&p, acc, ad_log_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_log10_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_sqrt_p(&p, b));
        acc = ad_plus_p(&p, acc, ad_pow_p(&p, b, a));
        acc = ad_plus_p(&p, acc, ad_pow_vd_p(&p, b, 1.75));
        acc = ad_plus_p(&p, acc, ad_pow_dv_p(&p, 1.5, a));
        out[id] = ad_commit_p(gs, &p, acc);
    }
}
#endif
