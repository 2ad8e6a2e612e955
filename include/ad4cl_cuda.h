/*
 * ad4cl_cuda.h -- C ABI of libad4cl_cuda.so, the B200 (sm_100a) replacement for
 * the OpenCL plumbing + host adjoint sweep that every AD4CL example hand-writes
 * around a launch.  Plain C: opaque handles, pointers and sizes only.
 *
 * Reference interface each entry point replaces (paths under the AD4CL tree):
 *
 *   ad4cl_cuda_init / _shutdown        platform, context, queue set-up
 *                                      test/admb/simple.cpp:111-200, main.cpp:159-250
 *   ad4cl_cuda_kernel_builtin          cl::Program build of ad.cl + user file, cl::Kernel
 *   ad4cl_cuda_kernel_from_source      test/admb/simple.cpp:138-175, 232  (NVRTC here)
 *   ad4cl_cuda_buffer_*                cl::Buffer + enqueueWrite/ReadBuffer
 *                                      test/admb/simple.cpp:206-230, 283-286
 *   ad4cl_cuda_kernel_set_arg_*        cl::Kernel::setArg, test/admb/simple.cpp:240-262,
 *                                      test/matrix/matrix_mul.cpp:184-192
 *   ad4cl_cuda_kernel_configure        global/local size, stack size
 *                                      test/admb/simple.cpp:235-238, simple.dat:6
 *   ad4cl_cuda_eval                    ONE objective+gradient evaluation, i.e. the whole of
 *                                      test/admb/simple.cpp:283-384: write gs/params ->
 *                                      enqueueNDRangeKernel -> read gs, stack, out ->
 *                                      gpu_restore (ad4cl.h:83-87) -> host sum of out[]
 *                                      (simple.cpp:357-361) -> epilogue (simple.cpp:365) ->
 *                                      compute_gradient (ad4cl.h:775-802) -> g[param.id]
 *   ad4cl_cuda_eval_device             same, parameters and result stay in device memory
 *                                      (multi-GPU drivers all-reduce the result in place)
 *   ad4cl_cuda_last_error              cl::Error::what() + build log, simple.cpp:264-266
 *
 * Every function returns AD4CL_OK (0) or a negative AD4CL_ERR_* code; the text
 * of the last failure on the calling thread is ad4cl_cuda_last_error().  There
 * is no CPU fallback: without a CUDA device ad4cl_cuda_init fails.
 *
 * Threading: calls on one context must be serialised by the caller (the
 * reference's host API is single-threaded, ad4cl.h:28-32); distinct contexts
 * are independent.
 */
#ifndef AD4CL_CUDA_H
#define AD4CL_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ad4cl_context ad4cl_context;
typedef struct ad4cl_kernel ad4cl_kernel;
typedef struct ad4cl_buffer ad4cl_buffer;
typedef struct ad4cl_comm ad4cl_comm;

enum {
    AD4CL_OK = 0,
    AD4CL_ERR_CUDA = -1,          /* a CUDA runtime / driver call failed */
    AD4CL_ERR_INVALID = -2,       /* bad argument, unknown kernel, unset kernel argument */
    AD4CL_ERR_TAPE_OVERFLOW = -3, /* a work-item recorded more than the configured tape holds */
    AD4CL_ERR_WINDOW = -4,        /* far operand met with the far-adjoint plane disabled */
    AD4CL_ERR_COMPILE = -5,       /* NVRTC rejected the source; the log is in last_error */
    AD4CL_ERR_NOMEM = -6
};

/* where the adjoints of the independents are accumulated during the sweep */
enum {
    AD4CL_ACC_AUTO = -1,
    AD4CL_ACC_THREAD = 0,   /* per-thread shared-memory accumulators + block tree */
    AD4CL_ACC_WARP = 1,     /* warp-shuffle reduction per contribution */
    AD4CL_ACC_GLOBAL = 2    /* fp64 atomics into a gradient array (many independents) */
};

/* tape tier */
enum {
    AD4CL_TAPE_AUTO = -1,
    AD4CL_TAPE_HBM = 0,     /* per-warp streams in an HBM arena (L2-resident when small) */
    AD4CL_TAPE_SHARED = 1,  /* the same streams in shared memory (short tapes) */
    AD4CL_TAPE_REGISTER = 2 /* short straight-line objectives: tape, adjoints and gradient accumulators are
                               REGISTERS of the thread (an instance of the kernel compiled for the number of
                               independents and the loop-bounding integer argument; the intent of the
                               reference's private stacks, ad.cl:51-57, 91-97).  AUTO picks it whenever such
                               an instance exists and was fully scalarised, else AD4CL_TAPE_HBM. */
};

/* L2 prefetch ahead of the sweep: the rows two batches below the ones being fetched (LANE recorder), the cells
   three batches below (UNIFORM recorder; a cell's address does not depend on its header).  Helps a sweep that
   waits on HBM latency (tape_stress 10^7 x 10^3: 35.8 -> 34.3 ms), costs one that is bound elsewhere a few
   percent; results are bit-identical either way.  AUTO measures. */
enum {
    AD4CL_PREFETCH_AUTO = -1,
    AD4CL_PREFETCH_OFF = 0,
    AD4CL_PREFETCH_ON = 1
};

/* How a work-item's tape is recorded in the arena / shared-memory tiers (ad4cl_tape.cuh):
     UNIFORM  optimistic: ONE 8-byte header per warp and slot + an 8-byte cell per lane only for partials that
              are not +1/-1 (a third of the bytes; short tapes never leave L2); a warp whose lanes turn out to
              record different tapes re-records the work-item with the LANE recorder.
     LANE     the round-1 stream, 12 bytes per lane and slot: any control flow, any ids.
     AUTO     the first gradient evaluation of a configured kernel times the candidates -- UNIFORM (in both
              launch shapes when the work-group size was left open), LANE, LANE + prefetch, then the winner
              without the lock-step barrier -- twice each, once, and keeps the fastest.  Never changes a
              result beyond the order in which an independent's contributions are summed over a warp
              (bit-identical with <= 16 independents). */
enum {
    AD4CL_RECORDER_AUTO = -1,
    AD4CL_RECORDER_UNIFORM = 0,
    AD4CL_RECORDER_LANE = 1
};

/* Lock step: the warps of a work-group wait for each other before every work-item.  The objective body of
   a statistical model is tens of KB of straight-line code, more than the 32 KB L1.5 instruction cache;
   24 warps at 24 different places in it starve on instruction fetch (ncu: stall_no_instruction), warps
   that walk it together share their fetches.  Costs a barrier per work-item: AUTO measures it. */
enum {
    AD4CL_LOCKSTEP_AUTO = -1,
    AD4CL_LOCKSTEP_OFF = 0,
    AD4CL_LOCKSTEP_ON = 1
};

/* scalar epilogue applied to S = sum_i out_i */
enum {
    AD4CL_EPILOGUE_SUM = 0,      /* f = S */
    AD4CL_EPILOGUE_HALF_N_LOG = 1 /* f = (n/2) log S      (simple.cpp:365, main.cpp:417) */
};

typedef struct ad4cl_kernel_config {
    long long global_size[2];  /* NDRange; global_size[1] = 1 for 1-D */
    int local_size[2];         /* work-group; product <= 768 and a multiple of 32; {0,0} = runtime's choice: 256
                                  (16 x 16 for 2-D), or one 768-thread group per SM if that measures faster */
    int max_ops;               /* most AD operators one work-item records */
    int max_slots;             /* most operand slots (sum of k) one work-item records; 0 = 2*max_ops */
    int window;                /* adjoint ring entries (power of two); 0 = runtime's choice */
    int far_plane;             /* 1: allow operands older than `window` operators (default 1) */
    int acc_mode;              /* AD4CL_ACC_* */
    int tape_tier;             /* AD4CL_TAPE_* */
    int epilogue;              /* AD4CL_EPILOGUE_* */
    double epilogue_n;         /* n of the epilogue; 0 = number of data work-items */
    int blocks_per_sm;         /* persistent grid = SMs x this; 0 = occupancy maximum */
    int reference_quirks;      /* builtin kernels only: 1 = reference arithmetic as written */
    int sweep_prefetch;        /* AD4CL_PREFETCH_* */
    int recorder;              /* AD4CL_RECORDER_* */
    int lockstep;              /* AD4CL_LOCKSTEP_* */
} ad4cl_kernel_config;

typedef struct ad4cl_eval_stats {
    double ops;             /* AD operators recorded by the evaluation */
    double slots;           /* operand slots recorded (12 B each, written once, read once) */
    int grid, block;        /* launch shape used */
    size_t smem_bytes;
    size_t arena_bytes;
    float kernel_ms;        /* device time of the last evaluation (both launches) */
    int sweep_prefetch;     /* the setting in use: 0 off, 1 on, -1 not decided yet (AUTO before the first gradient evaluation) */
    int tape_tier;          /* the AD4CL_TAPE_* tier in use (what AUTO resolved to) */
    int recorder;           /* the AD4CL_RECORDER_* in use, -1 not decided yet */
    int lockstep;           /* 1: work-groups run in lock step */
} ad4cl_eval_stats;

int ad4cl_cuda_init(int device, ad4cl_context** ctx);
int ad4cl_cuda_shutdown(ad4cl_context* ctx);
const char* ad4cl_cuda_last_error(void);
/* identifies the device-side headers this library was built with; a cubin from ad4cl_cuda_compile
   is only valid for the library that reports the same string */
const char* ad4cl_cuda_device_abi(void);
/* the CUDA stream of the context, as a cudaStream_t cast to void* */
void* ad4cl_cuda_stream(ad4cl_context* ctx);
/* run on a caller-owned stream (e.g. torch's current stream); NULL restores the context's own */
int ad4cl_cuda_use_stream(ad4cl_context* ctx, void* cuda_stream);
int ad4cl_cuda_device_sm_count(ad4cl_context* ctx);

int ad4cl_cuda_buffer_create(ad4cl_context* ctx, size_t bytes, ad4cl_buffer** buf);
/* wrap caller-owned device memory (never freed by the library) */
int ad4cl_cuda_buffer_wrap(ad4cl_context* ctx, void* device_ptr, size_t bytes, ad4cl_buffer** buf);
int ad4cl_cuda_buffer_upload(ad4cl_buffer* buf, size_t offset, const void* host, size_t bytes);
int ad4cl_cuda_buffer_download(ad4cl_buffer* buf, size_t offset, void* host, size_t bytes);
void* ad4cl_cuda_buffer_device_ptr(ad4cl_buffer* buf);
int ad4cl_cuda_buffer_free(ad4cl_buffer* buf);

/* a kernel compiled ahead of time into the library (ad4cl_b200/kernels/\*.cl) */
int ad4cl_cuda_kernel_builtin(ad4cl_context* ctx, const char* name, int reference_quirks,
        ad4cl_kernel** kernel);
/* number of built-in kernels; name/signature of the i-th */
int ad4cl_cuda_builtin_count(void);
const char* ad4cl_cuda_builtin_name(int i);
const char* ad4cl_cuda_builtin_signature(int i);
/*
 * Compile `source` (OpenCL-C-dialect kernel text written against ad.cl) with
 * NVRTC for sm_100a.  `signature` gives the kinds of `entry`'s parameters in
 * order:  G gs  S gradient_stack  O out  V ad_variable*  D double*  I int*
 * i int  d double   (e.g. "GSVVDDOi" for test/admb/simple.cl).  `options` is a
 * space-separated list of extra compiler options (may be NULL).
 */
int ad4cl_cuda_kernel_from_source(ad4cl_context* ctx, const char* source, const char* entry,
        const char* signature, const char* options, int reference_quirks, ad4cl_kernel** kernel);
/*
 * The two halves of _from_source: compile without touching a device (returns a malloc'd
 * sm_100a cubin, release with ad4cl_cuda_free_host), and load a cubin produced earlier.
 */
int ad4cl_cuda_compile(const char* source, const char* entry, const char* signature, const char* options,
        int reference_quirks, void** cubin, size_t* cubin_bytes);
void ad4cl_cuda_free_host(void* p);
int ad4cl_cuda_kernel_from_cubin(ad4cl_context* ctx, const void* cubin, size_t cubin_bytes, const char* entry,
        const char* signature, int reference_quirks, ad4cl_kernel** kernel);
int ad4cl_cuda_kernel_free(ad4cl_kernel* kernel);
const char* ad4cl_cuda_kernel_signature(ad4cl_kernel* kernel);

/* arguments are indexed like the user kernel's parameter list (cl::Kernel::setArg);
   G, S and O positions accept any value and ignore it */
int ad4cl_cuda_kernel_set_arg_buffer(ad4cl_kernel* kernel, int index, ad4cl_buffer* buf);
int ad4cl_cuda_kernel_set_arg_int(ad4cl_kernel* kernel, int index, int value);
int ad4cl_cuda_kernel_set_arg_double(ad4cl_kernel* kernel, int index, double value);
/* bind a V argument to `count` consecutive independents starting at id `first_id`;
   ad4cl_cuda_eval fills them from theta[first_id .. first_id+count) */
int ad4cl_cuda_kernel_set_arg_params(ad4cl_kernel* kernel, int index, int first_id, int count);

int ad4cl_cuda_kernel_config_default(ad4cl_kernel_config* cfg);
int ad4cl_cuda_kernel_configure(ad4cl_kernel* kernel, const ad4cl_kernel_config* cfg);

/*
 * One evaluation.  theta: nparams doubles (host).  On return *f is the
 * objective and grad[0..nparams) its gradient (grad may be NULL: objective
 * only, nothing is recorded).  Crossing PCIe: 8*nparams bytes in,
 * 8*(nparams+3) bytes out.
 */
int ad4cl_cuda_eval(ad4cl_kernel* kernel, const double* theta, int nparams, double* f, double* grad);
/*
 * Same with device-resident parameters and result, asynchronous on the
 * context's stream: theta_dev holds nparams doubles; result_dev receives
 * nparams+3 doubles [grad..., S_or_f, ops, slots].  With apply_epilogue = 0 the
 * raw sums are left (so partial results of several devices can be added before
 * the epilogue is applied).  Errors raised by the launch itself surface at the
 * next ad4cl_cuda_check().  ONE kernel launch per call: parameter packing, record + sweep,
 * both reduction stages and the epilogue run inside the user kernel's entry (with
 * AD4CL_ACC_GLOBAL, i.e. the matrix configs, packing and the second stage are separate
 * small launches).  Capturable into a CUDA graph.
 */
int ad4cl_cuda_eval_device(ad4cl_kernel* kernel, const double* theta_dev, int nparams,
        double* result_dev, int want_gradient, int apply_epilogue);
/* the same, predicated on device memory: skipped when *skip_dev != 0 at EXECUTION time (for evaluations
   captured in a CUDA graph behind a device-side decision, e.g. an optimiser that has converged) */
int ad4cl_cuda_eval_device_if(ad4cl_kernel* kernel, const double* theta_dev, int nparams,
        double* result_dev, int want_gradient, int apply_epilogue, const int* skip_dev);
ad4cl_context* ad4cl_cuda_kernel_context(ad4cl_kernel* kernel);
/* synchronise the stream and report tape overflow / window errors of queued evaluations */
int ad4cl_cuda_check(ad4cl_kernel* kernel);
int ad4cl_cuda_last_stats(ad4cl_kernel* kernel, ad4cl_eval_stats* stats);
/*
 * Tape export -- the debug / legacy path (what enqueueReadBuffer of the whole stack gave the
 * reference's host, test/admb/simple.cpp:315).  Records the kernel WITHOUT sweeping and rewrites
 * every work-item's tape as the reference's 40-byte ad_entry records (ad.cl:75-79), in work-item
 * order, into entries_out (capacity entries; *n_entries receives the count).  out_vars (n_out
 * 16-byte ad_variables, may be NULL) receives out[i] with its id in the exported numbering:
 * independents keep their ids, temporaries are numbered from first_free_id on.  Legacy host code can
 * continue as in simple.cpp:357-369: gs->stack_current += n, gs->current_variable_id = first_free_id
 * + n, host sum of out[], epilogue, compute_gradient.  Small problems only.
 */
int ad4cl_cuda_export_tape(ad4cl_kernel* kernel, const double* theta, int nparams, int first_free_id,
        void* entries_out, long long capacity, long long* n_entries, void* out_vars, long long n_out);

/*
 * Opt-in device timing of the fused record+sweep+reduce kernel ALONE (CUDA events
 * recorded on the launching stream around that one launch; the analogue of the
 * reference's CL_PROFILING event.getProfilingInfo, test/admb/simple.cpp:301-310).
 * _read synchronises the stream and returns the mean over the launches seen.
 */
int ad4cl_cuda_kernel_profile(ad4cl_kernel* kernel, int enable);
int ad4cl_cuda_kernel_profile_read(ad4cl_kernel* kernel, double* mean_ms, long long* launches, int reset);

/*
 * L-BFGS over ad4cl_cuda_eval: the optimiser the reference declares for its device runtime
 * (struct lbfgs_parameters_g / lbfgs_update_g, ad.cl:99-127) and never defines; its examples
 * are driven by ADMB's function_minimizer instead (test/admb/simple.cpp:481-492).  Field names
 * follow the reference's struct.  theta: start point in, minimiser out.
 */
typedef struct ad4cl_lbfgs_options {
    int max_iterations;              /* ad.cl:106 */
    int max_linesearch_iterations;   /* ad.cl:108 */
    int history;                     /* correction pairs kept (m) */
    double gradient_tolerance;       /* stop when max |g_i| <= this */
    double function_tolerance;       /* stop when f decreases by <= this * max(1, |f|) */
} ad4cl_lbfgs_options;

typedef struct ad4cl_lbfgs_result {
    double f;                        /* objective at the returned point */
    double gradient_max;             /* max |g_i| there */
    int converged;                   /* ad.cl:105: 1 gradient, 2 function tolerance, 0 iteration limit, -1 line search failed */
    int iterations;
    int linesearch_iteration;        /* ad.cl:107: line-search evaluations in total */
    int evaluations;                 /* objective+gradient evaluations (ad4cl_cuda_eval calls) */
    int number_of_parameters;        /* ad.cl:110 */
} ad4cl_lbfgs_result;

int ad4cl_cuda_lbfgs_options_default(ad4cl_lbfgs_options* options);
int ad4cl_cuda_minimize_lbfgs(ad4cl_kernel* kernel, double* theta, int number_of_parameters,
        const ad4cl_lbfgs_options* options, ad4cl_lbfgs_result* result);
/*
 * The same minimiser DEVICE RESIDENT (what lbfgs_parameters_g / lbfgs_update_g,
 * ad.cl:99-127, were declared for): parameters, history pairs and the line-search state live in device
 * memory; between two evaluations ONE single-CTA kernel (lbfgs_update) takes the Armijo decision, updates
 * the history, runs the two-loop recursion and writes the next trial point where the next evaluation reads
 * it.  The fit is replayed as CUDA graphs of 128 (evaluation, update) pairs: no host synchronisation and no
 * PCIe traffic between evaluations; once converged the remaining evaluations of the graph skip themselves
 * (ad4cl_cuda_eval_device_if).  Same options, same result fields; theta: start point in, minimiser out.
 */
int ad4cl_cuda_minimize_lbfgs_device(ad4cl_kernel* kernel, double* theta, int number_of_parameters,
        const ad4cl_lbfgs_options* options, ad4cl_lbfgs_result* result);

/*
 * Dense-contraction path for test/matrix (matrix_mul.cpp:39-57, matrixmul.cl:5-28): instead of
 * recording 2n AD operators per element of C = A*B, compute C, sum C and the derivative of
 * L = sum_{j,i} G[j,i] C[j,i] as two more products, dL/dA = G*B^T and dL/dB = A^T*G.
 * Row-major n x n doubles in device memory, n a multiple of 64.  G_dev NULL means G = 1 (the
 * reference's objective).  C_dev, dA_dev, dB_dev, sum_dev may each be NULL to skip that output.
 * workspace_dev: n*n + (n/64)^2 doubles.  variant 0: CUDA-core DFMA, 1: fp64 tensor-core DMMA (one launch per
 * product), 2: DMMA, every requested product in ONE persistent launch (tiles drawn from a counter; G = NULL costs no fill).
 * Asynchronous on the context's stream.
 */
int ad4cl_cuda_matrix_product_dense(ad4cl_context* ctx, int variant, int n, const double* A_dev,
        const double* B_dev, const double* G_dev, double* C_dev, double* dA_dev, double* dB_dev,
        double* sum_dev, double* workspace_dev);

/* measured fp64 FMA rate of the device in TFLOP/s: the denominator of the dense path's roofline */
int ad4cl_cuda_measure_fp64_peak(ad4cl_context* ctx, double* tflops);

/*
 * Multi-GPU (one process per GPU; no counterpart in the reference, which is single device):
 * a one-shot all-reduce of the short result vector [grad..., S, ops, slots] over NVLink peer
 * memory, summed in rank order (bit-identical on every rank), with the scalar epilogue fused.
 *   create   allocate this rank's inbox (max_doubles per peer)
 *   handle   cudaIpcMemHandle_t of the inbox (ad4cl_cuda_comm_handle_bytes() bytes); the caller
 *            exchanges the handles of all ranks with whatever it has (torch.distributed, MPI)
 *   connect  handles of ALL ranks, in rank order
 *   allreduce  in place on data_dev (n doubles), asynchronous on the context's stream; with
 *            epilogue = AD4CL_EPILOGUE_HALF_N_LOG, data[nparams] is S and the epilogue is applied
 *   check    synchronise and report a peer that never arrived
 */
int ad4cl_cuda_comm_create(ad4cl_context* ctx, int rank, int world, int max_doubles, ad4cl_comm** comm);
/* all `world` ranks in ONE process (a C++ host driving every GPU of the box, examples/main_cuda.cpp):
   contexts of distinct devices, peers reached by plain peer access, no handle exchange; comms_out
   receives `world` connected communicators, rank r on ctxs[r] */
int ad4cl_cuda_comm_create_local(ad4cl_context** ctxs, int world, int max_doubles, ad4cl_comm** comms_out);
/* how long a rank waits for its peers before the evaluation fails (default 30 s).  A timed-out
   all-reduce leaves NaN in the reduced vector, ad4cl_cuda_eval_sharded / _comm_check return an error,
   and the communicator must be destroyed and rebuilt (its epochs are no longer aligned). */
int ad4cl_cuda_comm_set_timeout(ad4cl_comm* comm, double seconds);
int ad4cl_cuda_comm_handle_bytes(void);
int ad4cl_cuda_comm_handle(ad4cl_comm* comm, void* handle_out);
int ad4cl_cuda_comm_connect(ad4cl_comm* comm, const void* handles);
int ad4cl_cuda_comm_allreduce(ad4cl_comm* comm, double* data_dev, int n, int nparams, int epilogue, double n_obs);
int ad4cl_cuda_comm_check(ad4cl_comm* comm);
int ad4cl_cuda_comm_destroy(ad4cl_comm* comm);
/*
 * One SHARDED evaluation in one call (SURVEY.md 8(b) `..._eval_sharded`): `kernel` is bound to this
 * rank's shard of the observations; theta (host, nparams doubles, the same on every rank) goes up,
 * the kernel's last CTA all-reduces the raw sums [grad, S, ops, slots] with the peers of `comm` over
 * NVLink and applies the epilogue, f and grad[nparams] (identical bits on every rank) come down.
 * Replayed as ONE CUDA graph: H2D, one kernel, D2H.  Every rank of the communicator must call it
 * the same number of times.  _device_ form: theta and result ([grad, f, ops, slots]) stay in device
 * memory, asynchronous on the context's stream.
 */
int ad4cl_cuda_eval_sharded(ad4cl_kernel* kernel, ad4cl_comm* comm, const double* theta, int nparams,
        double* f, double* grad);
int ad4cl_cuda_eval_device_sharded(ad4cl_kernel* kernel, ad4cl_comm* comm, const double* theta_dev, int nparams,
        double* result_dev, int want_gradient);
/*
 * General epilogues / non-uniform seeds.  The reference's sweep seeds only the LAST tape
 * entry (ad4cl.h:786) and its examples only ever reduce with f = h(sum_i out_i) (simple.cpp:357-365).  For
 * an objective f = h(out_1, .., out_n) that is not of that form, evaluate in two passes -- nothing is
 * retained between them, the tape still never leaves the SM / L2:
 *   pass 1   out_values_dev != NULL, grad == NULL: every work-item stores out_i.value (indexed like the
 *            kernel's own out[]; device memory, n doubles).  The caller computes the seeds
 *            s_i = dh / d out_i from them (on the device or the host).
 *   pass 2   seeds_dev != NULL, grad != NULL: work-item i's sweep starts from s_i instead of 1, so
 *            grad = sum_i s_i * d out_i / d theta = the gradient of f.
 * *sum receives sum_i out_i either way (no epilogue is applied).  theta and grad are host memory.
 */
int ad4cl_cuda_eval_general(ad4cl_kernel* kernel, const double* theta, int nparams, double* out_values_dev,
        const double* seeds_dev, double* sum, double* grad);
/* kernel launches the last evaluation of this kernel enqueued (1 unless AD4CL_ACC_GLOBAL) */
int ad4cl_cuda_launches_per_eval(ad4cl_kernel* kernel);

#ifdef __cplusplus
}
#endif

#endif /* AD4CL_CUDA_H */
