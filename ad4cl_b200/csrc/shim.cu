/*
 * shim.cu -- implementation of the C ABI declared in include/ad4cl_cuda.h.
 *
 * Host-side orchestration only: resource sizing, argument marshalling, the ONE
 * launch of an evaluation (parameter packing, record + sweep, both reduction
 * stages and -- multi-GPU -- the one-shot all-reduce all happen inside the
 * user kernel's entry, ad4cl_entry.cuh) and the two small PCIe copies.  The
 * arithmetic lives in ad.cuh / ad4cl_tape.cuh / ad4cl_entry.cuh.
 *
 * No CPU path exists here: every entry point needs a CUDA device.
 */
#include "../../include/ad4cl_cuda.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ad4cl_builtin.h"
#include "ad4cl_entry.cuh"
#include "ad4cl_reduce.cuh"

extern "C" const ad4cl_builtin_row ad4cl_builtin_table_fixed[];
extern "C" const ad4cl_builtin_row ad4cl_builtin_table_quirks[];
/* register-tier instances (builtin_static.cu), one table per (number of independents, quirks) */
extern "C" const ad4cl_static_row ad4cl_static_table_fixed_p2[];
extern "C" const ad4cl_static_row ad4cl_static_table_quirks_p2[];
extern "C" const ad4cl_static_row ad4cl_static_table_fixed_p10[];
extern "C" const ad4cl_static_row ad4cl_static_table_quirks_p10[];
extern "C" const char* const ad4cl_embedded_header_names[];
extern "C" const char* const ad4cl_embedded_header_texts[];
extern "C" const int ad4cl_embedded_header_count;
extern "C" const char ad4cl_embedded_header_hash[];

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(AD4CL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                    __FILE__, __LINE__);                                                    \
    } while (0)

/* writes {theta[i], id = i} for the V arguments (ad_init_var, ad4cl.h:90-93) */
__global__ void pack_params(const double* __restrict__ theta, struct ad_variable* __restrict__ params, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        params[i].value = theta[i];
        params[i].id = i;
    }
}

/* ---- lazily bound driver API + NVRTC (keeps the library loadable on a box without a GPU) ---- */
struct DriverApi {
    CUresult (*cuModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*cuModuleUnload)(CUmodule) = nullptr;
    CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
            unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*cuFuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*cuGetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
};

DriverApi& driver() {
    static DriverApi d;
    if (d.ok) return d;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return d;
#define BIND(name, sym) d.name = reinterpret_cast<decltype(d.name)> (dlsym(h, sym))
    BIND(cuModuleLoadData, "cuModuleLoadData");
    BIND(cuModuleUnload, "cuModuleUnload");
    BIND(cuModuleGetFunction, "cuModuleGetFunction");
    BIND(cuLaunchKernel, "cuLaunchKernel");
    BIND(cuFuncSetAttribute, "cuFuncSetAttribute");
    BIND(cuFuncGetAttribute, "cuFuncGetAttribute");
    BIND(cuOccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    BIND(cuGetErrorString, "cuGetErrorString");
#undef BIND
    d.ok = d.cuModuleLoadData && d.cuModuleGetFunction && d.cuLaunchKernel && d.cuFuncSetAttribute
            && d.cuFuncGetAttribute && d.cuOccupancyMaxActiveBlocksPerMultiprocessor;
    return d;
}

typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NvrtcApi& nvrtc() {
    static NvrtcApi n;
    if (n.ok) return n;
    void* h = nullptr;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
        "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return n;
#define BIND(name, sym) n.name = reinterpret_cast<decltype(n.name)> (dlsym(h, sym))
    BIND(CreateProgram, "nvrtcCreateProgram");
    BIND(DestroyProgram, "nvrtcDestroyProgram");
    BIND(CompileProgram, "nvrtcCompileProgram");
    BIND(GetProgramLogSize, "nvrtcGetProgramLogSize");
    BIND(GetProgramLog, "nvrtcGetProgramLog");
    BIND(GetCUBINSize, "nvrtcGetCUBINSize");
    BIND(GetCUBIN, "nvrtcGetCUBIN");
    BIND(GetErrorString, "nvrtcGetErrorString");
#undef BIND
    n.ok = n.CreateProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN && n.GetProgramLogSize
            && n.GetProgramLog && n.DestroyProgram;
    return n;
}

struct Arg {
    char kind = 0;
    bool set = false;
    ad4cl_buffer* buf = nullptr;
    int ival = 0;
    double dval = 0.0;
    int first_id = -1, count = 0;   /* V bound to the packed parameter array */
    void* ptr_storage = nullptr;    /* what the launch's void** points at */
};

int pow2_at_least(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

} // namespace

struct ad4cl_context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
};

struct ad4cl_buffer {
    ad4cl_context* ctx = nullptr;
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;
};

struct ad4cl_kernel {
    ad4cl_context* ctx = nullptr;
    std::string name, signature;
    const void* entries[2][3] = {};  /* ahead-of-time kernel: [recorder 0 uniform / 1 lane][accumulator mode] */
    const void* entry = nullptr;   /* the one selected by prepare() */
    std::vector<const ad4cl_static_row*> static_rows; /* register-tier instances of this built-in kernel */
    bool use_static = false;       /* prepare() selected one of them */
    CUmodule module = nullptr;     /* NVRTC kernel */
    CUfunction functions[2][3] = {};
    /* JIT register tier: a kernel created from SOURCE can be rebuilt as a register-tier instance
       (AD4CL_TAPE_STATIC) for the independents and integer arguments it is bound to; built lazily by
       prepare(), kept while the key (independents, bindings, integer arguments) stays the same */
    std::string source, options;
    CUmodule static_module = nullptr;
    CUfunction static_function = nullptr;
    std::string static_key;        /* what static_module was built for ("" = never tried) */
    bool static_usable = false;    /* ... and whether that build was scalarised (no local memory) */
    int recorder = 1;              /* the recorder (kernel) in use: 0 uniform-first, 1 lane */
    int acc = 0;                   /* the accumulator mode in use */
    CUfunction function = nullptr;
    std::vector<Arg> args;
    ad4cl_kernel_config cfg{};
    bool configured = false;

    /* resources derived from (cfg, nparams); rebuilt when either changes */
    bool ready = false;
    int nparams = -1;
    int block = 0, grid = 0;
    size_t smem = 0, arena_bytes = 0;
    /* Launch shapes the tuner may choose between: [0] the configured (or default 256-thread) work-group,
       [1] one 768-thread work-group per SM, offered when the caller left the work-group size to the
       library (OpenCL's local_work_size == NULL).  Same number of resident warps, same arena. */
    struct Shape { int block = 0, grid = 0, l0 = 0; size_t smem = 0; long long g0 = 0, n_items = 0; };
    Shape shapes[2];
    int nshapes = 1, shape = 0;
    void use_shape(int i) {
        shape = i;
        block = shapes[i].block; grid = shapes[i].grid; smem = shapes[i].smem;
        L.l0 = shapes[i].l0; L.g0 = (int) shapes[i].g0; L.n_items = shapes[i].n_items;
        stats.grid = grid; stats.block = block; stats.smem_bytes = smem;
    }
    int launches_per_eval = 1;
    const double* seeds_dev = nullptr;   /* set around one enqueue by ad4cl_cuda_eval_general */
    double* out_values_dev = nullptr;
    ad4cl::LaunchInfo L{};
    long long data_items = 0;
    struct ad_variable* params_dev = nullptr;
    double* theta_dev = nullptr;                 /* staging for the host path: nparams plain doubles */
    unsigned* ticket_dev = nullptr;              /* last-CTA ticket of the fused second stage */
    bool fused = false;                          /* one launch per evaluation (acc != global, nparams <= kInKernelPackMax) */
    double* partials = nullptr;
    double* result_dev = nullptr;
    double* global_grad = nullptr;
    int* status_dev = nullptr;
    char* arena = nullptr;
    double* theta_pinned = nullptr;              /* the caller's theta, staged for the asynchronous H2D copy */
    double* result_pinned = nullptr;             /* nparams + 4 doubles: [grad, f, ops, slots, status] */
    int* status_pinned = nullptr;
    /* the host evaluation path replayed as one CUDA graph: [0] objective only, [1] with gradient */
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    ad4cl_comm* graph_comm = nullptr;   /* the communicator the captured graphs were built with */
    bool tune_pending = false;          /* AD4CL_RECORDER_AUTO / AD4CL_PREFETCH_AUTO not decided yet */
    bool graph_disabled = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ad4cl_eval_stats stats{};
    /* opt-in timing of the fused kernel alone (ad4cl_cuda_kernel_profile) */
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events; /* pairs */
    size_t prof_next = 0, prof_used = 0;
    double prof_sum_ms = 0.0;
    long long prof_count = 0;

    void drop_graphs() {
        for (auto& g : graph) {
            if (g) cudaGraphExecDestroy(g);
            g = nullptr;
        }
    }

    void release() {
        cudaFree(params_dev); cudaFree(partials); cudaFree(result_dev);
        cudaFree(global_grad); cudaFree(status_dev); cudaFree(arena);
        cudaFree(theta_dev); cudaFree(ticket_dev);
        cudaFreeHost(result_pinned); cudaFreeHost(status_pinned);
        cudaFreeHost(theta_pinned);
        theta_pinned = nullptr; theta_dev = nullptr; ticket_dev = nullptr;
        drop_graphs();
        params_dev = nullptr; partials = nullptr; result_dev = nullptr;
        global_grad = nullptr; status_dev = nullptr; arena = nullptr;
        result_pinned = nullptr; status_pinned = nullptr;
        ready = false;
    }
};

namespace {

int signature_ok(const char* s) {
    if (!s || !*s) return 0;
    int outs = 0;
    for (const char* p = s; *p; p++) {
        if (!strchr("GSOVDIid", *p)) return 0;
        if (*p == 'O') outs++;
    }
    return outs == 1;
}

ad4cl_kernel* new_kernel(ad4cl_context* ctx, const char* name, const char* sig) {
    ad4cl_kernel* k = new ad4cl_kernel();
    k->ctx = ctx;
    k->name = name;
    k->signature = sig;
    k->args.resize(strlen(sig));
    for (size_t i = 0; i < k->args.size(); i++) {
        k->args[i].kind = sig[i];
        if (strchr("GSO", sig[i])) k->args[i].set = true;
    }
    ad4cl_cuda_kernel_config_default(&k->cfg);
    return k;
}

int occupancy(ad4cl_kernel* k, int block, size_t smem, int* blocks) {
    /* the entry's static shared memory (7 KB, plus the user kernel's own __local arrays) counts against the 48 KB
       default as well: opt in early */
    if (!k->module) {
        if (smem > 8 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(k->entry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k->entry, block, smem));
    } else {
        DriverApi& d = driver();
        if (smem > 8 * 1024) {
            CUresult r = d.cuFuncSetAttribute(k->function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int) smem);
            if (r != CUDA_SUCCESS) return fail(AD4CL_ERR_CUDA, "cuFuncSetAttribute(smem=%zu) failed: %d", smem, (int) r);
        }
        CUresult r = d.cuOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k->function, block, smem);
        if (r != CUDA_SUCCESS) return fail(AD4CL_ERR_CUDA, "cuOccupancyMaxActiveBlocksPerMultiprocessor failed: %d", (int) r);
    }
    return AD4CL_OK;
}

struct StaticSpec {
    int nparams = 0;
    const std::vector<Arg>* args = nullptr;
};
/* Which integer arguments a register-tier instance is specialised on: the ones small enough to be the trip
   count of a loop the 64-slot static tape can hold unrolled.  Large ones (the problem size of `id < size`)
   stay run-time parameters -- as literals they buy nothing and were measured to cost registers. */
inline bool static_freezes(int v) { return v >= 0 && v <= 64; }
int compile_to_cubin(const char* source, const char* entry, const char* signature,
        const char* options, int reference_quirks, std::vector<char>* cubin, const StaticSpec* stat = nullptr);

/* JIT register tier: (re)build the register-tier instance of a kernel created from source for exactly these
   independents / bindings / integer arguments, and report whether it can be used (it compiled and the
   compiler scalarised the tape: no local memory).  Never fails the evaluation: a kernel that does not
   qualify simply stays in the arena tiers. */
bool jit_static_instance(ad4cl_kernel* k, int nparams, unsigned cap_slots) {
    if (!k->module || k->source.empty()) return false;
    if (nparams < 1 || nparams > 16 || cap_slots > 64) return false;   /* register accumulators / the 64-slot static tape */
    if (k->functions[0][0] == nullptr) return false;                   /* work-group barriers: not a per-thread kernel */
    std::string key = std::to_string(nparams) + "/" + std::to_string(k->cfg.reference_quirks);
    int next = 0;
    for (const Arg& a : k->args) {
        if (a.kind == 'V') {
            if (!a.set || a.first_id != next) return false;            /* independents bound in argument order */
            next += a.count;
            key += ":V" + std::to_string(a.first_id) + "+" + std::to_string(a.count);
        } else if (a.kind == 'i') {
            if (!a.set) return false;
            key += static_freezes(a.ival) ? ":i" + std::to_string(a.ival) : std::string(":i*");
        }
    }
    if (next != nparams) return false;
    if (key == k->static_key) return k->static_usable;
    if (k->static_module) driver().cuModuleUnload(k->static_module);
    k->static_module = nullptr;
    k->static_function = nullptr;
    k->static_key = key;
    k->static_usable = false;
    StaticSpec spec;
    spec.nparams = nparams;
    spec.args = &k->args;
    std::vector<char> bin;
    if (compile_to_cubin(k->source.c_str(), k->name.c_str(), k->signature.c_str(), k->options.empty() ? nullptr : k->options.c_str(),
            k->cfg.reference_quirks, &bin, &spec) != AD4CL_OK)
        return false;   /* e.g. the kernel indexes its independents in a way the local copies cannot serve */
    if (driver().cuModuleLoadData(&k->static_module, bin.data()) != CUDA_SUCCESS) {
        k->static_module = nullptr;
        return false;
    }
    const std::string fn = k->name + "__entry_s";
    int local_bytes = 1;
    if (driver().cuModuleGetFunction(&k->static_function, k->static_module, fn.c_str()) != CUDA_SUCCESS
            || driver().cuFuncGetAttribute(&local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, k->static_function) != CUDA_SUCCESS
            || local_bytes != 0) {
        driver().cuModuleUnload(k->static_module);
        k->static_module = nullptr;
        k->static_function = nullptr;
        return false;   /* the tape did not become registers: the arena tiers are the faster home */
    }
    k->static_usable = true;
    return true;
}

/* size every device resource for `nparams` independents */
int prepare(ad4cl_kernel* k, int nparams) {
    if (k->ready && k->nparams == nparams) return AD4CL_OK;
    if (!k->configured) return fail(AD4CL_ERR_INVALID, "kernel %s: call ad4cl_cuda_kernel_configure first", k->name.c_str());
    if (nparams < 0) return fail(AD4CL_ERR_INVALID, "nparams < 0");
    k->release();
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    const ad4cl_kernel_config& c = k->cfg;

    long long g0 = c.global_size[0], g1 = c.global_size[1] > 0 ? c.global_size[1] : 1;
    int l0 = c.local_size[0], l1 = c.local_size[1] > 0 ? c.local_size[1] : 1;
    if (g0 <= 0) return fail(AD4CL_ERR_INVALID, "global_size[0] must be positive");
    if (l0 <= 0) {
        l0 = g1 > 1 ? 16 : 256;
        l1 = g1 > 1 ? 16 : 1;
    }
    const int block = l0 * l1;
    if (block > AD4CL_MAX_BLOCK || block % 32 != 0)
        return fail(AD4CL_ERR_INVALID, "work-group size %d must be a multiple of 32 and <= %d", block, AD4CL_MAX_BLOCK);
    /* pad the NDRange to whole work-groups; kernels guard with `id < size` (simple.cl:121) */
    g0 = (g0 + l0 - 1) / l0 * l0;
    g1 = (g1 + l1 - 1) / l1 * l1;
    if (g0 >= (1LL << 31) || g0 * g1 >= (1LL << 40)) return fail(AD4CL_ERR_INVALID, "NDRange too large");
    k->data_items = c.global_size[0] * (c.global_size[1] > 0 ? c.global_size[1] : 1);

    int acc = c.acc_mode;
    if (acc == AD4CL_ACC_AUTO) acc = nparams <= 16 ? AD4CL_ACC_THREAD : (nparams <= 512 ? AD4CL_ACC_WARP : AD4CL_ACC_GLOBAL);
    k->acc = acc;
    k->recorder = c.recorder == AD4CL_RECORDER_UNIFORM ? 0 : 1;   /* AUTO starts with the lane kernel until tuned */
    if (k->module && k->functions[0][acc] == nullptr) k->recorder = 1;   /* a kernel with work-group barriers has no uniform build */
    if (k->module) k->function = k->functions[k->recorder][acc];
    else k->entry = k->entries[k->recorder][acc];
    const int max_ops = std::max(c.max_ops, 1);
    const unsigned cap_slots = (unsigned) (c.max_slots > 0 ? c.max_slots : 2 * max_ops);
    const bool far = c.far_plane != 0;
    /* without the far plane the ring must reach every operand: the distance is counted in SLOTS */
    int window = c.window > 0 ? pow2_at_least(c.window) : (far ? 16 : pow2_at_least((int) cap_slots + 1));
    window = std::max(window, 4); /* the uniform sweep pops a batch of 4 consecutive ring entries */
    int tier = c.tape_tier;
    /* Tier choice.  REGISTER (ad4cl_tape.cuh, AD4CL_TAPE_STATIC): the kernel has an instance compiled for
       exactly this many independents, bound as params(0..P) in argument order, and for the value of the
       integer argument that instance froze; AUTO additionally requires that the instance really was
       scalarised (no local memory).  Otherwise AUTO = the HBM arena, which stays L2 resident for short
       tapes; the shared-memory tier is only ever chosen explicitly. */
    const ad4cl_static_row* srow = nullptr;
    bool jstat = false;
    if (tier == AD4CL_TAPE_AUTO || tier == AD4CL_TAPE_REGISTER) {
        for (const ad4cl_static_row* r : k->static_rows) {
            if (r->nparams != nparams) continue;
            bool ok = (unsigned) r->slots <= cap_slots;   /* a deliberately short tape keeps its overflow report */
            if (g1 != 1) ok = false;                      /* the instances are 1-D kernels */
            if (block > AD4CL_STATIC_MAX_BLOCK) ok = false; /* ... built for work-groups of at most 256 */
            /* AUTO leaves explicit choices of the arena machinery alone */
            if (tier == AD4CL_TAPE_AUTO && (c.window != 0 || c.far_plane == 0
                    || (c.acc_mode != AD4CL_ACC_AUTO && c.acc_mode != AD4CL_ACC_THREAD))) ok = false;
            int next = 0;
            for (const Arg& a : k->args) {
                if (a.kind != 'V') continue;
                if (a.first_id != next) ok = false;
                next += a.count;
            }
            if (next != nparams) ok = false;
            if (ok && r->frozen_arg >= 0 && (r->frozen_arg >= (int) k->args.size() || !k->args[r->frozen_arg].set
                    || k->args[r->frozen_arg].ival != r->frozen_value)) ok = false;
            if (ok && tier == AD4CL_TAPE_AUTO) {
                cudaFuncAttributes fa{};
                if (cudaFuncGetAttributes(&fa, r->entry) != cudaSuccess || fa.localSizeBytes != 0) ok = false;
                cudaGetLastError();
            }
            if (ok) {
                srow = r;
                break;
            }
        }
        /* a kernel created from source: build its register-tier instance now (AUTO under the same conditions
           as for the built-in instances) */
        if (!srow && k->module && g1 == 1 && block <= AD4CL_STATIC_MAX_BLOCK
                && (tier == AD4CL_TAPE_REGISTER || (c.window == 0 && c.far_plane != 0
                        && (c.acc_mode == AD4CL_ACC_AUTO || c.acc_mode == AD4CL_ACC_THREAD))))
            jstat = jit_static_instance(k, nparams, cap_slots);
        if (tier == AD4CL_TAPE_REGISTER && !srow && !jstat)
            return fail(AD4CL_ERR_INVALID, "kernel %s has no register-tier instance for %d independents with these "
                    "arguments (built-in instances: linreg2_g / linreg2_p with 2, regression_* with 10; kernels created "
                    "from source: 1-D, <= 16 independents bound in argument order, max_slots <= 64, no barrier, and a "
                    "tape the compiler can scalarise)", k->name.c_str(), nparams);
        tier = (srow || jstat) ? AD4CL_TAPE_REGISTER : AD4CL_TAPE_HBM;
    }
    const bool stat = srow != nullptr || jstat;
    k->use_static = stat;
    if (stat) acc = AD4CL_ACC_THREAD; /* per-thread REGISTER accumulators */
    if (srow) k->entry = srow->entry;
    if (jstat) k->function = k->static_function;
    const bool fused = stat || (acc != AD4CL_ACC_GLOBAL && nparams <= ad4cl::kInKernelPackMax);

    const size_t smem = stat ? sizeof (double) * (size_t) (block / 32) * (nparams + 3)
            : ad4cl::entry_smem_bytes(block, nparams, window, acc, tier == AD4CL_TAPE_SHARED, cap_slots, far);
    if (smem > k->ctx->prop.sharedMemPerBlockOptin)
        return fail(AD4CL_ERR_INVALID, "kernel %s needs %zu B of shared memory per block (limit %zu): "
                "lower window/local size or use AD4CL_ACC_WARP", k->name.c_str(), smem,
                (size_t) k->ctx->prop.sharedMemPerBlockOptin);
    int per_sm = 0;
    int rc = occupancy(k, block, smem, &per_sm);
    if (rc != AD4CL_OK) return rc;
    if (!stat) { /* the other recorder's kernel may be launched too (AUTO times both): same opt-in, the smaller occupancy */
        const int other = 1 - k->recorder;
        if (k->module ? k->functions[other][acc] != nullptr : k->entries[other][acc] != nullptr) {
            const void* e0 = k->entry;
            CUfunction f0 = k->function;
            if (k->module) k->function = k->functions[other][acc];
            else k->entry = k->entries[other][acc];
            int per_sm2 = 0;
            rc = occupancy(k, block, smem, &per_sm2);
            k->entry = e0;
            k->function = f0;
            if (rc != AD4CL_OK) return rc;
            per_sm = std::min(per_sm, per_sm2);
        }
    }
    if (per_sm < 1) return fail(AD4CL_ERR_INVALID, "kernel %s does not fit on an SM (block %d, smem %zu)", k->name.c_str(), block, smem);
    if (c.blocks_per_sm > 0) per_sm = std::min(per_sm, c.blocks_per_sm);
    const long long ngroups = (g0 / l0) * (g1 / l1);
    int grid = (int) std::min<long long>((long long) k->ctx->prop.multiProcessorCount * per_sm, ngroups);
    grid = std::max(grid, 1);
    k->shapes[0] = {block, grid, l0, smem, g0, g0 * g1};
    k->nshapes = 1;

    /* The alternative shape: ONE work-group of AD4CL_MAX_BLOCK threads per SM.  With the work-groups in
       lock step (LaunchInfo::lockstep) all resident warps then walk the same stretch of a long objective
       together, which is what a body larger than the instruction caches wants (profiles/README.md).
       Only when the caller left the work-group size open, for 1-D arena kernels. */
    if (!stat && tier == AD4CL_TAPE_HBM && c.local_size[0] <= 0 && g1 == 1 && c.blocks_per_sm <= 0
            && AD4CL_MAX_BLOCK > block && c.global_size[0] >= (long long) AD4CL_MAX_BLOCK * k->ctx->prop.multiProcessorCount) {
        const int b2 = AD4CL_MAX_BLOCK;
        const size_t smem2 = ad4cl::entry_smem_bytes(b2, nparams, window, acc, false, cap_slots, far);
        if (smem2 <= k->ctx->prop.sharedMemPerBlockOptin) {
            int ps = 0, ps2 = 0;
            rc = occupancy(k, b2, smem2, &ps);
            if (rc != AD4CL_OK) return rc;
            const int other = 1 - k->recorder;
            if (k->module ? k->functions[other][acc] != nullptr : k->entries[other][acc] != nullptr) {
                const void* e0 = k->entry;
                CUfunction f0 = k->function;
                if (k->module) k->function = k->functions[other][acc];
                else k->entry = k->entries[other][acc];
                rc = occupancy(k, b2, smem2, &ps2);
                k->entry = e0;
                k->function = f0;
                if (rc != AD4CL_OK) return rc;
                ps = std::min(ps, ps2);
            }
            const long long g0b = (c.global_size[0] + b2 - 1) / b2 * b2;
            if (ps >= 1 && g0b < (1LL << 31) && ps * b2 >= per_sm * block) { /* never fewer resident warps */
                const int grid2 = (int) std::min<long long>((long long) k->ctx->prop.multiProcessorCount * ps, g0b / b2);
                k->shapes[1] = {b2, grid2, b2, smem2, g0b, g0b};
                k->nshapes = 2;
            }
        }
    }
    int max_grid = grid, max_warps = grid * (block / 32);
    if (k->nshapes > 1) {
        max_grid = std::max(max_grid, k->shapes[1].grid);
        max_warps = std::max(max_warps, k->shapes[1].grid * (k->shapes[1].block / 32));
    }

    const size_t region = ad4cl::warp_region_bytes(cap_slots, far);
    const size_t arena_bytes = tier != AD4CL_TAPE_HBM ? 0 : (size_t) max_warps * region;
    const int ncols = nparams + 3;
    const int pcols = (acc == AD4CL_ACC_GLOBAL ? 0 : nparams) + 3; /* columns of the per-block partials */

    CUDA_TRY(cudaMalloc(&k->params_dev, sizeof (struct ad_variable) * (size_t) std::max(nparams, 1)));
    CUDA_TRY(cudaMalloc(&k->partials, sizeof (double) * (size_t) max_grid * pcols));
    CUDA_TRY(cudaMalloc(&k->result_dev, sizeof (double) * (size_t) (ncols + 1)));
    CUDA_TRY(cudaMalloc(&k->status_dev, sizeof (int)));
    CUDA_TRY(cudaMemset(k->status_dev, 0, sizeof (int)));
    if (acc == AD4CL_ACC_GLOBAL) CUDA_TRY(cudaMalloc(&k->global_grad, sizeof (double) * (size_t) std::max(nparams, 1)));
    if (arena_bytes) {
        cudaError_t e = cudaMalloc(&k->arena, arena_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(AD4CL_ERR_NOMEM, "tape arena of %zu bytes (%d warps x %zu) does not fit: lower blocks_per_sm or max_ops",
                    arena_bytes, max_warps, region);
        }
        CUDA_TRY(cudaMemset(k->arena, 0, arena_bytes)); /* far bitmaps must start clear */
    }
    CUDA_TRY(cudaMallocHost(&k->result_pinned, sizeof (double) * (size_t) (ncols + 1)));
    CUDA_TRY(cudaMallocHost(&k->theta_pinned, sizeof (double) * (size_t) std::max(nparams, 1)));
    CUDA_TRY(cudaMalloc(&k->theta_dev, sizeof (double) * (size_t) std::max(nparams, 1)));
    CUDA_TRY(cudaMalloc(&k->ticket_dev, sizeof (unsigned)));
    CUDA_TRY(cudaMemset(k->ticket_dev, 0, sizeof (unsigned)));
    CUDA_TRY(cudaMallocHost(&k->status_pinned, sizeof (int)));
    if (!k->ev0) CUDA_TRY(cudaEventCreate(&k->ev0));
    if (!k->ev1) CUDA_TRY(cudaEventCreate(&k->ev1));

    ad4cl::LaunchInfo& L = k->L;
    L.g0_user = (int) c.global_size[0];
    L.n_user = c.global_size[0] * (c.global_size[1] > 0 ? c.global_size[1] : 1);
    L.nparams = nparams;
    L.lockstep = c.lockstep == AD4CL_LOCKSTEP_OFF ? 0 : 1;
    L.recording = 1;
    L.acc_mode = acc;
    L.window = window;
    L.use_far = far ? 1 : 0;
    L.tape_in_smem = tier == AD4CL_TAPE_SHARED ? 1 : 0;
    L.sweep_prefetch = c.sweep_prefetch == AD4CL_PREFETCH_ON ? 1 : 0;
    /* AUTO: decided by autotune_recorder on the first gradient evaluation (the register tier has no choice to make) */
    k->tune_pending = !stat && (c.recorder == AD4CL_RECORDER_AUTO || c.sweep_prefetch == AD4CL_PREFETCH_AUTO
                                || c.lockstep == AD4CL_LOCKSTEP_AUTO || k->nshapes > 1);
    k->stats.recorder = k->tune_pending && c.recorder == AD4CL_RECORDER_AUTO ? -1 : (stat ? -1 : k->recorder);
    k->stats.sweep_prefetch = k->tune_pending && c.sweep_prefetch == AD4CL_PREFETCH_AUTO ? -1 : L.sweep_prefetch;
    L.cap_slots = cap_slots;
    L.arena = k->arena;
    L.warp_region = region;
    L.partials = k->partials;
    L.global_grad = k->global_grad;
    L.status = k->status_dev;
    L.params = k->params_dev;
    L.comm.world = 1;
    k->fused = fused;

    k->arena_bytes = arena_bytes;
    k->nparams = nparams;
    k->use_shape(0);
    k->stats.lockstep = L.lockstep;
    k->stats.arena_bytes = arena_bytes;
    k->stats.tape_tier = tier;
    k->ready = true;
    return AD4CL_OK;
}

/* fold the recorded event pairs into the running mean (synchronises the stream) */
int drain_profile(ad4cl_kernel* k) {
    if (k->prof_used == 0) return AD4CL_OK;
    CUDA_TRY(cudaStreamSynchronize(k->ctx->stream));
    const size_t pairs = k->prof_events.size() / 2;
    size_t idx = (k->prof_next + pairs - k->prof_used) % pairs;
    for (size_t i = 0; i < k->prof_used; i++) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, k->prof_events[2 * idx], k->prof_events[2 * idx + 1]));
        k->prof_sum_ms += ms;
        k->prof_count++;
        idx = (idx + 1) % pairs;
    }
    k->prof_used = 0;
    return AD4CL_OK;
}

/*
 * Enqueue one evaluation on the context's stream.  theta_dev: nparams plain doubles in device memory.
 * Normal case (k->fused): ONE launch -- the entry converts theta to ad_variables, records, sweeps,
 * reduces per block, and the CTA that finishes last runs the second stage, the all-reduce over the
 * ranks of `comm` (if any) and the epilogue.  With fp64-atomic accumulation (matrix configs: 2 n^2
 * independents) the parameters are packed and the partial rows reduced by separate small launches.
 */
int enqueue(ad4cl_kernel* k, const double* theta_dev, double* result_dev, int want_gradient, int apply_epilogue,
        double* status_out = nullptr, ad4cl_comm* comm = nullptr, const int* skip_dev = nullptr);

int status_to_error(ad4cl_kernel* k, int st);

} // namespace

struct ad4cl_comm {
    ad4cl_context* ctx = nullptr;
    int rank = 0, world = 1, stride = 0;
    double* inbox = nullptr;           /* [2][world][stride] */
    ad4cl::CommInfo info{};            /* peers' inboxes as mapped into this process, epoch counter, timeout */
    bool connected = false;
    bool local_group = false;          /* created by ad4cl_cuda_comm_create_local: peers are plain pointers, not IPC mappings */
    int* status_dev = nullptr;
    int* status_pinned = nullptr;
};

namespace {

int enqueue(ad4cl_kernel* k, const double* theta_dev, double* result_dev, int want_gradient, int apply_epilogue,
        double* status_out, ad4cl_comm* comm, const int* skip_dev) {
    cudaStream_t st = k->ctx->stream;
    const int P = k->nparams;
    if (comm != nullptr && comm->world > 1 && !k->fused)
        return fail(AD4CL_ERR_INVALID, "kernel %s: the sharded evaluation needs the fused second stage "
                "(accumulator mode thread or warp, at most %d independents)", k->name.c_str(), ad4cl::kInKernelPackMax);
    ad4cl::LaunchInfo L = k->L;
    L.recording = want_gradient ? 1 : 0;
    L.skip = skip_dev;
    L.seeds = k->seeds_dev;
    L.out_values = k->out_values_dev;
    const double n_obs = k->cfg.epilogue_n > 0 ? k->cfg.epilogue_n : (double) k->data_items;
    const int epi = apply_epilogue ? k->cfg.epilogue : 0;
    if (k->fused) {
        L.theta = P > 0 ? theta_dev : nullptr;
        L.ticket = k->ticket_dev;
        L.result = result_dev;
        L.status_out = status_out;
        L.epilogue = epi;
        L.n_obs = n_obs;
        if (comm != nullptr && comm->world > 1) L.comm = comm->info;
    } else {
        if (P > 0 && theta_dev != nullptr)
            pack_params<<<(P + 255) / 256, 256, 0, st>>>(theta_dev, k->params_dev, P);
        if (k->global_grad)
            CUDA_TRY(cudaMemsetAsync(k->global_grad, 0, sizeof (double) * (size_t) std::max(P, 1), st));
    }

    /* marshal: LaunchInfo first, then the user kernel's non-runtime parameters in order */
    std::vector<void*> argv;
    argv.push_back(&L);
    for (size_t i = 0; i < k->args.size(); i++) {
        Arg& a = k->args[i];
        if (strchr("GSO", a.kind)) continue;
        if (!a.set) return fail(AD4CL_ERR_INVALID, "kernel %s: argument %zu ('%c') is not set", k->name.c_str(), i, a.kind);
        switch (a.kind) {
            case 'V':
                if (a.first_id >= 0) {
                    if (a.first_id + a.count > P)
                        return fail(AD4CL_ERR_INVALID, "kernel %s: argument %zu binds independents [%d,%d) but nparams = %d",
                                k->name.c_str(), i, a.first_id, a.first_id + a.count, P);
                    a.ptr_storage = k->params_dev + a.first_id;
                } else {
                    a.ptr_storage = a.buf->ptr;
                }
                argv.push_back(&a.ptr_storage);
                break;
            case 'D':
            case 'I':
                a.ptr_storage = a.buf ? a.buf->ptr : nullptr;
                argv.push_back(&a.ptr_storage);
                break;
            case 'i':
                argv.push_back(&a.ival);
                break;
            case 'd':
                argv.push_back(&a.dval);
                break;
        }
    }
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (k->profiling) {
        if (k->prof_used == k->prof_events.size() / 2) {
            int rc = drain_profile(k);
            if (rc != AD4CL_OK) return rc;
        }
        pe0 = k->prof_events[2 * k->prof_next];
        pe1 = k->prof_events[2 * k->prof_next + 1];
        k->prof_next = (k->prof_next + 1) % (k->prof_events.size() / 2);
        k->prof_used++;
        CUDA_TRY(cudaEventRecord(pe0, st));
    }
    if (!k->module) {
        CUDA_TRY(cudaLaunchKernel(k->entry, dim3(k->grid), dim3(k->block), argv.data(), k->smem, st));
    } else {
        CUresult r = driver().cuLaunchKernel(k->function, k->grid, 1, 1, k->block, 1, 1, (unsigned) k->smem,
                (CUstream) st, argv.data(), nullptr);
        if (r != CUDA_SUCCESS) return fail(AD4CL_ERR_CUDA, "cuLaunchKernel(%s) failed: %d", k->name.c_str(), (int) r);
    }
    if (pe1) CUDA_TRY(cudaEventRecord(pe1, st));
    k->launches_per_eval = 1;
    if (!k->fused) {
        if (k->global_grad) {
            ad4cl_reduce_partials<<<3, 256, 0, st>>>(k->partials, k->grid, 3, 0, P, epi, n_obs, result_dev,
                    k->status_dev, status_out);
            if (P > 0)
                ad4cl_finish_global_grad<<<(P + 255) / 256, 256, 0, st>>>(k->global_grad, P, k->partials, k->grid,
                        epi, n_obs, result_dev);
            k->launches_per_eval = 2 + (P > 0 ? 2 : 0);
        } else {
            ad4cl_reduce_partials<<<P + 3, 256, 0, st>>>(k->partials, k->grid, P + 3, P, 0, epi, n_obs, result_dev,
                    k->status_dev, status_out);
            k->launches_per_eval = 2 + (P > 0 ? 1 : 0);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return AD4CL_OK;
}

/*
 * AD4CL_RECORDER_AUTO / AD4CL_PREFETCH_AUTO: time one evaluation (CUDA events on the context's stream) for
 * each admissible combination of recorder and sweep prefetch and keep the fastest for this configuration.
 * Runs once, on the first gradient evaluation, never inside a stream capture (a sharded evaluation tunes
 * with local evaluations: the peers take no part).  The first run also warms up; a combination that takes
 * longer than 200 ms is not run a second time.  The choice changes no result beyond the order in which an
 * independent's contributions are summed over a warp.
 */
int autotune_recorder(ad4cl_kernel* k, const double* theta_dev) {
    cudaStream_t st = k->ctx->stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return AD4CL_OK; /* a caller-owned stream under graph capture cannot be timed: decide on a later call */
    }
    k->tune_pending = false;
    struct Cand { int recorder, prefetch, shape, lockstep; float ms; };
    std::vector<Cand> cands;
    const ad4cl_kernel_config& c = k->cfg;
    const int ls0 = c.lockstep == AD4CL_LOCKSTEP_OFF ? 0 : 1;
    for (int rec = 0; rec < 2; rec++) {
        if (c.recorder != AD4CL_RECORDER_AUTO && (c.recorder == AD4CL_RECORDER_LANE) != (rec == 1)) continue;
        if (rec == 0 && k->module && k->functions[0][k->acc] == nullptr) continue;
        for (int pf = 0; pf < 2; pf++) {
            if (rec == 1 && c.sweep_prefetch != AD4CL_PREFETCH_AUTO && (c.sweep_prefetch == AD4CL_PREFETCH_ON) != (pf == 1)) continue;
            if (k->L.tape_in_smem && pf == 1) continue;
            for (int sh = 0; sh < k->nshapes; sh++) {
                if (sh == 1 && (rec == 1 || !ls0 || pf == 1)) continue; /* the one-group-per-SM shape pays with the uniform recorder in lock step only (measured); it is for tapes that stay in L2, where a prefetch buys nothing */
                cands.push_back({rec, pf, sh, ls0, 1e30f});
            }
        }
    }
    if (cands.empty())   /* e.g. the uniform recorder was asked for a kernel with work-group barriers: keep what prepare() chose */
        cands.push_back({k->recorder, k->L.sweep_prefetch, 0, ls0, 1e30f});
    auto apply = [&](const Cand& cd) {
        k->recorder = cd.recorder;
        if (k->module) k->function = k->functions[k->recorder][k->acc];
        else k->entry = k->entries[k->recorder][k->acc];
        k->L.sweep_prefetch = cd.prefetch;
        k->L.lockstep = cd.lockstep;
        k->use_shape(cd.shape);
    };
    auto time_all = [&](std::vector<Cand>& cs, size_t from) -> int {
        for (int round = 0; round < 2; round++) {
            for (size_t i = from; i < cs.size(); i++) {
                Cand& cd = cs[i];
                if (round == 1 && cd.ms > 200.f && cd.ms < 1e29f) continue;
                apply(cd);
                CUDA_TRY(cudaEventRecord(k->ev0, st));
                int rc = enqueue(k, theta_dev, k->result_dev, 1, 0);
                if (rc != AD4CL_OK) return rc;
                CUDA_TRY(cudaEventRecord(k->ev1, st));
                CUDA_TRY(cudaEventSynchronize(k->ev1));
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, k->ev0, k->ev1));
                if (round == 1 || ms > 200.f) cd.ms = std::min(cd.ms, ms);
            }
        }
        return AD4CL_OK;
    };
    auto best_of = [&]() {
        size_t b = 0;
        for (size_t i = 1; i < cands.size(); i++)
            if (cands[i].ms < cands[b].ms) b = i;
        return b;
    };
    if (cands.size() > 1 || c.lockstep == AD4CL_LOCKSTEP_AUTO) {
        int rc = time_all(cands, 0);
        if (rc != AD4CL_OK) return rc;
        if (c.lockstep == AD4CL_LOCKSTEP_AUTO && cands[best_of()].shape == 0) {
            /* second stage: the winner without the per-work-group barrier (work-items of very different
               lengths in one work-group, or a body that fits the instruction cache anyway) */
            Cand free_running = cands[best_of()];
            free_running.lockstep = 0;
            free_running.ms = 1e30f;
            cands.push_back(free_running);
            rc = time_all(cands, cands.size() - 1);
            if (rc != AD4CL_OK) return rc;
        }
    }
    const Cand best = cands[best_of()];
    apply(best);
    k->stats.recorder = k->recorder;
    k->stats.sweep_prefetch = k->L.sweep_prefetch;
    k->stats.lockstep = k->L.lockstep;
    CUDA_TRY(cudaMemsetAsync(k->status_dev, 0, sizeof (int), st)); /* an error shows up again in the evaluation proper */
    return AD4CL_OK;
}

int status_to_error(ad4cl_kernel* k, int st) {
    if (st == 0) return AD4CL_OK;
    /* leave the arena clean for the next evaluation (far bitmaps may be dirty) */
    if (k->arena) cudaMemsetAsync(k->arena, 0, k->arena_bytes, k->ctx->stream);
    cudaMemsetAsync(k->status_dev, 0, sizeof (int), k->ctx->stream);
    cudaStreamSynchronize(k->ctx->stream);
    if (st & ad4cl::kErrBounds)
        return fail(AD4CL_ERR_CUDA, "kernel %s: internal bounds check failed (AD4CL_DEBUG_BOUNDS build)", k->name.c_str());
    if (st & ad4cl::kErrTapeOverflow)
        return fail(AD4CL_ERR_TAPE_OVERFLOW, "kernel %s: a work-item recorded more than max_ops=%d operators / %u slots "
                "(the reference overflows its stack silently here, SURVEY Q8)", k->name.c_str(), k->cfg.max_ops, k->L.cap_slots);
    return fail(AD4CL_ERR_WINDOW, "kernel %s: an operand older than window=%d operators was used and far_plane is off",
            k->name.c_str(), k->L.window);
}

} // namespace

/*
 * Stand-alone form of the one-shot all-reduce (ad4cl_cuda_comm_allreduce): the same CTA routine the
 * evaluation kernel's last CTA runs (ad4cl_entry.cuh), in place on a device vector, epilogue fused.
 */
namespace {

__global__ void __launch_bounds__(256) oneshot_allreduce_kernel(ad4cl::CommInfo C, int n, double* __restrict__ data,
        int nparams, int epilogue, double n_obs, int* __restrict__ status) {
    __shared__ unsigned long long epoch_s;
    __shared__ double S_s;
    if (threadIdx.x == 0) epoch_s = ++(*C.epoch);
    __syncthreads();
    const bool ok = ad4cl::oneshot_allreduce_cta(C, data, n, epoch_s);
    if (!ok && threadIdx.x == 0) atomicOr(status, 4);
    if (epilogue == 1 && ok) {
        if (threadIdx.x == 0) S_s = data[nparams];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) data[i] = ad4cl::apply_epilogue(data[i], i, nparams, S_s, n_obs);
    }
}

} // namespace

/* ============================================================ C ABI */
#pragma GCC visibility push(default)
extern "C" {

const char* ad4cl_cuda_last_error(void) { return g_last_error.c_str(); }

/* for the library's other translation units (dense_matrix.cu): set the last-error text */
int ad4cl_cuda_set_error_(int code, const char* msg) { return fail(code, "%s", msg); }

const char* ad4cl_cuda_device_abi(void) { return ad4cl_embedded_header_hash; }

int ad4cl_cuda_init(int device, ad4cl_context** out) {
    if (!out) return fail(AD4CL_ERR_INVALID, "ctx is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(AD4CL_ERR_CUDA, "no CUDA device (%s): libad4cl_cuda has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(AD4CL_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    ad4cl_context* c = new ad4cl_context();
    c->device = device;
    CUDA_TRY(cudaGetDeviceProperties(&c->prop, device));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    *out = c;
    return AD4CL_OK;
}

int ad4cl_cuda_shutdown(ad4cl_context* ctx) {
    if (!ctx) return AD4CL_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return AD4CL_OK;
}

void* ad4cl_cuda_stream(ad4cl_context* ctx) { return ctx ? (void*) ctx->stream : nullptr; }

int ad4cl_cuda_use_stream(ad4cl_context* ctx, void* cuda_stream) {
    if (!ctx) return fail(AD4CL_ERR_INVALID, "ctx is NULL");
    ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
    return AD4CL_OK;
}

int ad4cl_cuda_device_sm_count(ad4cl_context* ctx) { return ctx ? ctx->prop.multiProcessorCount : 0; }

int ad4cl_cuda_buffer_create(ad4cl_context* ctx, size_t bytes, ad4cl_buffer** out) {
    if (!ctx || !out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    ad4cl_buffer* b = new ad4cl_buffer();
    b->ctx = ctx;
    b->bytes = bytes;
    cudaError_t e = cudaMalloc(&b->ptr, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
        delete b;
        cudaGetLastError();
        return fail(AD4CL_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    *out = b;
    return AD4CL_OK;
}

int ad4cl_cuda_buffer_wrap(ad4cl_context* ctx, void* device_ptr, size_t bytes, ad4cl_buffer** out) {
    if (!ctx || !out || !device_ptr) return fail(AD4CL_ERR_INVALID, "NULL argument");
    ad4cl_buffer* b = new ad4cl_buffer();
    b->ctx = ctx;
    b->ptr = device_ptr;
    b->bytes = bytes;
    b->owned = false;
    *out = b;
    return AD4CL_OK;
}

int ad4cl_cuda_buffer_upload(ad4cl_buffer* buf, size_t offset, const void* host, size_t bytes) {
    if (!buf || (!host && bytes)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (offset + bytes > buf->bytes) return fail(AD4CL_ERR_INVALID, "upload of %zu bytes at %zu exceeds buffer of %zu", bytes, offset, buf->bytes);
    CUDA_TRY(cudaSetDevice(buf->ctx->device));
    CUDA_TRY(cudaMemcpyAsync((char*) buf->ptr + offset, host, bytes, cudaMemcpyHostToDevice, buf->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(buf->ctx->stream)); /* blocking, like enqueueWriteBuffer(CL_TRUE) */
    return AD4CL_OK;
}

int ad4cl_cuda_buffer_download(ad4cl_buffer* buf, size_t offset, void* host, size_t bytes) {
    if (!buf || (!host && bytes)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (offset + bytes > buf->bytes) return fail(AD4CL_ERR_INVALID, "download of %zu bytes at %zu exceeds buffer of %zu", bytes, offset, buf->bytes);
    CUDA_TRY(cudaSetDevice(buf->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(host, (char*) buf->ptr + offset, bytes, cudaMemcpyDeviceToHost, buf->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(buf->ctx->stream));
    return AD4CL_OK;
}

void* ad4cl_cuda_buffer_device_ptr(ad4cl_buffer* buf) { return buf ? buf->ptr : nullptr; }

int ad4cl_cuda_buffer_free(ad4cl_buffer* buf) {
    if (!buf) return AD4CL_OK;
    if (buf->owned) {
        cudaSetDevice(buf->ctx->device);
        cudaFree(buf->ptr);
    }
    delete buf;
    return AD4CL_OK;
}

int ad4cl_cuda_builtin_count(void) {
    int n = 0;
    while (ad4cl_builtin_table_fixed[n].name) n++;
    return n;
}
const char* ad4cl_cuda_builtin_name(int i) {
    return (i >= 0 && i < ad4cl_cuda_builtin_count()) ? ad4cl_builtin_table_fixed[i].name : nullptr;
}
const char* ad4cl_cuda_builtin_signature(int i) {
    return (i >= 0 && i < ad4cl_cuda_builtin_count()) ? ad4cl_builtin_table_fixed[i].signature : nullptr;
}

int ad4cl_cuda_kernel_builtin(ad4cl_context* ctx, const char* name, int reference_quirks, ad4cl_kernel** out) {
    if (!ctx || !name || !out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    const ad4cl_builtin_row* tab = reference_quirks ? ad4cl_builtin_table_quirks : ad4cl_builtin_table_fixed;
    for (const ad4cl_builtin_row* r = tab; r->name; r++) {
        if (strcmp(r->name, name) == 0) {
            ad4cl_kernel* k = new_kernel(ctx, name, r->signature);
            for (int rc2 = 0; rc2 < 2; rc2++)
                for (int m = 0; m < 3; m++) k->entries[rc2][m] = r->entry[rc2][m];
            const ad4cl_static_row* const stabs[2][2] = {{ad4cl_static_table_fixed_p2, ad4cl_static_table_fixed_p10},
                {ad4cl_static_table_quirks_p2, ad4cl_static_table_quirks_p10}};
            for (const ad4cl_static_row* st : stabs[reference_quirks ? 1 : 0])
                for (; st->name; st++)
                    if (strcmp(st->name, name) == 0) k->static_rows.push_back(st);
            k->cfg.reference_quirks = reference_quirks;
            *out = k;
            return AD4CL_OK;
        }
    }
    return fail(AD4CL_ERR_INVALID, "no built-in kernel named '%s'", name);
}

extern "C++" {
namespace {
/* OpenCL prelude + user text + generated entry -> sm_100a cubin (no device needed).  With `stat` the
   generated entry is the register-tier one: the independents are local copies `th` with literal ids and the
   integer arguments are literals (they bound the kernel's loops, which must unroll for the tape to scalarise). */
int compile_to_cubin(const char* source, const char* entry, const char* signature,
        const char* options, int reference_quirks, std::vector<char>* cubin, const StaticSpec* stat) {
    if (!source || !entry) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (!signature_ok(signature)) return fail(AD4CL_ERR_INVALID, "bad signature '%s' (letters GSOVDIid, exactly one O)", signature ? signature : "");
    NvrtcApi& n = nvrtc();
    if (!n.ok) return fail(AD4CL_ERR_COMPILE, "libnvrtc could not be loaded: %s", dlerror());

    std::string params, call;
    for (size_t i = 0; signature[i]; i++) {
        char nm[32];
        snprintf(nm, sizeof nm, "a%zu", i);
        const char c = signature[i];
        std::string passed = c == 'G' ? "gs" : c == 'S' ? "gradient_stack" : c == 'O' ? "out" : nm;
        if (stat && c == 'V') passed = "th + " + std::to_string((*stat->args)[i].first_id);
        if (stat && c == 'i' && static_freezes((*stat->args)[i].ival)) passed = "(" + std::to_string((*stat->args)[i].ival) + ")";
        if (!call.empty()) call += ", ";
        call += passed;
        const char* type = c == 'V' ? "struct ad_variable* " : c == 'D' ? "double* " : c == 'I' ? "int* "
                : c == 'i' ? "int " : c == 'd' ? "double " : nullptr;
        if (type) {
            if (!params.empty()) params += ", ";
            params += type;
            params += nm;
        }
    }
    std::string tu = "#include \"ad4cl_opencl_compat.cuh\"\n#line 1 \"user_kernel.cl\"\n";
    tu += source;
    tu += stat ? "\n#line 1 \"ad4cl_generated_entry\"\nAD4CL_STATIC_ENTRY(" : "\n#line 1 \"ad4cl_generated_entry\"\nAD4CL_OBJECTIVE_ENTRY(";
    tu += entry;
    tu += ", (" + params + "), (" + call + "))\n";

    nvrtcProgram prog = nullptr;
    int rc = n.CreateProgram(&prog, tu.c_str(), "ad4cl_user.cu", ad4cl_embedded_header_count,
            ad4cl_embedded_header_texts, ad4cl_embedded_header_names);
    if (rc != 0) return fail(AD4CL_ERR_COMPILE, "nvrtcCreateProgram failed: %d", rc);
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device",
        std::string("-DAD4CL_REFERENCE_QUIRKS=") + (reference_quirks ? "1" : "0")};
    if (stat) {
        opts.push_back("-DAD4CL_TAPE_STATIC=1");
        opts.push_back("-DAD4CL_P=" + std::to_string(stat->nparams));
    }
    /* a kernel that synchronises its work-group cannot be re-run by a single warp after a failed optimistic
       recording attempt (ad4cl_entry.cuh): build it with the general recorder only */
    if (strstr(source, "barrier") != nullptr || strstr(source, "__syncthreads") != nullptr || strstr(source, "__local") != nullptr
            || strstr(source, "__shared__") != nullptr)
        opts.push_back("-DAD4CL_BODY_HAS_BARRIER=1");
    if (options) {
        std::string o(options);
        size_t pos = 0;
        while (pos < o.size()) {
            size_t sp = o.find(' ', pos);
            if (sp == std::string::npos) sp = o.size();
            if (sp > pos) opts.push_back(o.substr(pos, sp - pos));
            pos = sp + 1;
        }
    }
    std::vector<const char*> optv;
    for (auto& o : opts) optv.push_back(o.c_str());
    rc = n.CompileProgram(prog, (int) optv.size(), optv.data());
    if (rc != 0) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        n.GetProgramLog(prog, &log[0]);
        n.DestroyProgram(&prog);
        return fail(AD4CL_ERR_COMPILE, "NVRTC build of '%s' failed:\n%s", entry, log.c_str());
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    cubin->resize(cs);
    n.GetCUBIN(prog, cubin->data());
    n.DestroyProgram(&prog);
    return AD4CL_OK;
}
} // namespace
} // extern "C++"

int ad4cl_cuda_compile(const char* source, const char* entry, const char* signature, const char* options,
        int reference_quirks, void** cubin, size_t* cubin_bytes) {
    if (!cubin || !cubin_bytes) return fail(AD4CL_ERR_INVALID, "NULL argument");
    std::vector<char> bin;
    int rc = compile_to_cubin(source, entry, signature, options, reference_quirks, &bin);
    if (rc != AD4CL_OK) return rc;
    *cubin = malloc(bin.size());
    if (!*cubin) return fail(AD4CL_ERR_NOMEM, "malloc(%zu)", bin.size());
    memcpy(*cubin, bin.data(), bin.size());
    *cubin_bytes = bin.size();
    return AD4CL_OK;
}

void ad4cl_cuda_free_host(void* p) { free(p); }

int ad4cl_cuda_kernel_from_cubin(ad4cl_context* ctx, const void* cubin, size_t cubin_bytes, const char* entry,
        const char* signature, int reference_quirks, ad4cl_kernel** out) {
    if (!ctx || !cubin || !cubin_bytes || !entry || !out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (!signature_ok(signature)) return fail(AD4CL_ERR_INVALID, "bad signature '%s'", signature ? signature : "");
    if (!driver().ok) return fail(AD4CL_ERR_CUDA, "libcuda.so.1 could not be loaded");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaFree(0)); /* make the primary context current for the driver API */
    ad4cl_kernel* k = new_kernel(ctx, entry, signature);
    CUresult r = driver().cuModuleLoadData(&k->module, cubin);
    if (r != CUDA_SUCCESS) {
        delete k;
        return fail(AD4CL_ERR_CUDA, "cuModuleLoadData failed: %d", (int) r);
    }
    for (int rec = 0; rec < 2; rec++) {
        for (int m = 0; m < 3; m++) {
            std::string fn = std::string(entry) + (rec == 0 ? "__entry_u" : "__entry_m") + std::to_string(m);
            r = driver().cuModuleGetFunction(&k->functions[rec][m], k->module, fn.c_str());
            if (r != CUDA_SUCCESS) {
                k->functions[rec][m] = nullptr;
                if (rec == 0) continue; /* no uniform build: a kernel with work-group barriers */
                driver().cuModuleUnload(k->module);
                delete k;
                return fail(AD4CL_ERR_CUDA, "cuModuleGetFunction(%s) failed: %d", fn.c_str(), (int) r);
            }
        }
    }
    k->cfg.reference_quirks = reference_quirks;
    *out = k;
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_from_source(ad4cl_context* ctx, const char* source, const char* entry,
        const char* signature, const char* options, int reference_quirks, ad4cl_kernel** out) {
    if (!ctx || !out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    std::vector<char> bin;
    int rc = compile_to_cubin(source, entry, signature, options, reference_quirks, &bin);
    if (rc != AD4CL_OK) return rc;
    rc = ad4cl_cuda_kernel_from_cubin(ctx, bin.data(), bin.size(), entry, signature, reference_quirks, out);
    if (rc == AD4CL_OK) { /* kept for the register-tier instance prepare() may build (jit_static_instance) */
        (*out)->source = source;
        (*out)->options = options ? options : "";
    }
    return rc;
}

int ad4cl_cuda_kernel_free(ad4cl_kernel* k) {
    if (!k) return AD4CL_OK;
    cudaSetDevice(k->ctx->device);
    cudaStreamSynchronize(k->ctx->stream);
    k->release();
    if (k->ev0) cudaEventDestroy(k->ev0);
    if (k->ev1) cudaEventDestroy(k->ev1);
    for (auto& e : k->prof_events) cudaEventDestroy(e);
    if (k->static_module && driver().cuModuleUnload) driver().cuModuleUnload(k->static_module);
    if (k->module && driver().cuModuleUnload) driver().cuModuleUnload(k->module);
    delete k;
    return AD4CL_OK;
}

const char* ad4cl_cuda_kernel_signature(ad4cl_kernel* k) { return k ? k->signature.c_str() : nullptr; }

static int arg_at(ad4cl_kernel* k, int index, const char* kinds, Arg** a) {
    if (!k) return fail(AD4CL_ERR_INVALID, "kernel is NULL");
    if (index < 0 || index >= (int) k->args.size())
        return fail(AD4CL_ERR_INVALID, "kernel %s has %zu parameters; index %d is out of range", k->name.c_str(), k->args.size(), index);
    *a = &k->args[index];
    if (strchr("GSO", (*a)->kind)) return 1; /* runtime-provided: accept and ignore */
    if (!strchr(kinds, (*a)->kind))
        return fail(AD4CL_ERR_INVALID, "kernel %s: parameter %d has kind '%c'", k->name.c_str(), index, (*a)->kind);
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_set_arg_buffer(ad4cl_kernel* k, int index, ad4cl_buffer* buf) {
    Arg* a = nullptr;
    int rc = arg_at(k, index, "VDI", &a);
    if (rc != AD4CL_OK) return rc > 0 ? AD4CL_OK : rc;
    a->buf = buf;
    a->first_id = -1;
    a->set = buf != nullptr || a->kind != 'V';
    k->drop_graphs();
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_set_arg_int(ad4cl_kernel* k, int index, int value) {
    Arg* a = nullptr;
    int rc = arg_at(k, index, "i", &a);
    if (rc != AD4CL_OK) return rc > 0 ? AD4CL_OK : rc;
    if (k->use_static && (!a->set || a->ival != value)) k->ready = false;   /* register-tier instances freeze their small integer arguments */
    a->ival = value;
    a->set = true;
    k->drop_graphs();
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_set_arg_double(ad4cl_kernel* k, int index, double value) {
    Arg* a = nullptr;
    int rc = arg_at(k, index, "d", &a);
    if (rc != AD4CL_OK) return rc > 0 ? AD4CL_OK : rc;
    a->dval = value;
    a->set = true;
    k->drop_graphs();
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_set_arg_params(ad4cl_kernel* k, int index, int first_id, int count) {
    Arg* a = nullptr;
    int rc = arg_at(k, index, "V", &a);
    if (rc != AD4CL_OK) return rc > 0 ? AD4CL_OK : rc;
    if (first_id < 0 || count < 1) return fail(AD4CL_ERR_INVALID, "bad independent range [%d,+%d)", first_id, count);
    if (k->use_static && (a->first_id != first_id || a->count != count)) k->ready = false;
    a->first_id = first_id;
    a->count = count;
    a->buf = nullptr;
    a->set = true;
    k->drop_graphs();
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_config_default(ad4cl_kernel_config* cfg) {
    if (!cfg) return fail(AD4CL_ERR_INVALID, "cfg is NULL");
    memset(cfg, 0, sizeof *cfg);
    cfg->global_size[1] = 1;
    cfg->max_ops = 64;
    cfg->far_plane = 1;
    cfg->acc_mode = AD4CL_ACC_AUTO;
    cfg->tape_tier = AD4CL_TAPE_AUTO;
    cfg->epilogue = AD4CL_EPILOGUE_SUM;
    cfg->sweep_prefetch = AD4CL_PREFETCH_AUTO;
    cfg->recorder = AD4CL_RECORDER_AUTO;
    cfg->lockstep = AD4CL_LOCKSTEP_AUTO;
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_configure(ad4cl_kernel* k, const ad4cl_kernel_config* cfg) {
    if (!k || !cfg) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (cfg->global_size[0] <= 0) return fail(AD4CL_ERR_INVALID, "global_size[0] must be positive");
    if (cfg->max_ops <= 0) return fail(AD4CL_ERR_INVALID, "max_ops must be positive");
    const int quirks = k->cfg.reference_quirks;
    k->cfg = *cfg;
    k->cfg.reference_quirks = quirks; /* fixed when the kernel was created */
    k->configured = true;
    k->ready = false;
    return AD4CL_OK;
}

int ad4cl_cuda_eval_device(ad4cl_kernel* k, const double* theta_dev, int nparams, double* result_dev,
        int want_gradient, int apply_epilogue) {
    if (!k || !result_dev || (!theta_dev && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    if (k->tune_pending && want_gradient) {
        rc = autotune_recorder(k, theta_dev);
        if (rc != AD4CL_OK) return rc;
    }
    return enqueue(k, theta_dev, result_dev, want_gradient, apply_epilogue);
}

int ad4cl_cuda_eval_device_if(ad4cl_kernel* k, const double* theta_dev, int nparams, double* result_dev,
        int want_gradient, int apply_epilogue, const int* skip_dev) {
    if (!k || !result_dev || (!theta_dev && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    if (skip_dev != nullptr && !k->fused)
        return fail(AD4CL_ERR_INVALID, "kernel %s: a predicated evaluation needs the single-launch path "
                "(accumulator mode thread or warp)", k->name.c_str());
    return enqueue(k, theta_dev, result_dev, want_gradient, apply_epilogue, nullptr, nullptr, skip_dev);
}

ad4cl_context* ad4cl_cuda_kernel_context(ad4cl_kernel* k) { return k ? k->ctx : nullptr; }

int ad4cl_cuda_eval_device_sharded(ad4cl_kernel* k, ad4cl_comm* comm, const double* theta_dev, int nparams,
        double* result_dev, int want_gradient) {
    if (!k || !comm || !result_dev || (!theta_dev && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (comm->ctx != k->ctx) return fail(AD4CL_ERR_INVALID, "communicator and kernel belong to different contexts");
    if (!comm->connected && comm->world > 1) return fail(AD4CL_ERR_INVALID, "communicator is not connected");
    if (nparams + 3 > comm->stride - 1)
        return fail(AD4CL_ERR_INVALID, "vector of %d doubles exceeds the communicator's %d", nparams + 3, comm->stride - 1);
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    if (k->tune_pending && want_gradient) { /* local evaluations only: the peers take no part */
        rc = autotune_recorder(k, theta_dev);
        if (rc != AD4CL_OK) return rc;
    }
    return enqueue(k, theta_dev, result_dev, want_gradient, 1, nullptr, comm);
}

int ad4cl_cuda_check(ad4cl_kernel* k) {
    if (!k) return fail(AD4CL_ERR_INVALID, "kernel is NULL");
    if (!k->ready) return AD4CL_OK;
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(k->status_pinned, k->status_dev, sizeof (int), cudaMemcpyDeviceToHost, k->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(k->ctx->stream));
    return status_to_error(k, *k->status_pinned);
}

/* the host path on the stream: theta up, ONE launch, [grad, f, ops, slots, status] down */
static int enqueue_host_path(ad4cl_kernel* k, int want_gradient, ad4cl_comm* comm) {
    cudaStream_t st = k->ctx->stream;
    const int P = k->nparams;
    if (P > 0)
        CUDA_TRY(cudaMemcpyAsync(k->theta_dev, k->theta_pinned, sizeof (double) * (size_t) P, cudaMemcpyHostToDevice, st));
    int rc = enqueue(k, k->theta_dev, k->result_dev, want_gradient, 1, k->result_dev + P + 3, comm);
    if (rc != AD4CL_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(k->result_pinned, k->result_dev, sizeof (double) * (size_t) (P + 4), cudaMemcpyDeviceToHost, st));
    return AD4CL_OK;
}

static int eval_host(ad4cl_kernel* k, ad4cl_comm* comm, const double* theta, int nparams, double* f, double* grad) {
    if (!k || !f || (!theta && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    cudaStream_t st = k->ctx->stream;
    const int want = grad != nullptr ? 1 : 0;
    if (nparams > 0) memcpy(k->theta_pinned, theta, sizeof (double) * (size_t) nparams);
    if (k->tune_pending && want) { /* local evaluations only: the peers of a sharded call take no part */
        if (nparams > 0)
            CUDA_TRY(cudaMemcpyAsync(k->theta_dev, k->theta_pinned, sizeof (double) * (size_t) nparams, cudaMemcpyHostToDevice, st));
        rc = autotune_recorder(k, k->theta_dev);
        if (rc != AD4CL_OK) return rc;
        k->drop_graphs();
    }
    /* Replay the whole evaluation as one CUDA graph (captured on first use; the launch's
       arguments are fixed until an argument, the configuration or the communicator changes; the
       all-reduce epoch lives in device memory).  Profiling needs events around the kernel, so it
       takes the plain path. */
    if (k->graph_comm != comm) {
        k->drop_graphs();
        k->graph_comm = comm;
    }
    const bool use_graph = !k->profiling && !k->graph_disabled;
    if (use_graph && !k->graph[want]) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = enqueue_host_path(k, want, comm);
            cudaError_t e = cudaStreamEndCapture(st, &g);
            if (rc == AD4CL_OK && e == cudaSuccess && g && cudaGraphInstantiate(&k->graph[want], g, 0) != cudaSuccess)
                k->graph[want] = nullptr;
            if (g) cudaGraphDestroy(g);
            if (rc != AD4CL_OK) {
                cudaGetLastError();
                return rc; /* an argument error, not a capture problem */
            }
        }
        if (!k->graph[want]) { /* capture unavailable on this stream: fall back to plain launches for good */
            cudaGetLastError();
            k->graph_disabled = true;
        }
    }
    if (use_graph && k->graph[want]) {
        CUDA_TRY(cudaGraphLaunch(k->graph[want], st));
        k->stats.kernel_ms = 0.f;
    } else {
        CUDA_TRY(cudaEventRecord(k->ev0, st));
        rc = enqueue_host_path(k, want, comm);
        if (rc != AD4CL_OK) return rc;
        CUDA_TRY(cudaEventRecord(k->ev1, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!(use_graph && k->graph[want])) CUDA_TRY(cudaEventElapsedTime(&k->stats.kernel_ms, k->ev0, k->ev1));
    const int status = (int) k->result_pinned[nparams + 3];
    if (status & 4) {
        cudaMemsetAsync(k->status_dev, 0, sizeof (int), st);
        cudaStreamSynchronize(st);
        return fail(AD4CL_ERR_CUDA, "kernel %s: the all-reduce timed out waiting for a peer (rank %d of %d); "
                "the communicator must be rebuilt", k->name.c_str(), comm ? comm->rank : 0, comm ? comm->world : 1);
    }
    rc = status_to_error(k, status);
    if (rc != AD4CL_OK) return rc;
    *f = k->result_pinned[nparams];
    if (grad) memcpy(grad, k->result_pinned, sizeof (double) * (size_t) nparams);
    k->stats.ops = k->result_pinned[nparams + 1];
    k->stats.slots = k->result_pinned[nparams + 2];
    return AD4CL_OK;
}

int ad4cl_cuda_eval(ad4cl_kernel* k, const double* theta, int nparams, double* f, double* grad) {
    return eval_host(k, nullptr, theta, nparams, f, grad);
}

int ad4cl_cuda_eval_sharded(ad4cl_kernel* k, ad4cl_comm* comm, const double* theta, int nparams, double* f, double* grad) {
    if (!comm) return fail(AD4CL_ERR_INVALID, "communicator is NULL");
    if (k && comm->ctx != k->ctx) return fail(AD4CL_ERR_INVALID, "communicator and kernel belong to different contexts");
    if (!comm->connected && comm->world > 1) return fail(AD4CL_ERR_INVALID, "communicator is not connected");
    if (nparams + 3 > comm->stride - 1)
        return fail(AD4CL_ERR_INVALID, "vector of %d doubles exceeds the communicator's %d", nparams + 3, comm->stride - 1);
    return eval_host(k, comm, theta, nparams, f, grad);
}

int ad4cl_cuda_launches_per_eval(ad4cl_kernel* k) { return k ? k->launches_per_eval : 0; }

int ad4cl_cuda_eval_general(ad4cl_kernel* k, const double* theta, int nparams, double* out_values_dev,
        const double* seeds_dev, double* sum, double* grad) {
    if (!k || (!theta && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    cudaStream_t st = k->ctx->stream;
    if (nparams > 0) {
        memcpy(k->theta_pinned, theta, sizeof (double) * (size_t) nparams);
        CUDA_TRY(cudaMemcpyAsync(k->theta_dev, k->theta_pinned, sizeof (double) * (size_t) nparams, cudaMemcpyHostToDevice, st));
    }
    k->seeds_dev = seeds_dev;
    k->out_values_dev = out_values_dev;
    rc = enqueue(k, k->theta_dev, k->result_dev, grad != nullptr ? 1 : 0, 0, k->result_dev + nparams + 3);
    k->seeds_dev = nullptr;
    k->out_values_dev = nullptr;
    if (rc != AD4CL_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(k->result_pinned, k->result_dev, sizeof (double) * (size_t) (nparams + 4), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!k->fused) { /* the separate second stage writes the status only on the fused path */
        rc = ad4cl_cuda_check(k);
        if (rc != AD4CL_OK) return rc;
    } else {
        rc = status_to_error(k, (int) k->result_pinned[nparams + 3]);
        if (rc != AD4CL_OK) return rc;
    }
    if (sum) *sum = k->result_pinned[nparams];
    if (grad) memcpy(grad, k->result_pinned, sizeof (double) * (size_t) nparams);
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_profile(ad4cl_kernel* k, int enable) {
    if (!k) return fail(AD4CL_ERR_INVALID, "kernel is NULL");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    if (enable && k->prof_events.empty()) {
        k->prof_events.resize(2 * 64);
        for (auto& e : k->prof_events) CUDA_TRY(cudaEventCreate(&e));
    }
    if (!enable) {
        int rc = drain_profile(k);
        if (rc != AD4CL_OK) return rc;
    }
    k->profiling = enable != 0;
    return AD4CL_OK;
}

int ad4cl_cuda_kernel_profile_read(ad4cl_kernel* k, double* mean_ms, long long* launches, int reset) {
    if (!k || !mean_ms || !launches) return fail(AD4CL_ERR_INVALID, "NULL argument");
    int rc = drain_profile(k);
    if (rc != AD4CL_OK) return rc;
    *launches = k->prof_count;
    *mean_ms = k->prof_count ? k->prof_sum_ms / (double) k->prof_count : 0.0;
    if (reset) {
        k->prof_sum_ms = 0.0;
        k->prof_count = 0;
    }
    return AD4CL_OK;
}

static int comm_alloc(ad4cl_context* ctx, int rank, int world, int max_doubles, ad4cl_comm** out) {
    if (!ctx || !out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (world < 1 || world > ad4cl::kMaxWorld || rank < 0 || rank >= world || max_doubles < 1)
        return fail(AD4CL_ERR_INVALID, "bad communicator shape rank=%d world=%d n=%d (world <= %d)", rank, world,
                max_doubles, ad4cl::kMaxWorld);
    CUDA_TRY(cudaSetDevice(ctx->device));
    ad4cl_comm* c = new ad4cl_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    c->stride = max_doubles + 1; /* + the flag word */
    const size_t bytes = sizeof (double) * 2 * (size_t) world * c->stride;
    CUDA_TRY(cudaMalloc(&c->inbox, bytes + sizeof (unsigned long long)));
    CUDA_TRY(cudaMemset(c->inbox, 0, bytes + sizeof (unsigned long long)));
    CUDA_TRY(cudaMalloc(&c->status_dev, sizeof (int)));
    CUDA_TRY(cudaMemset(c->status_dev, 0, sizeof (int)));
    CUDA_TRY(cudaMallocHost(&c->status_pinned, sizeof (int)));
    CUDA_TRY(cudaDeviceSynchronize());
    c->info.inbox[rank] = c->inbox;
    c->info.epoch = reinterpret_cast<unsigned long long*> (c->inbox + 2 * (size_t) world * c->stride);
    c->info.rank = rank;
    c->info.world = world;
    c->info.stride = c->stride;
    c->info.timeout_ns = 30ull * 1000000000ull; /* ad4cl_cuda_comm_set_timeout */
    *out = c;
    return AD4CL_OK;
}

int ad4cl_cuda_comm_create(ad4cl_context* ctx, int rank, int world, int max_doubles, ad4cl_comm** out) {
    return comm_alloc(ctx, rank, world, max_doubles, out);
}

/* all ranks in ONE process (one context per GPU): peers are reached through plain peer access */
int ad4cl_cuda_comm_create_local(ad4cl_context** ctxs, int world, int max_doubles, ad4cl_comm** comms_out) {
    if (!ctxs || !comms_out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    for (int r = 0; r < world; r++) comms_out[r] = nullptr;
    for (int r = 0; r < world; r++) {
        int rc = comm_alloc(ctxs[r], r, world, max_doubles, &comms_out[r]);
        if (rc != AD4CL_OK) return rc;
        comms_out[r]->local_group = true;
    }
    for (int r = 0; r < world; r++) {
        CUDA_TRY(cudaSetDevice(ctxs[r]->device));
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            if (ctxs[q]->device != ctxs[r]->device) {
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device));
                if (!can) return fail(AD4CL_ERR_CUDA, "device %d cannot access device %d", ctxs[r]->device, ctxs[q]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(AD4CL_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctxs[r]->device, ctxs[q]->device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            comms_out[r]->info.inbox[q] = comms_out[q]->inbox;
        }
        comms_out[r]->connected = true;
    }
    return AD4CL_OK;
}

int ad4cl_cuda_comm_set_timeout(ad4cl_comm* c, double seconds) {
    if (!c || !(seconds > 0.0)) return fail(AD4CL_ERR_INVALID, "bad timeout");
    c->info.timeout_ns = (unsigned long long) (seconds * 1e9);
    return AD4CL_OK;
}

int ad4cl_cuda_comm_handle_bytes(void) { return (int) sizeof (cudaIpcMemHandle_t); }

int ad4cl_cuda_comm_handle(ad4cl_comm* c, void* handle_out) {
    if (!c || !handle_out) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->inbox));
    memcpy(handle_out, &h, sizeof h);
    return AD4CL_OK;
}

int ad4cl_cuda_comm_connect(ad4cl_comm* c, const void* handles) {
    if (!c || !handles) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*) handles + (size_t) r * sizeof h, sizeof h);
        void* p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->info.inbox[r] = (double*) p;
    }
    c->connected = true;
    return AD4CL_OK;
}

int ad4cl_cuda_comm_allreduce(ad4cl_comm* c, double* data_dev, int n, int nparams, int epilogue, double n_obs) {
    if (!c || !data_dev) return fail(AD4CL_ERR_INVALID, "NULL argument");
    if (!c->connected && c->world > 1) return fail(AD4CL_ERR_INVALID, "communicator is not connected");
    if (n < 1 || n > c->stride - 1) return fail(AD4CL_ERR_INVALID, "vector of %d doubles exceeds the communicator's %d", n, c->stride - 1);
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    oneshot_allreduce_kernel<<<1, 256, 0, c->ctx->stream>>>(c->info, n, data_dev, nparams, epilogue, n_obs, c->status_dev);
    CUDA_TRY(cudaGetLastError());
    return AD4CL_OK;
}

int ad4cl_cuda_comm_check(ad4cl_comm* c) {
    if (!c) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(c->status_pinned, c->status_dev, sizeof (int), cudaMemcpyDeviceToHost, c->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(c->ctx->stream));
    if (*c->status_pinned)
        return fail(AD4CL_ERR_CUDA, "one-shot all-reduce timed out waiting for a peer (rank %d of %d): the reduced vector "
                "holds NaN and the communicator must be rebuilt", c->rank, c->world);
    return AD4CL_OK;
}

int ad4cl_cuda_comm_destroy(ad4cl_comm* c) {
    if (!c) return AD4CL_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (!c->local_group)
        for (int r = 0; r < c->world; r++)
            if (r != c->rank && c->info.inbox[r]) cudaIpcCloseMemHandle(c->info.inbox[r]);
    cudaFree(c->inbox);
    cudaFree(c->status_dev);
    cudaFreeHost(c->status_pinned);
    delete c;
    return AD4CL_OK;
}

/*
 * Tape export (debug / legacy path).  Records only -- no sweep -- with one arena column per WORK-ITEM,
 * copies the rows to the host and rewrites them as the reference's 40-byte entries (ad.cl:75-79) in
 * work-item order, i.e. the stack a serial run of the reference would have produced, so that legacy host
 * code (host sum of out[], epilogue, compute_gradient: simple.cpp:357-369) keeps working unchanged.
 * Ids: independents keep theirs; the temporaries of work-item i become first_free_id + (operators of
 * the work-items before i) + operator index.  In-place operators appear with a fresh result id (same
 * sweep result); a leaf (ad_init_var_g) appears as an entry of size 0.  Small problems only: the arena
 * holds one warp region (ad4cl_tape.cuh: header, cell, lane-word and far planes) per 32 work-items.
 */
int ad4cl_cuda_export_tape(ad4cl_kernel* k, const double* theta, int nparams, int first_free_id,
        void* entries_out, long long capacity, long long* n_entries, void* out_vars, long long n_out) {
    struct HostEntry { double dx0; int id0; int pad0; double dx1; int id1; int pad1; int id; int size; };
    struct HostVar { double value; int id; int pad; };
    static_assert(sizeof (HostEntry) == 40 && sizeof (HostVar) == 16, "reference layouts (ad.cl:65-79)");
    if (!k || !entries_out || !n_entries || (!theta && nparams > 0)) return fail(AD4CL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(k->ctx->device));
    /* the export reads tape rows out of the arena: a kernel that would run in the register or the
       shared-memory tier is prepared for the arena tier for this call */
    struct TierGuard {
        ad4cl_kernel* k;
        int saved;
        ~TierGuard() {
            if (k->cfg.tape_tier != saved) {
                k->cfg.tape_tier = saved;
                k->ready = false;
            }
        }
    } guard{k, k->cfg.tape_tier};
    if (k->cfg.tape_tier != AD4CL_TAPE_HBM) {
        k->cfg.tape_tier = AD4CL_TAPE_HBM;
        k->ready = false;
    }
    int rc = prepare(k, nparams);
    if (rc != AD4CL_OK) return rc;
    cudaStream_t st = k->ctx->stream;
    const long long items = k->L.n_items;
    const int B = k->block, nw = B / 32;
    const long long ngroups = (items + B - 1) / B;
    const size_t region = k->L.warp_region;
    const size_t arena_bytes = (size_t) ngroups * nw * region;
    if (arena_bytes > ((size_t) 8 << 30))
        return fail(AD4CL_ERR_NOMEM, "tape export of %lld work-items needs a %zu-byte arena: this path is for small problems", items, arena_bytes);
    char* arena = nullptr;
    double *meta = nullptr, *partials = nullptr;
    CUDA_TRY(cudaMalloc(&arena, arena_bytes));
    CUDA_TRY(cudaMemsetAsync(arena, 0, arena_bytes, st));
    CUDA_TRY(cudaMalloc(&meta, sizeof (double) * 3 * (size_t) items));
    CUDA_TRY(cudaMemsetAsync(meta, 0, sizeof (double) * 3 * (size_t) items, st));
    CUDA_TRY(cudaMalloc(&partials, sizeof (double) * (size_t) ngroups * (nparams + 3)));

    /* record-only launch: same entry, one work-group per CTA, private arena */
    const ad4cl::LaunchInfo saved = k->L;
    const int saved_grid = k->grid;
    double* const saved_partials = k->partials;
    k->partials = partials;
    k->L.arena = arena;
    k->L.partials = partials;
    k->L.retain_meta = meta;
    k->L.tape_in_smem = 0;
    k->grid = (int) ngroups;
    if (nparams > 0) {
        memcpy(k->theta_pinned, theta, sizeof (double) * (size_t) nparams);
        CUDA_TRY(cudaMemcpyAsync(k->theta_dev, k->theta_pinned, sizeof (double) * (size_t) nparams, cudaMemcpyHostToDevice, st));
    }
    rc = enqueue(k, k->theta_dev, k->result_dev, 1, 0);
    k->L = saved;
    k->grid = saved_grid;
    k->partials = saved_partials;
    std::vector<char> rows(arena_bytes);
    std::vector<double> m(3 * (size_t) items);
    if (rc == AD4CL_OK) {
        cudaError_t e = cudaMemcpyAsync(rows.data(), arena, arena_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m.data(), meta, sizeof (double) * m.size(), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(k->status_pinned, k->status_dev, sizeof (int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(AD4CL_ERR_CUDA, "tape export copy failed: %s", cudaGetErrorString(e));
    }
    cudaFree(arena);
    cudaFree(meta);
    cudaFree(partials);
    if (rc != AD4CL_OK) return rc;
    rc = status_to_error(k, *k->status_pinned);
    if (rc != AD4CL_OK) return rc;

    /* rows -> reference entries, work-item by work-item */
    HostEntry* E = static_cast<HostEntry*> (entries_out);
    HostVar* O = static_cast<HostVar*> (out_vars);
    long long ne = 0;
    int next_id = first_free_id;
    std::vector<int> op_of_slot;
    for (long long item = 0; item < items; item++) {
        /* work-item -> (CTA, thread): the NDRange walk of run_objective (1-D and 2-D) */
        const int g0 = k->L.g0, l0 = k->L.l0, l1 = B / l0;
        const long long gid0 = item % g0, gid1 = item / g0;
        const long long grp = (gid1 / l1) * (g0 / l0) + gid0 / l0;
        const int tid = (int) ((gid1 % l1) * l0 + gid0 % l0);
        const char* col = rows.data() + ((size_t) grp * nw + tid / 32) * region;
        const ad4cl::RegionLayout lay = ad4cl::region_layout(k->L.cap_slots, k->L.use_far != 0);
        const int lane = tid % 32;
        const int nslots = (int) m[3 * item + 2];
        op_of_slot.assign((size_t) nslots, -1);
        const int base = next_id;
        int op = 0;
        HostEntry cur{};
        int have = 0;
        for (int s = 0; s < nslots; s++) {
            /* the export launch records with the general recorder: row s = 384 bytes, the 32 lanes' partials
               then their id words, after the two constant cells of the cell plane (ad4cl_tape.cuh) */
            const char* row = col + lay.off_cells + (size_t) ad4cl::kUnitCells * ad4cl::kCellRowBytes + (size_t) s * ad4cl::kRowBytes;
            int word;
            memcpy(&word, row + ad4cl::kRowIdOffset + lane * 4, 4);
            double dx;
            memcpy(&dx, row + lane * 8, 8);
            if (!(word & ad4cl::kNopFlag)) {
                const int payload = word & ad4cl::kIdMask;
                const int id = (word & ad4cl::kParamFlag) ? payload : base + op_of_slot[(size_t) payload];
                if (have == 0) { cur.dx0 = dx; cur.id0 = id; }
                else { cur.dx1 = dx; cur.id1 = id; }
                have++;
            }
            if (word & ad4cl::kEndFlag) {
                if (ne >= capacity) return fail(AD4CL_ERR_TAPE_OVERFLOW, "the exported tape exceeds the caller's %lld entries", capacity);
                cur.id = base + op;
                cur.size = have;
                E[ne++] = cur;
                op_of_slot[(size_t) s] = op++;
                cur = HostEntry{};
                have = 0;
            }
        }
        next_id = base + op;
        if (O && item < n_out) {
            const int oid = (int) m[3 * item + 1];
            O[item].value = m[3 * item];
            O[item].id = oid < 0 ? 0 : (oid < nparams ? oid : base + op_of_slot[(size_t) (oid - nparams)]);
            O[item].pad = 0;
        }
    }
    *n_entries = ne;
    return AD4CL_OK;
}

int ad4cl_cuda_last_stats(ad4cl_kernel* k, ad4cl_eval_stats* stats) {
    if (!k || !stats) return fail(AD4CL_ERR_INVALID, "NULL argument");
    *stats = k->stats;
    return AD4CL_OK;
}

} // extern "C"
#pragma GCC visibility pop
