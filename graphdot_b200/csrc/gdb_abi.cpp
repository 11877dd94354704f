// gdb_abi.cpp -- host side of the C ABI in include/graphdot_b200.h:
// device context, NVRTC program cache, graph-set upload and solve launches.
//
// Replaces the pycuda plumbing of reference
// graphdot/kernel/marginalized/_backend_cuda.py (context :49-61, JIT :118-155,
// code generation :157-228, launch :247-367) and graphdot/cuda/*.py.
// CUDA runtime (static) is used for memory/streams/events on the device's
// primary context; the driver entry points needed for NVRTC modules are
// fetched through cudaGetDriverEntryPoint, so the library loads (and its
// pure-host functions work) on machines without a GPU driver.
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <sstream>
#include <string>
#include <vector>

#include "gdb_internal.h"

extern const char *const gdb_embedded_prelude;
extern const char *const gdb_embedded_solver;
extern const char *const gdb_embedded_small;
extern const char *const gdb_embedded_large;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string t_last_error;

int gdb_fail(int code, const char *fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

extern "C" const char *gdb_last_error(void) { return t_last_error.c_str(); }
extern "C" const char *gdb_version(void) { return "graphdot_b200 0.1.0 (sm_100a, NVRTC)"; }
extern "C" void gdb_free(void *p) { free(p); }

extern "C" const char *gdb_solver_template(void) {
    static std::string joined = std::string(gdb_embedded_prelude) + "\n/* <generated splice> */\n" + gdb_embedded_solver +
                                "\n" + gdb_embedded_small + "\n" + gdb_embedded_large;
    return joined.c_str();
}

#define RT(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return gdb_fail(GDB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

#define DRV(ctx, call)                                                                   \
    do {                                                                                 \
        CUresult r_ = (call);                                                            \
        if (r_ != CUDA_SUCCESS) {                                                        \
            const char *s_ = nullptr;                                                    \
            if ((ctx)->cuGetErrorString) (ctx)->cuGetErrorString(r_, &s_);               \
            return gdb_fail(GDB_ERR_CUDA, "%s: %s", #call, s_ ? s_ : "unknown driver error"); \
        }                                                                                \
    } while (0)

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
};

struct gdb_context_s {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host copies of finished column blocks
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_copy = nullptr;        // last copy issued on copy_stream
    std::vector<cudaEvent_t> tile_ev;     // per launch: kernel done, copy done
    struct gdb_pool *pool = nullptr;      // host threads of the collection step
    // driver entry points
    CUresult (*cuModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*cuModuleUnload)(CUmodule) = nullptr;
    CUresult (*cuModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*cuModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
    CUresult (*cuFuncGetAttribute)(int *, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                               CUstream, void **, void **) = nullptr;
    CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*cuGetErrorString)(CUresult, const char **) = nullptr;
    // persistent device buffers, grown on demand
    DevBuf jobs, starts, gram, grad, scratch, counters, norm_diag, norm_ddiag;
    uint32_t norm_n = 0, norm_nj = 0;
    gdb_graphset_t norm_gs = nullptr;  // graph set the stored self-similarities belong to
    std::mutex mu;
    std::map<std::string, gdb_program_t> programs;  // source text -> program
};

static int dev_reserve(DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return GDB_OK;
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    RT(cudaMalloc(&b.ptr, want));
    b.cap = want;
    return GDB_OK;
}

template<class F> static int load_entry(const char *name, F *&fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    RT(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st));
    if (!p || st != cudaDriverEntryPointSuccess) return gdb_fail(GDB_ERR_CUDA, "driver entry point %s unavailable", name);
    fn = reinterpret_cast<F *>(p);
    return GDB_OK;
}

extern "C" int gdb_context_create(int device, gdb_context_t *out) {
    if (!out) return gdb_fail(GDB_ERR_INVALID, "gdb_context_create: null out");
    int n = 0;
    RT(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return gdb_fail(GDB_ERR_CUDA, "CUDA device %d not present (%d visible)", device, n);
    RT(cudaSetDevice(device));
    RT(cudaFree(0));  // establishes the primary context
    gdb_context_s *c = new gdb_context_s;
    c->device = device;
    RT(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major < 10)
        fprintf(stderr, "graphdot_b200: warning: device %s is sm_%d%d; kernels are built for sm_100a\n", c->prop.name,
                c->prop.major, c->prop.minor);
    RT(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RT(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) RT(cudaEventCreate(&e));
    RT(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    RT(cudaEventRecord(c->ev_copy, c->copy_stream));
    int rc;
    if ((rc = load_entry("cuModuleLoadData", c->cuModuleLoadData))) return rc;
    if ((rc = load_entry("cuModuleUnload", c->cuModuleUnload))) return rc;
    if ((rc = load_entry("cuModuleGetFunction", c->cuModuleGetFunction))) return rc;
    if ((rc = load_entry("cuModuleGetGlobal", c->cuModuleGetGlobal))) return rc;
    if ((rc = load_entry("cuFuncGetAttribute", c->cuFuncGetAttribute))) return rc;
    if ((rc = load_entry("cuFuncSetAttribute", c->cuFuncSetAttribute))) return rc;
    if ((rc = load_entry("cuLaunchKernel", c->cuLaunchKernel))) return rc;
    if ((rc = load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", c->cuOccupancyMaxActiveBlocksPerMultiprocessor)))
        return rc;
    if ((rc = load_entry("cuGetErrorString", c->cuGetErrorString))) return rc;
    RT(cudaMalloc(&c->counters.ptr, 4096));
    c->counters.cap = 4096;
    *out = c;
    return GDB_OK;
}

extern "C" int gdb_context_info(gdb_context_t c, gdb_device_info *o) {
    if (!c || !o) return gdb_fail(GDB_ERR_INVALID, "gdb_context_info: null argument");
    memset(o, 0, sizeof *o);
    o->device = c->device;
    o->sm_count = c->prop.multiProcessorCount;
    o->cc_major = c->prop.major;
    o->cc_minor = c->prop.minor;
    o->max_smem_per_block_optin = (int)c->prop.sharedMemPerBlockOptin;
    o->max_smem_per_sm = (int)c->prop.sharedMemPerMultiprocessor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    o->clock_khz = khz;
    o->l2_bytes = c->prop.l2CacheSize;
    o->total_mem = c->prop.totalGlobalMem;
    snprintf(o->name, sizeof o->name, "%s", c->prop.name);
    return GDB_OK;
}

extern "C" int gdb_context_synchronize(gdb_context_t c) {
    if (!c) return gdb_fail(GDB_ERR_INVALID, "null context");
    RT(cudaSetDevice(c->device));
    RT(cudaStreamSynchronize(c->stream));
    RT(cudaStreamSynchronize(c->copy_stream));
    RT(cudaGetLastError());
    return GDB_OK;
}

extern "C" int gdb_program_destroy(gdb_program_t p);
void gdb_pool_destroy(struct gdb_pool *p);

extern "C" int gdb_context_destroy(gdb_context_t c) {
    if (!c) return GDB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->programs) {
        gdb_program_t p = kv.second;
        kv.second = nullptr;
        (void)p;  // programs are owned by their handles; the cache only aliases them
    }
    for (DevBuf *b : {&c->jobs, &c->starts, &c->gram, &c->grad, &c->scratch, &c->counters, &c->norm_diag, &c->norm_ddiag})
        if (b->ptr) cudaFree(b->ptr);
    for (auto &e : c->ev) cudaEventDestroy(e);
    for (auto &e : c->tile_ev) cudaEventDestroy(e);
    cudaEventDestroy(c->ev_copy);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    if (c->pool) gdb_pool_destroy(c->pool);
    delete c;
    return GDB_OK;
}

extern "C" int gdb_host_alloc(size_t bytes, void **out) {
    if (!out) return gdb_fail(GDB_ERR_INVALID, "gdb_host_alloc: null out");
    *out = nullptr;
    RT(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return GDB_OK;
}

extern "C" int gdb_host_free(void *p) {
    if (p) RT(cudaFreeHost(p));
    return GDB_OK;
}

extern "C" int gdb_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return gdb_fail(GDB_ERR_INVALID, "gdb_host_register: null argument");
    RT(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return GDB_OK;
}

extern "C" int gdb_host_unregister(void *p) {
    if (p) RT(cudaHostUnregister(p));
    return GDB_OK;
}

// ---------------------------------------------------------------------------
// program
// ---------------------------------------------------------------------------
struct gdb_program_s {
    gdb_context_t ctx = nullptr;
    // one NVRTC module per kernel: the general and the small-pair kernel are compiled
    // concurrently when the program is created; the large-pair kernel (80 % of the
    // compile time of the whole template) only when a graph set first needs it
    CUmodule mod = nullptr, mod_small = nullptr, mod_large = nullptr;
    std::string extra;
    std::mutex large_mu;
    bool large_tried = false;
    CUfunction fn = nullptr;        // mlgk_solve: any pair size
    CUfunction fn_small = nullptr;  // mlgk_solve_small: pair resident in shared memory
    CUfunction fn_large = nullptr;  // mlgk_solve_large: one cluster per pair (graph-level outputs)
    int small_regs = 0, small_static_smem = 0, large_static_smem = 0;
    int cluster = 4, lcpt = 16, lell = 12;
    uint32_t edge_size = 0, ell_entry = 8, large_block = 1024, large_ltr = 1;
    unsigned layout[8] = {};
    std::string source, log;
    gdb_program_info info{};
    uint32_t theta_size[3] = {};
    int eval_gradient = 0, nodal = 0, symmetric = 0, diagonal = 0, wpt = 1, rpw = 8, adj = 4;
    int refcount = 1;
};

static void emit_functor(std::ostringstream &o, const char *name, const gdb_functor_src &f, bool binary) {
    const char *args = binary ? "X const &x1, X const &x2" : "X const &n";
    o << "struct " << name << "_theta_t { " << (f.theta_decl ? f.theta_decl : "") << " };\n";
    o << "struct " << name << "_t : " << name << "_theta_t {\n";
    o << "    static constexpr int jac_dims = " << f.n_jac << ";\n";
    o << "    template<class X> __device__ __forceinline__ float operator()(" << args << ") const {\n";
    o << "        return (" << f.expr << ");\n    }\n";
    o << "    template<class X> __device__ __forceinline__ void jacobian(" << args << ", float *j) const {\n";
    for (uint32_t k = 0; k < f.n_jac; ++k) o << "        j[" << k << "] = (" << f.jac[k] << ");\n";
    o << "    }\n};\n";
}

static int pick_block(const gdb_program_desc *d) {
    int b = d->block_size;
    if (b <= 0) b = 64;
    if (b % 32 || b > 1024) return -1;
    return b;
}

static int pick_wpt(const gdb_program_desc *d) { return d->workers_per_thread <= 0 ? 1 : d->workers_per_thread; }

static int pick_rpw(const gdb_program_desc *d) { return d->rows_per_warp <= 0 ? 8 : d->rows_per_warp; }
static int pick_adj(const gdb_program_desc *d) { return d->slots_per_lane <= 0 ? 4 : d->slots_per_lane; }
static int pick_cluster(const gdb_program_desc *d) {
    int c = d->cluster_size <= 0 ? 2 : d->cluster_size;
    if (const char *env = getenv("GDB_CLUSTER")) c = atoi(env);  // tuning hook
    return c;
}
static int pick_lcpt(const gdb_program_desc *d) { return d->cols_per_lane <= 0 ? 16 : d->cols_per_lane; }
static int pick_lell(const gdb_program_desc *d) { return d->ell_slots <= 0 ? 12 : d->ell_slots; }

static int render(const gdb_program_desc *d, std::string &src) {
    if (!d || !d->node_decl || !d->edge_decl || !d->node_kernel.expr || !d->edge_kernel.expr || !d->p_start.expr)
        return gdb_fail(GDB_ERR_INVALID, "gdb_program_desc: missing source strings");
    const int block = pick_block(d);
    if (block < 0) return gdb_fail(GDB_ERR_INVALID, "block_size must be a multiple of 32 up to 1024");
    if (d->nodal < 0 || d->nodal > 2 || d->lmin < 0 || d->lmin > 1) return gdb_fail(GDB_ERR_INVALID, "invalid traits");
    if (pick_wpt(d) > 4) return gdb_fail(GDB_ERR_INVALID, "workers_per_thread must be 1..4");
    if (pick_rpw(d) > 8) return gdb_fail(GDB_ERR_INVALID, "rows_per_warp must be 1..8");
    if (pick_adj(d) != 2 && pick_adj(d) != 4) return gdb_fail(GDB_ERR_INVALID, "slots_per_lane must be 2 or 4");
    {
        const int c = pick_cluster(d);
        if (c != 1 && c != 2 && c != 4 && c != 8) return gdb_fail(GDB_ERR_INVALID, "cluster_size must be 1, 2, 4 or 8");
        if (pick_lcpt(d) > 32) return gdb_fail(GDB_ERR_INVALID, "cols_per_lane must be 1..32");
        if (pick_lell(d) > 64) return gdb_fail(GDB_ERR_INVALID, "ell_slots must be 1..64");
    }
    std::ostringstream o;
    o << gdb_embedded_prelude << "\n";
    o << "// ---- generated splice ----\n";
    o << "#define GDB_BLOCK " << block << "\n";
    o << "#define GDB_MIN_BLOCKS " << std::max(1, std::min(16, 1024 / block)) << "\n";
    {
        // small-pair kernel: the per-thread state is rows_per_warp x (x, r, Ap, diag)
        // [float2 with gradients] + ~45 registers of indices and loop state.  The
        // kernel is bound by instruction issue and latency, so more and lighter warps
        // win: with at most 6 rows per warp ask ptxas for 640 threads per SM (96
        // registers, no spills) instead of 512.  Measured on the C3 workload
        // (M pairs/s): block 96 / 8 rows / 128 regs 12.4; 128 / 6 / 96 regs 17.6;
        // 128 / 6 / 80 regs (A p spilled) 17.6; 72 regs 11.6.
        int threads = (pick_wpt(d) == 1 && pick_rpw(d) <= 6) ? 640 : 512;
        if (const char *env = getenv("GDB_SMALL_THREADS")) threads = atoi(env);  // tuning hook
        o << "#define GDB_MIN_BLOCKS_SMALL " << std::max(1, threads / block) << "\n";
    }
    o << "#define GDB_WPT " << pick_wpt(d) << "\n";
    o << "#define GDB_RPW " << pick_rpw(d) << "\n";
    o << "#define GDB_ADJ " << pick_adj(d) << "\n";
    o << "#define GDB_CLUSTER " << pick_cluster(d) << "\n";
    o << "#define GDB_LCPT " << pick_lcpt(d) << "\n";
    o << "#define GDB_LELL " << pick_lell(d) << "\n";
    o << "#define GDB_WEIGHTED " << (d->weighted ? 1 : 0) << "\n";
    o << "#define GDB_DIAGONAL " << (d->diagonal ? 1 : 0) << "\n";
    o << "#define GDB_SYMMETRIC " << (d->symmetric ? 1 : 0) << "\n";
    o << "#define GDB_NODAL " << d->nodal << "\n";
    o << "#define GDB_LMIN " << d->lmin << "\n";
    o << "#define GDB_GRADIENT " << (d->eval_gradient ? 1 : 0) << "\n";
    o << "#define GDB_NP " << d->p_start.n_jac << "\n";
    o << "#define GDB_NV " << d->node_kernel.n_jac << "\n";
    o << "#define GDB_NE " << d->edge_kernel.n_jac << "\n";
    o << "struct node_t { " << d->node_decl << " };\n";
    o << "struct edge_label_t { " << d->edge_decl << " };\n";
    o << "static_assert(sizeof(node_t) == " << d->node_size << ", \"node_t layout differs from the host dtype\");\n";
    o << "static_assert(sizeof(edge_label_t) == " << d->edge_label_size
      << ", \"edge label layout differs from the host dtype\");\n";
    o << "static_assert(alignof(edge_label_t) == " << d->edge_label_align << ", \"edge label alignment\");\n";
    emit_functor(o, "node_kernel", d->node_kernel, true);
    emit_functor(o, "edge_kernel", d->edge_kernel, true);
    emit_functor(o, "p_start", d->p_start, false);
    const gdb_functor_src *fs[3] = {&d->node_kernel, &d->edge_kernel, &d->p_start};
    const char *names[3] = {"node_kernel", "edge_kernel", "p_start"};
    for (int k = 0; k < 3; ++k)
        if (fs[k]->theta_size)
            o << "static_assert(sizeof(" << names[k] << "_theta_t) == " << fs[k]->theta_size << ", \"" << names[k]
              << " hyper-parameter layout differs from the host dtype\");\n";
    o << "// ---- end of generated splice ----\n";
    o << gdb_embedded_solver << "\n" << gdb_embedded_small << "\n" << gdb_embedded_large << "\n";
    src = o.str();
    return GDB_OK;
}

extern "C" int gdb_render_source(const gdb_program_desc *d, char **out) {
    if (!out) return gdb_fail(GDB_ERR_INVALID, "null out");
    std::string s;
    int rc = render(d, s);
    if (rc) return rc;
    *out = static_cast<char *>(malloc(s.size() + 1));
    if (!*out) return gdb_fail(GDB_ERR_NOMEM, "out of memory");
    memcpy(*out, s.c_str(), s.size() + 1);
    return GDB_OK;
}

static int compile_cubin(const std::string &src, const char *extra, int mask, std::vector<char> &cubin, std::string &log) {
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src.c_str(), "mlgk_solver.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return gdb_fail(GDB_ERR_COMPILE, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "--use_fast_math", "-lineinfo",
                                     "-default-device", "--extra-device-vectorization"};
    opts.push_back("-DGDB_BUILD_MASK=" + std::to_string(mask));
    if (extra) {
        std::istringstream is(extra);
        std::string tok;
        while (is >> tok) opts.push_back(tok);
    }
    std::vector<const char *> copts;
    for (auto &s : opts) copts.push_back(s.c_str());
    r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
    size_t ls = 0;
    nvrtcGetProgramLogSize(prog, &ls);
    log.assign(ls ? ls - 1 : 0, '\0');
    if (ls > 1) nvrtcGetProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return gdb_fail(GDB_ERR_COMPILE, "NVRTC compilation failed: %s\n%s", nvrtcGetErrorString(r), log.c_str());
    }
    size_t cs = 0;
    nvrtcGetCUBINSize(prog, &cs);
    cubin.resize(cs);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return GDB_OK;
}

// kernels 0..n-1 (masks 1, 2, 4) compiled on their own threads; the error text of a thread
// is carried back explicitly because gdb_last_error is per thread
static void compile_parallel(const std::string &src, const char *extra, int n, std::vector<char> *cubin,
                             std::string *logs, int *rcs, std::string *errs) {
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k)
        th.emplace_back([&, k] {
            rcs[k] = compile_cubin(src, extra, 1 << k, cubin[k], logs[k]);
            if (rcs[k]) errs[k] = gdb_last_error();
        });
    for (auto &t : th) t.join();
}

extern "C" int gdb_program_compile_only(const gdb_program_desc *d, uint64_t *cubin_bytes) {
    std::string src, log;
    int rc = render(d, src);
    if (rc) return rc;
    std::vector<char> cubin[3];
    std::string logs[3], errs[3];
    int rcs[3] = {0, 0, 0};
    const int n = d->nodal == GDB_NODAL_NONE ? 3 : 2;
    compile_parallel(src, d->extra_options, n, cubin, logs, rcs, errs);
    for (int k = 0; k < n; ++k)
        if (rcs[k]) return gdb_fail(rcs[k], "%s", errs[k].c_str());
    if (cubin_bytes) *cubin_bytes = cubin[0].size() + cubin[1].size() + cubin[2].size();
    return GDB_OK;
}

extern "C" int gdb_program_create(gdb_context_t c, const gdb_program_desc *d, gdb_program_t *out) {
    if (!c || !out) return gdb_fail(GDB_ERR_INVALID, "gdb_program_create: null argument");
    std::string src;
    int rc = render(d, src);
    if (rc) return rc;
    std::string key = src + "\n//opts:" + (d->extra_options ? d->extra_options : "");
    {
        std::lock_guard<std::mutex> lk(c->mu);
        auto it = c->programs.find(key);
        if (it != c->programs.end() && it->second) {
            it->second->refcount++;
            it->second->info.from_cache = 1;
            *out = it->second;
            return GDB_OK;
        }
    }
    RT(cudaSetDevice(c->device));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin[2];
    std::string logs[2], errs[2];
    int rcs[2] = {0, 0};
    compile_parallel(src, d->extra_options, 2, cubin, logs, rcs, errs);
    for (int k = 0; k < 2; ++k)
        if (rcs[k]) return gdb_fail(rcs[k], "%s", errs[k].c_str());
    gdb_program_s *p = new gdb_program_s;
    p->ctx = c;
    p->source = src;
    p->extra = d->extra_options ? d->extra_options : "";
    p->log = logs[0] + logs[1];
    p->eval_gradient = d->eval_gradient;
    p->nodal = d->nodal;
    p->symmetric = d->symmetric ? 1 : 0;
    p->diagonal = d->diagonal ? 1 : 0;
    p->wpt = pick_wpt(d);
    p->rpw = pick_rpw(d);
    p->adj = pick_adj(d);
    p->theta_size[0] = d->node_kernel.theta_size;
    p->theta_size[1] = d->edge_kernel.theta_size;
    p->theta_size[2] = d->p_start.theta_size;
    DRV(c, c->cuModuleLoadData(&p->mod, cubin[0].data()));
    DRV(c, c->cuModuleLoadData(&p->mod_small, cubin[1].data()));
    DRV(c, c->cuModuleGetFunction(&p->fn, p->mod, "mlgk_solve"));
    CUdeviceptr lay = 0;
    size_t lay_bytes = 0;
    DRV(c, c->cuModuleGetGlobal(&lay, &lay_bytes, p->mod, "gdb_param_layout"));
    if (lay_bytes != sizeof p->layout) return gdb_fail(GDB_ERR_LAYOUT, "gdb_param_layout has %zu bytes", lay_bytes);
    RT(cudaMemcpy(p->layout, reinterpret_cast<void *>(lay), sizeof p->layout, cudaMemcpyDeviceToHost));
    if (p->layout[4] != sizeof(gdb_params_fixed_host))
        return gdb_fail(GDB_ERR_LAYOUT, "device gdb_params_fixed is %u bytes, host mirror %zu", p->layout[4],
                        sizeof(gdb_params_fixed_host));
    if (p->layout[5] != d->node_size) return gdb_fail(GDB_ERR_LAYOUT, "node_t is %u bytes on device", p->layout[5]);
    int v = 0;
    p->info.block_size = pick_block(d);
    c->cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, p->fn);
    p->info.num_regs = v;
    c->cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, p->fn);
    p->info.static_smem = v;
    c->cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, p->fn);
    p->info.local_bytes = v;
    p->info.max_dynamic_smem = (int)c->prop.sharedMemPerBlockOptin - p->info.static_smem;
    DRV(c, c->cuFuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->info.max_dynamic_smem));
    DRV(c, c->cuModuleGetFunction(&p->fn_small, p->mod_small, "mlgk_solve_small"));
    c->cuFuncGetAttribute(&p->small_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, p->fn_small);
    c->cuFuncGetAttribute(&p->small_static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, p->fn_small);
    DRV(c, c->cuFuncSetAttribute(p->fn_small, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                 (int)c->prop.sharedMemPerBlockOptin - p->small_static_smem));
    p->info.num_regs_small = p->small_regs;
    p->cluster = pick_cluster(d);
    p->lcpt = pick_lcpt(d);
    p->lell = pick_lell(d);
    p->edge_size = p->layout[6];
    p->info.n_jac = (int)p->layout[7];
    p->info.from_cache = 0;
    p->info.compile_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->programs[key] = p;
        p->refcount++;  // the cache keeps programs alive for the context's lifetime
    }
    *out = p;
    return GDB_OK;
}

extern "C" int gdb_program_info_get(gdb_program_t p, gdb_program_info *o) {
    if (!p || !o) return gdb_fail(GDB_ERR_INVALID, "null argument");
    *o = p->info;
    return GDB_OK;
}
extern "C" const char *gdb_program_log(gdb_program_t p) { return p ? p->log.c_str() : ""; }
extern "C" const char *gdb_program_source(gdb_program_t p) { return p ? p->source.c_str() : ""; }

extern "C" int gdb_program_destroy(gdb_program_t p) {
    if (!p) return GDB_OK;
    if (--p->refcount > 0) return GDB_OK;
    for (CUmodule m : {p->mod, p->mod_small, p->mod_large})
        if (m && p->ctx) p->ctx->cuModuleUnload(m);
    delete p;
    return GDB_OK;
}

// ---------------------------------------------------------------------------
// graph set
// ---------------------------------------------------------------------------
struct gdb_graphset_s {
    gdb_context_t ctx = nullptr;
    gdb_layout layout{};
    uint32_t n = 0;
    uint8_t *host = nullptr;  // pinned image: [table | blobs]
    uint8_t *dev = nullptr;
    uint64_t bytes = 0, table_bytes = 0;
    std::vector<uint32_t> n_node, blob_bytes, nnz;
    uint32_t max_blob[2] = {0, 0};  // two largest blobs
    uint32_t max_node[2] = {0, 0};  // two largest node counts
    uint32_t max_nnz[2] = {0, 0};   // two largest element counts
    // small-pair kernel: most neighbour slots of one graph left without a helper lane,
    // for [slots per lane 2, 4][workers per thread 1..4] (mirrors the lane tables of
    // mlgk_small.cuh)
    uint32_t max_ovf[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    uint32_t max_idx[2] = {0, 0};   // two largest (row index + edge elements) byte counts
    uint32_t max_tc = 0, max_degree = 0;  // large-pair kernel: longest neighbour-row list, largest degree
    uint32_t max_tile_nnz = 0;            // ... and most elements in one tile row (8 rows)
    uint64_t sum_node = 0;
    bool index16 = true;            // every graph carries a valid 16-bit row index
};

extern "C" int gdb_graphset_create(gdb_context_t c, const gdb_layout *L, uint32_t n, const void *const *blobs,
                                   const uint64_t *blob_bytes, gdb_graphset_t *out) {
    if (!c || !L || !n || !blobs || !blob_bytes || !out) return gdb_fail(GDB_ERR_INVALID, "gdb_graphset_create: null argument");
    RT(cudaSetDevice(c->device));
    gdb_graphset_s *gs = new gdb_graphset_s;
    gs->ctx = c;
    gs->layout = *L;
    gs->n = n;
    gs->table_bytes = ((uint64_t)n * sizeof(gdb_graph_ref_host) + 255u) & ~255ull;
    uint64_t total = gs->table_bytes;
    for (uint32_t k = 0; k < n; ++k) {
        if (blob_bytes[k] % 16 || blob_bytes[k] < GDB_HDR_BYTES) {
            delete gs;
            return gdb_fail(GDB_ERR_INVALID, "graph %u: malformed blob", k);
        }
        total += blob_bytes[k];
    }
    gs->bytes = total;
    {
        cudaError_t e1 = cudaHostAlloc((void **)&gs->host, total, cudaHostAllocPortable);
        cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc((void **)&gs->dev, total) : e1;
        if (e2 != cudaSuccess) {
            if (gs->host) cudaFreeHost(gs->host);
            delete gs;
            return gdb_fail(GDB_ERR_CUDA, "graph set of %llu bytes: %s", (unsigned long long)total, cudaGetErrorString(e2));
        }
    }
    gdb_graph_ref_host *table = reinterpret_cast<gdb_graph_ref_host *>(gs->host);
    uint64_t off = gs->table_bytes;
    for (uint32_t k = 0; k < n; ++k) {
        uint8_t *dst = gs->host + off;
        memcpy(dst, blobs[k], blob_bytes[k]);
        const gdb_graph_hdr_host *h = reinterpret_cast<const gdb_graph_hdr_host *>(dst);
        if (h->blob_bytes != blob_bytes[k]) {
            gdb_graphset_destroy(gs);
            return gdb_fail(GDB_ERR_INVALID, "graph %u: blob size mismatch", k);
        }
        const uint64_t dev_base = reinterpret_cast<uint64_t>(gs->dev) + off;
        // relocate frozen_array data pointers: blob-relative -> device address
        for (int32_t i = 0; i < h->n_node; ++i)
            for (uint32_t f = 0; f < L->n_node_ptr; ++f)
                *reinterpret_cast<uint64_t *>(dst + h->off_node + (size_t)i * L->node_size + L->node_ptr_offset[f]) += dev_base;
        if (L->n_edge_ptr) {
            uint32_t label_off = 0, edge_size = L->edge_label_size;
            if (L->weighted) {
                const uint32_t a = std::max<uint32_t>(4u, L->edge_label_align);
                label_off = (4u + L->edge_label_align - 1u) / L->edge_label_align * L->edge_label_align;
                edge_size = (label_off + L->edge_label_size + a - 1u) / a * a;
            }
            for (int32_t e = 0; e < h->nnz; ++e)
                for (uint32_t f = 0; f < L->n_edge_ptr; ++f)
                    *reinterpret_cast<uint64_t *>(dst + h->off_edge + (size_t)e * edge_size + label_off + L->edge_ptr_offset[f]) +=
                        dev_base;
        }
        table[k].blob = dev_base;
        table[k].bytes = (uint32_t)blob_bytes[k];
        table[k].n_node = (uint32_t)h->n_node;
        gs->n_node.push_back((uint32_t)h->n_node);
        gs->blob_bytes.push_back((uint32_t)blob_bytes[k]);
        gs->nnz.push_back((uint32_t)h->nnz);
        auto top2 = [](uint32_t *m, uint32_t v) {
            if (v > m[0]) {
                m[1] = m[0];
                m[0] = v;
            } else if (v > m[1]) {
                m[1] = v;
            }
        };
        top2(gs->max_blob, (uint32_t)blob_bytes[k]);
        top2(gs->max_node, (uint32_t)h->n_node);
        top2(gs->max_nnz, (uint32_t)h->nnz);
        if (!(h->flags & 2u)) gs->index16 = false;
        gs->max_tc = std::max(gs->max_tc, h->max_tc);
        gs->max_degree = std::max(gs->max_degree, h->max_degree);
        gs->sum_node += (uint64_t)h->n_node;
        {
            const uint32_t *te = reinterpret_cast<const uint32_t *>(dst + h->off_tileelem);
            for (int32_t t = 0; t < h->n_tile; ++t) gs->max_tile_nnz = std::max(gs->max_tile_nnz, te[t + 1] - te[t]);
        }
        {
            const uint32_t *rowptr = reinterpret_cast<const uint32_t *>(dst + h->off_rowptr);
            const uint32_t *lanemap = reinterpret_cast<const uint32_t *>(dst + h->off_lanemap);
            for (int a = 0; a < 2; ++a)
                for (int w = 0; w < 4; ++w) {
                    const uint32_t adj = a ? 4u : 2u, VL = 32u * (w + 1);
                    uint32_t next = (uint32_t)h->n_node, ovf = 0;
                    for (int32_t pos = 0; pos < h->n_node; ++pos) {  // nodes by decreasing degree
                        const uint32_t c = lanemap[pos] & 0xffffu, deg = rowptr[c + 1] - rowptr[c];
                        const uint32_t want = deg > adj ? (deg - 1u) / adj : 0u;
                        const uint32_t got = next >= VL ? 0u : std::min(want, VL - next);
                        next += want;
                        if (deg > adj * (1u + got)) ovf += deg - adj * (1u + got);
                    }
                    gs->max_ovf[a][w] = std::max(gs->max_ovf[a][w], ovf);
                }
        }
        top2(gs->max_idx, (uint32_t)(((h->n_node + 1) * 4u + 15u) & ~15u) +
                              (uint32_t)((h->nnz * 4u + 15u) & ~15u) + (uint32_t)(h->off_emeta - h->off_edge));
        off += blob_bytes[k];
    }
    if (n == 1) {
        gs->max_blob[1] = gs->max_blob[0];
        gs->max_node[1] = gs->max_node[0];
        gs->max_nnz[1] = gs->max_nnz[0];
        gs->max_idx[1] = gs->max_idx[0];
    }
    *out = gs;
    return gdb_graphset_upload(gs);
}

extern "C" int gdb_graphset_upload(gdb_graphset_t gs) {
    if (!gs) return gdb_fail(GDB_ERR_INVALID, "null graph set");
    RT(cudaSetDevice(gs->ctx->device));
    RT(cudaMemcpyAsync(gs->dev, gs->host, gs->bytes, cudaMemcpyHostToDevice, gs->ctx->stream));
    RT(cudaStreamSynchronize(gs->ctx->stream));
    return GDB_OK;
}

extern "C" int gdb_graphset_bytes(gdb_graphset_t gs, uint64_t *bytes) {
    if (!gs || !bytes) return gdb_fail(GDB_ERR_INVALID, "null argument");
    *bytes = gs->bytes;
    return GDB_OK;
}

extern "C" int gdb_graphset_destroy(gdb_graphset_t gs) {
    if (!gs) return GDB_OK;
    if (gs->ctx && gs->ctx->norm_gs == gs) gs->ctx->norm_gs = nullptr;
    cudaSetDevice(gs->ctx->device);
    if (gs->dev) cudaFree(gs->dev);
    if (gs->host) cudaFreeHost(gs->host);
    delete gs;
    return GDB_OK;
}

// ---------------------------------------------------------------------------
// host thread pool (collection of finished column blocks)
// ---------------------------------------------------------------------------
struct gdb_pool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<void(size_t)> job;
    size_t n_parts = 0, next = 0, done = 0;
    uint64_t epoch = 0;
    bool stop = false;

    explicit gdb_pool(unsigned n) {
        for (unsigned k = 0; k < n; ++k) workers.emplace_back([this] { loop(); });
    }
    ~gdb_pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &w : workers) w.join();
    }
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            cv.wait(lk, [&] { return stop || (epoch != seen && next < n_parts); });
            if (stop) return;
            while (next < n_parts) {
                const size_t part = next++;
                lk.unlock();
                job(part);
                lk.lock();
                if (++done == n_parts) cv_done.notify_all();
            }
            seen = epoch;
        }
    }
    // run f(0..n-1) on the workers and the calling thread; returns when all are done
    void parallel_for(size_t n, std::function<void(size_t)> f) {
        if (n == 0) return;
        std::unique_lock<std::mutex> lk(mu);
        job = std::move(f);
        n_parts = n;
        next = done = 0;
        ++epoch;
        cv.notify_all();
        while (next < n_parts) {
            const size_t part = next++;
            lk.unlock();
            job(part);
            lk.lock();
            ++done;
        }
        cv_done.wait(lk, [&] { return done == n_parts; });
        n_parts = 0;
    }
};

static gdb_pool *ctx_pool(gdb_context_t c) {
    if (!c->pool) {
        unsigned n = std::thread::hardware_concurrency();
        if (const char *env = getenv("GDB_HOST_THREADS")) n = (unsigned)strtoul(env, nullptr, 10);
        n = std::max(1u, std::min(n, 16u));
        c->pool = new gdb_pool(n - 1);  // the calling thread takes part
    }
    return c->pool;
}

void gdb_pool_destroy(gdb_pool *p) { delete p; }

// ---------------------------------------------------------------------------
// solve
// ---------------------------------------------------------------------------
namespace {

struct LaunchCfg {
    CUfunction fn = nullptr;
    int kind = 0;  // 0 general, 1 small, 2 large (cluster)
    uint64_t grid = 0, smem = 0, scratch_stride = 0;
    uint32_t graphs_need = 0, cluster = 1, row_cap = 0;
    int block = 0;
};

struct Tile {
    uint32_t i0, i1, j0, j1;
    uint64_t n_jobs;
    uint64_t c0, c1;  // output columns complete once this launch (and all earlier ones) finished
};

uint64_t triu_jobs(uint64_t lo, uint64_t hi, uint64_t j1) {
    const uint64_t rows = hi - lo, m = j1 - lo;
    return rows * m - rows * (rows - 1) / 2;
}

}  // namespace

// compile and load the large-pair kernel the first time a graph set needs it (graph-level
// outputs only).  A failure is reported once; later solves fall back to the general kernel.
static int ensure_large(gdb_context_t c, gdb_program_t p) {
    if (p->nodal != GDB_NODAL_NONE) return GDB_OK;
    std::lock_guard<std::mutex> lk(p->large_mu);
    if (p->large_tried) return GDB_OK;
    p->large_tried = true;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin;
    std::string log;
    int rc = compile_cubin(p->source, p->extra.empty() ? nullptr : p->extra.c_str(), 4, cubin, log);
    if (rc) return rc;
    p->log += log;
    CUfunction fn = nullptr;
    DRV(c, c->cuModuleLoadData(&p->mod_large, cubin.data()));
    DRV(c, c->cuModuleGetFunction(&fn, p->mod_large, "mlgk_solve_large"));
    CUdeviceptr ll = 0;
    size_t ll_bytes = 0;
    unsigned vals[3] = {0, 0, 0};
    DRV(c, c->cuModuleGetGlobal(&ll, &ll_bytes, p->mod_large, "gdb_large_layout"));
    if (ll_bytes != sizeof vals) return gdb_fail(GDB_ERR_LAYOUT, "gdb_large_layout has %zu bytes", ll_bytes);
    RT(cudaMemcpy(vals, reinterpret_cast<void *>(ll), sizeof vals, cudaMemcpyDeviceToHost));
    p->ell_entry = vals[0];
    p->large_block = vals[1];
    p->large_ltr = vals[2];
    int v = 0;
    c->cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, fn);
    p->info.num_regs_large = v;
    c->cuFuncGetAttribute(&p->large_static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, fn);
    DRV(c, c->cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                 (int)c->prop.sharedMemPerBlockOptin - p->large_static_smem));
    p->fn_large = fn;
    p->info.compile_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return GDB_OK;
}

static int pick_kernel(gdb_context_t c, gdb_program_t p, gdb_graphset_t gs, const gdb_solve_args *a, LaunchCfg &k) {
    const int block = p->info.block_size;
    const int nvec = p->eval_gradient ? 6 : 5;
    const uint64_t maxN = (uint64_t)gs->max_node[0] * gs->max_node[0];
    const uint64_t maxNpad = (maxN + 3) & ~3ull;
    const uint64_t graphs_need = (uint64_t)gs->max_blob[0] + gs->max_blob[1];  // small kernel: whole blobs
    const uint64_t idx_need = (uint64_t)gs->max_idx[0] + gs->max_idx[1];       // general kernel: row index + edges
    const uint64_t full_need = idx_need + nvec * maxNpad * 4;
    uint64_t cap = (uint64_t)p->info.max_dynamic_smem;
    if (const char *env = getenv("GDB_SMEM_CAP")) {  // testing / tuning hook
        const uint64_t v = strtoull(env, nullptr, 10);
        cap = std::min<uint64_t>(cap, v & ~15ull);
    }
    uint64_t smem = 0;
    bool spill = true;
    if (full_need <= cap) {
        smem = full_need;
        spill = false;
    } else if (idx_need <= std::min<uint64_t>(cap, 72 * 1024)) {
        smem = idx_need;  // index staged, vectors in the global arena
    } else {
        smem = 0;
    }
    k.fn = p->fn;
    k.kind = 0;
    k.block = block;
    k.graphs_need = (uint32_t)graphs_need;
    // small-pair kernel: blobs + cached edge products + diag + 4 vectors (x2 with
    // gradients) of the largest possible pair must fit in shared memory
    {
        const uint64_t nrhs = p->eval_gradient ? 2 : 1;
        // work area of mlgk_solve_small (same layout, from the maxima of the graph set):
        // step table | row table | lane tables (vown, vhelp, vovf, wslot, vinfo) | p | W[k1][lanes' slots + overflow]
        const uint64_t pad4nnz = ((uint64_t)gs->max_nnz[0] + 3) & ~3ull;
        const uint64_t wrow = ((32ull * p->wpt * p->adj + 3) & ~3ull) +
                              (((uint64_t)gs->max_ovf[p->adj == 4][std::min(p->wpt, 4) - 1] + 3) & ~3ull);
        const uint64_t wmax = (uint64_t)gs->max_nnz[0] * wrow;
        const uint64_t tables = 32ull * p->wpt + 2 * (((uint64_t)gs->max_node[0] + 3) & ~3ull) + pad4nnz + 4;
        // two blob staging buffers (double-buffered TMA prefetch) + the work area
        const uint64_t small_need = 2 * graphs_need + (((uint64_t)gs->max_nnz[0] * 8 + 15) & ~15ull) +
                                    (((uint64_t)gs->max_node[0] * 8 + 15) & ~15ull) + tables * 4 +
                                    nrhs * maxNpad * 4 + std::max(nrhs * maxNpad, wmax) * 4;  // p; W (W p aliases it)
        const uint64_t small_cap = std::min<uint64_t>(cap, (uint64_t)c->prop.sharedMemPerBlockOptin - p->small_static_smem);
        // one warp per tile row of G1, lanes (x workers per thread) over the columns of G2
        const bool mapped = (uint64_t)gs->max_node[0] <= (uint64_t)p->rpw * (uint64_t)(block / 32) &&
                            (uint64_t)gs->max_node[0] <= 32ull * p->wpt;
        // nodal Jacobians (forward sensitivities): a second W-shaped buffer for dW, x by node
        // and the right-hand sides of a round
        const bool nodal_grad = p->eval_gradient && p->nodal != GDB_NODAL_NONE;
        const uint64_t ng_need = nodal_grad ? (((wmax + 3) & ~3ull) * 2 - wmax) * 4 + maxNpad * 4 + maxNpad * 8 + 64 : 0;
        if (gs->index16 && small_need + ng_need <= small_cap && mapped && wmax < (1u << 24) &&
            !getenv("GDB_FORCE_GENERAL")) {
            k.fn = p->fn_small;
            k.kind = 1;
            smem = small_need + ng_need;
            spill = false;
        }
    }
    (void)a;
    // large-pair kernel: one cluster per pair, when the vectors of the largest pair would
    // otherwise live in a per-CTA arena (graph-level outputs; 16-bit row index)
    const bool large_fits = spill && k.kind == 0 && p->nodal == GDB_NODAL_NONE && gs->index16 &&
                            !getenv("GDB_FORCE_GENERAL") && (uint64_t)gs->max_node[0] <= 32ull * p->lcpt;
    if (large_fits) {
        int rc = ensure_large(c, p);
        if (rc) return rc;
    }
    if (large_fits && p->fn_large) {
        const uint64_t n2p = ((uint64_t)gs->max_node[0] + 3) & ~3ull;
        const uint64_t D = std::min<uint64_t>(gs->max_degree, (uint64_t)p->lell);
        const uint64_t ell_entry = p->ell_entry;  // sizeof(gdb_ell_t), read back from the module
        // elements of one CTA's rows of the first graph: at most its share of the tile rows
        const uint64_t tiles_per_cta = (((uint64_t)gs->max_node[0] + 7) / 8 + p->cluster - 1) / p->cluster;
        const uint64_t row_cap = std::min<uint64_t>((uint64_t)gs->max_tile_nnz * tiles_per_cta, 6144);
        const uint64_t ell = ((D * n2p * ell_entry + 15) & ~15ull) + ((n2p * 4 + 15) & ~15ull) +
                             ((row_cap * ell_entry + 15) & ~15ull);
        const uint64_t buf = (uint64_t)p->large_ltr * gs->max_tc * n2p * 4;
        const uint64_t lcap = std::min<uint64_t>(cap, (uint64_t)c->prop.sharedMemPerBlockOptin - p->large_static_smem);
        // staging: ONE buffer.  Two resident CTAs of different pairs per SM hide the staging
        // latency of each other, and the shared memory a second buffer would take is worth
        // more as L1 (the node-kernel set-up re-reads the feature pools of both graphs; measured
        // on C4: 2 x 115 KB double buffered 32.5 k pairs/s, 2 x 80 KB single buffered 42.9 k).
        // GDB_LARGE_DOUBLE=1 forces double buffering (A-B hook).
        uint64_t need = ell + buf;
        int ctas_dbl = 0, ctas_sgl = 0;
        if (need <= lcap) DRV(c, c->cuOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_sgl, p->fn_large, (int)p->large_block, (size_t)need));
        if (getenv("GDB_LARGE_DOUBLE") && ell + 2 * buf <= lcap) {
            need = ell + 2 * buf;
            DRV(c, c->cuOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_dbl, p->fn_large, (int)p->large_block, (size_t)need));
        }
        if (need <= lcap) {
            // clusters in flight: what the device can co-schedule (an optional cap keeps
            // the vectors of all of them L2-resident: GDB_L2_BUDGET_MB)
            const int nvecs = p->eval_gradient ? 6 : 5;
            const uint64_t per_cluster = (uint64_t)nvecs * gs->max_node[0] * n2p * 4;
            const int ctas_per_sm = std::max(ctas_sgl > ctas_dbl ? ctas_sgl : ctas_dbl, 1);
            uint64_t clusters = (uint64_t)c->prop.multiProcessorCount * ctas_per_sm / p->cluster;
            const double mean_n = (double)gs->sum_node / gs->n;
            const double typical = nvecs * mean_n * mean_n * 4.0 * 1.15;
            double budget = typical * 1e6;  // no cap unless asked for
            if (const char *env = getenv("GDB_L2_BUDGET_MB")) budget = atof(env) * 1048576.0;
            // the arena must fit in a quarter of the device memory
            clusters = std::min<uint64_t>(clusters, std::max<uint64_t>(1, c->prop.totalGlobalMem / 4 / per_cluster));
            clusters = std::max<uint64_t>(1, std::min<uint64_t>(clusters, (uint64_t)(budget / typical)));
            if (const char *env = getenv("GDB_LARGE_CLUSTERS")) clusters = std::max(1, atoi(env));
            int rc;
            if ((rc = dev_reserve(c->scratch, clusters * per_cluster))) return rc;
            k.fn = p->fn_large;
            k.kind = 2;
            k.block = (int)p->large_block;
            k.row_cap = (uint32_t)row_cap;
            k.cluster = (uint32_t)p->cluster;
            k.grid = clusters * p->cluster;
            k.smem = need;
            k.scratch_stride = per_cluster / 4;
            return GDB_OK;
        }
    }
    int blocks_per_sm = 0;
    DRV(c, c->cuOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k.fn, block, (size_t)smem));
    uint64_t grid = (uint64_t)c->prop.multiProcessorCount * blocks_per_sm;
    if (spill) {
        // bound the arena: at most ~1/4 of device memory
        const uint64_t per_cta = nvec * maxNpad * 4;
        const uint64_t budget = c->prop.totalGlobalMem / 4;
        grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, budget / std::max<uint64_t>(per_cta, 1)));
        int rc;
        if ((rc = dev_reserve(c->scratch, grid * per_cta))) return rc;
        k.scratch_stride = nvec * maxNpad;
    }
    k.grid = grid;
    k.smem = smem;
    return GDB_OK;
}

// Every output position a job set can touch must lie inside the nX x nY
// output: a wrong starts / row0 / nX from an ABI caller is an error here, not
// an out-of-bounds device write.
static int check_extents(gdb_program_t p, gdb_graphset_t gs, const gdb_solve_args *a, uint64_t n_jobs) {
    auto ext = [&](uint32_t g) -> uint64_t {
        if (p->nodal == GDB_NODAL_FULL) return gs->n_node[g];
        if (p->nodal == GDB_NODAL_BLOCK) return (uint64_t)gs->n_node[g] * gs->n_node[g];
        return 1;
    };
    auto row_ok = [&](uint32_t g) { return (uint64_t)a->starts[g] >= a->row0 && (uint64_t)a->starts[g] - a->row0 + ext(g) <= a->nX; };
    auto col_ok = [&](uint32_t g) { return (uint64_t)a->starts[g] >= a->col0 && (uint64_t)a->starts[g] - a->col0 + ext(g) <= a->nY; };
    const bool cols = !p->diagonal && p->nodal != GDB_NODAL_BLOCK;
    if (a->job_mode == GDB_JOBS_LIST) {
        for (uint64_t k = 0; k < n_jobs; ++k) {
            const uint32_t i = a->jobs[2 * k], j = a->jobs[2 * k + 1];
            if (!row_ok(i) || (cols && !col_ok(j)) || (cols && p->symmetric && (!row_ok(j) || !col_ok(i))))
                return gdb_fail(GDB_ERR_INVALID, "job %llu (%u, %u) writes outside the %u x %u output", (unsigned long long)k, i, j, a->nX, a->nY);
        }
        return GDB_OK;
    }
    const uint32_t jlo = a->job_mode == GDB_JOBS_TRIU ? a->i0 : a->j0;
    for (uint32_t i = a->i0; i < a->i1; ++i)
        if (!row_ok(i) || (cols && p->symmetric && !col_ok(i)))
            return gdb_fail(GDB_ERR_INVALID, "graph %u: starts[] places its row outside the %u x %u output", i, a->nX, a->nY);
    if (cols)
        for (uint32_t j = jlo; j < a->j1; ++j)
            if (!col_ok(j) || (p->symmetric && !row_ok(j)))
                return gdb_fail(GDB_ERR_INVALID, "graph %u: starts[] places its column outside the %u x %u output", j, a->nX, a->nY);
    return GDB_OK;
}

extern "C" int gdb_solve(gdb_context_t c, gdb_program_t p, gdb_graphset_t gs, gdb_solve_args *a) {
    if (!c || !p || !gs || !a) return gdb_fail(GDB_ERR_INVALID, "gdb_solve: null argument");
    if (p->ctx != c || gs->ctx != c) return gdb_fail(GDB_ERR_INVALID, "program / graph set belong to another context");
    if (!a->starts) return gdb_fail(GDB_ERR_INVALID, "gdb_solve: starts is required");
    const bool own_dev = a->gramian_dev != nullptr;
    const bool to_host = !a->keep_on_device && (a->gramian != nullptr || !own_dev);
    // host output buffers are only needed when the results are copied back
    if (to_host && !a->gramian) return gdb_fail(GDB_ERR_INVALID, "gdb_solve: gramian buffer required");
    if (to_host && p->eval_gradient && !a->gradient)
        return gdb_fail(GDB_ERR_INVALID, "program evaluates gradients: gradient buffer required");
    if (own_dev && p->eval_gradient && !a->gradient_dev)
        return gdb_fail(GDB_ERR_INVALID, "program evaluates gradients: gradient_dev required next to gramian_dev");
    if (p->eval_gradient && a->nJ != p->layout[7])
        return gdb_fail(GDB_ERR_INVALID, "nJ = %u but the program has %u hyper-parameters", a->nJ, p->layout[7]);
    if (a->out_dtype != GDB_OUT_NONE) {
        if (a->out_dtype != GDB_OUT_F64 && a->out_dtype != GDB_OUT_F32) return gdb_fail(GDB_ERR_INVALID, "unknown out_dtype %d", a->out_dtype);
        if (!to_host || !a->out_gram || (p->eval_gradient && !a->out_grad))
            return gdb_fail(GDB_ERR_INVALID, "out_dtype needs host staging (gramian / gradient) and out_gram / out_grad");
        if (a->async) return gdb_fail(GDB_ERR_INVALID, "async solves cannot collect (out_dtype)");
    }
    uint64_t n_jobs = 0;
    if (a->job_mode == GDB_JOBS_LIST) {
        if (!a->jobs && a->n_jobs) return gdb_fail(GDB_ERR_INVALID, "job list missing");
        n_jobs = a->n_jobs;
        for (uint64_t k = 0; k < 2 * n_jobs; ++k)
            if (a->jobs[k] >= gs->n) return gdb_fail(GDB_ERR_INVALID, "job %llu references graph %u of %u", (unsigned long long)(k / 2), a->jobs[k], gs->n);
    } else if (a->job_mode == GDB_JOBS_RECT) {
        if (a->i1 > gs->n || a->j1 > gs->n || a->i0 > a->i1 || a->j0 > a->j1) return gdb_fail(GDB_ERR_INVALID, "job rectangle out of range");
        n_jobs = (uint64_t)(a->i1 - a->i0) * (a->j1 - a->j0);
    } else if (a->job_mode == GDB_JOBS_TRIU) {
        if (a->i1 > gs->n || a->j1 > gs->n || a->i0 > a->i1 || a->i1 > a->j1) return gdb_fail(GDB_ERR_INVALID, "job triangle out of range");
        n_jobs = triu_jobs(a->i0, a->i1, a->j1);
    } else {
        return gdb_fail(GDB_ERR_INVALID, "unknown job_mode %d", a->job_mode);
    }
    if (a->n_starts < gs->n) return gdb_fail(GDB_ERR_INVALID, "starts has %u entries for %u graphs", a->n_starts, gs->n);
    a->kernel_ms = a->h2d_ms = a->d2h_ms = 0.f;
    a->cg_iterations = a->matvec_products = a->vector_elements = 0;
    a->h2d_bytes = a->d2h_bytes = 0;
    a->n_launches = 0;
    if (n_jobs == 0) return GDB_OK;
    if (a->store_diag && (a->nX != gs->n || a->nY != 1 || n_jobs != gs->n))
        return gdb_fail(GDB_ERR_INVALID, "store_diag needs one (i, i) job per graph and nX = number of graphs");
    if (a->normalize) {
        if (p->nodal) return gdb_fail(GDB_ERR_INVALID, "fused normalization is defined for graph-level outputs only");
        if (c->norm_gs != gs || c->norm_n != gs->n) return gdb_fail(GDB_ERR_INVALID, "normalize: no self-similarities stored for this graph set");
        if (p->eval_gradient && c->norm_nj != a->nJ) return gdb_fail(GDB_ERR_INVALID, "normalize: stored self-similarities carry no matching Jacobian");
    }
    int rc;
    if ((rc = check_extents(p, gs, a, n_jobs))) return rc;

    RT(cudaSetDevice(c->device));
    cudaStream_t st = a->stream ? static_cast<cudaStream_t>(a->stream) : c->stream;
    const uint64_t plane = (uint64_t)a->nX * a->nY;
    const uint64_t grad_floats = p->eval_gradient ? plane * a->nJ : 0;
    float *d_gram = a->gramian_dev, *d_grad = a->gradient_dev;
    if (!own_dev) {
        if ((rc = dev_reserve(c->gram, plane * 4))) return rc;
        if (grad_floats && (rc = dev_reserve(c->grad, grad_floats * 4))) return rc;
        d_gram = static_cast<float *>(c->gram.ptr);
        d_grad = static_cast<float *>(c->grad.ptr);
    }
    if ((rc = dev_reserve(c->starts, (size_t)a->n_starts * 4))) return rc;
    if (a->job_mode == GDB_JOBS_LIST && (rc = dev_reserve(c->jobs, n_jobs * 8))) return rc;

    LaunchCfg cfg;
    if ((rc = pick_kernel(c, p, gs, a, cfg))) return rc;
    a->used_small_kernel = cfg.kind;

    // ---- launch plan: one launch, or one per block of rows / columns -------------
    std::vector<Tile> tiles;
    auto col_at = [&](uint32_t g) -> uint64_t {  // first output column of graph g
        if (g >= a->n_starts) return a->nY;
        const uint64_t s = (uint64_t)a->starts[g] - a->col0;
        return std::min<uint64_t>(s, a->nY);
    };
    const bool can_tile = a->tile > 0 && !p->diagonal && p->nodal != GDB_NODAL_BLOCK &&
                          ((a->job_mode == GDB_JOBS_TRIU && p->symmetric && a->i1 == a->j1) || a->job_mode == GDB_JOBS_RECT);
    if (!can_tile) {
        // one launch.  A rectangle completes the columns of its graphs j0..j1 only
        // (a tile of a larger device-resident matrix is copied back alone)
        const bool rect = a->job_mode == GDB_JOBS_RECT && !p->diagonal && p->nodal != GDB_NODAL_BLOCK;
        tiles.push_back({a->i0, a->i1, a->j0, a->j1, n_jobs, rect ? col_at(a->j0) : 0, rect ? col_at(a->j1) : a->nY});
    } else if (a->job_mode == GDB_JOBS_TRIU) {
        double width = a->tile;
        for (uint32_t lo = a->i0; lo < a->i1;) {
            const uint32_t hi = std::min<uint64_t>(a->i1, (uint64_t)lo + std::max<uint32_t>(1u, (uint32_t)width));
            tiles.push_back({lo, hi, lo, a->j1, triu_jobs(lo, hi, a->j1), col_at(lo), col_at(hi)});
            lo = hi;
            if (a->tile_shrink > 0.f && a->tile_shrink < 1.f) width = std::max(64.0, width * a->tile_shrink);
        }
    } else {
        double width = a->tile;
        for (uint32_t lo = a->j0; lo < a->j1;) {
            const uint32_t hi = std::min<uint64_t>(a->j1, (uint64_t)lo + std::max<uint32_t>(1u, (uint32_t)width));
            tiles.push_back({a->i0, a->i1, lo, hi, (uint64_t)(a->i1 - a->i0) * (hi - lo), col_at(lo), col_at(hi)});
            lo = hi;
            if (a->tile_shrink > 0.f && a->tile_shrink < 1.f) width = std::max(64.0, width * a->tile_shrink);
        }
    }
    const size_t nt = tiles.size();
    if ((rc = dev_reserve(c->counters, nt * 32))) return rc;
    // events: [0] start, [1] inputs done, [2] kernels done, [3] copies done + per tile (kernel done, copy done)
    while (c->tile_ev.size() < 2 * nt) {
        cudaEvent_t e;
        RT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->tile_ev.push_back(e);
    }

    // ---- inputs -----------------------------------------------------------------
    RT(cudaEventRecord(c->ev[0], st));
    if (!own_dev) RT(cudaStreamWaitEvent(st, c->ev_copy, 0));  // earlier async copies out of the context's buffers are done
    if (a->upload_graphs) {
        RT(cudaMemcpyAsync(gs->dev, gs->host, gs->bytes, cudaMemcpyHostToDevice, st));
        a->h2d_bytes += gs->bytes;
    }
    if (a->job_mode == GDB_JOBS_LIST) {
        RT(cudaMemcpyAsync(c->jobs.ptr, a->jobs, n_jobs * 8, cudaMemcpyHostToDevice, st));
        a->h2d_bytes += n_jobs * 8;
    }
    RT(cudaMemcpyAsync(c->starts.ptr, a->starts, (size_t)a->n_starts * 4, cudaMemcpyHostToDevice, st));
    a->h2d_bytes += (uint64_t)a->n_starts * 4;
    RT(cudaMemsetAsync(c->counters.ptr, 0, nt * 32, st));
    if (!own_dev) {
        RT(cudaMemsetAsync(d_gram, 0, plane * 4, st));
        if (grad_floats) RT(cudaMemsetAsync(d_grad, 0, grad_floats * 4, st));
    }

    std::vector<unsigned char> params(p->layout[0], 0);
    gdb_params_fixed_host f{};
    f.graphs = reinterpret_cast<uint64_t>(gs->dev);
    f.jobs = reinterpret_cast<uint64_t>(c->jobs.ptr);
    f.starts = reinterpret_cast<uint64_t>(c->starts.ptr);
    f.gram = reinterpret_cast<uint64_t>(d_gram);
    f.grad = reinterpret_cast<uint64_t>(d_grad);
    f.scratch = reinterpret_cast<uint64_t>(c->scratch.ptr);
    f.scratch_stride = cfg.scratch_stride;
    f.job_mode = (uint32_t)a->job_mode;
    f.nX = a->nX, f.nY = a->nY, f.nJ = a->nJ;
    f.q = a->q, f.eps = a->eps, f.ftol = a->ftol, f.gtol = a->gtol;
    f.smem_bytes = (uint32_t)cfg.smem;
    f.row0 = a->row0, f.col0 = a->col0;
    f.blob_slot = cfg.graphs_need;
    f.pad3 = cfg.row_cap;
    if (a->normalize) {
        f.norm_n = c->norm_n;
        f.norm_diag = reinterpret_cast<uint64_t>(c->norm_diag.ptr);
        f.norm_ddiag = reinterpret_cast<uint64_t>(c->norm_ddiag.ptr);
    }
    const void *thetas[3] = {a->node_theta, a->edge_theta, a->p_theta};
    for (int k = 0; k < 3; ++k) {
        if (!p->theta_size[k]) continue;
        if (!thetas[k]) return gdb_fail(GDB_ERR_INVALID, "hyper-parameter block %d missing", k);
        if (p->layout[1 + k] + p->theta_size[k] > p->layout[0]) return gdb_fail(GDB_ERR_LAYOUT, "theta block %d overflows the parameter struct", k);
        memcpy(params.data() + p->layout[1 + k], thetas[k], p->theta_size[k]);
    }
    void *kargs[1] = {params.data()};

    // planes copied back: all of them, or only the active ones when collecting
    std::vector<uint32_t> planes;  // Jacobian planes to copy / collect, in output order
    for (uint32_t k = 0; k < (p->eval_gradient ? a->nJ : 0u); ++k)
        if (a->out_dtype == GDB_OUT_NONE || !a->plane_mask || a->plane_mask[k]) planes.push_back(k);

    RT(cudaEventRecord(c->ev[1], st));
    for (size_t t = 0; t < nt; ++t) {
        const Tile &T = tiles[t];
        f.counters = reinterpret_cast<uint64_t>(c->counters.ptr) + t * 32;
        f.n_jobs = T.n_jobs;
        f.i0 = T.i0, f.i1 = T.i1, f.j0 = T.j0, f.j1 = T.j1;
        memcpy(params.data(), &f, sizeof f);
        uint64_t grid = std::min<uint64_t>(cfg.grid, T.n_jobs);
        if (cfg.kind == 2) grid = std::min<uint64_t>(cfg.grid, T.n_jobs * cfg.cluster);  // whole clusters
        DRV(c, c->cuLaunchKernel(cfg.fn, (unsigned)grid, 1, 1, (unsigned)cfg.block, 1, 1, (unsigned)cfg.smem, (CUstream)st, kargs, nullptr));
        a->n_launches++;
        a->grid = (uint32_t)grid;
        if (to_host && T.c1 > T.c0) {
            // the finished column block leaves on the copy stream while the next launch runs
            RT(cudaEventRecord(c->tile_ev[2 * t], st));
            RT(cudaStreamWaitEvent(c->copy_stream, c->tile_ev[2 * t], 0));
            const uint64_t off = T.c0 * a->nX, cnt = (T.c1 - T.c0) * a->nX;
            RT(cudaMemcpyAsync(a->gramian + off, d_gram + off, cnt * 4, cudaMemcpyDeviceToHost, c->copy_stream));
            for (uint32_t k : planes)
                RT(cudaMemcpyAsync(a->gradient + k * plane + off, d_grad + k * plane + off, cnt * 4, cudaMemcpyDeviceToHost, c->copy_stream));
            a->d2h_bytes += cnt * 4 * (1 + planes.size());
            RT(cudaEventRecord(c->tile_ev[2 * t + 1], c->copy_stream));
        }
    }
    a->smem_bytes = (uint32_t)cfg.smem;
    RT(cudaEventRecord(c->ev[2], st));
    if (a->store_diag) {
        if ((rc = dev_reserve(c->norm_diag, plane * 4))) return rc;
        RT(cudaMemcpyAsync(c->norm_diag.ptr, d_gram, plane * 4, cudaMemcpyDeviceToDevice, st));
        c->norm_nj = 0;
        if (grad_floats) {
            if ((rc = dev_reserve(c->norm_ddiag, grad_floats * 4))) return rc;
            RT(cudaMemcpyAsync(c->norm_ddiag.ptr, d_grad, grad_floats * 4, cudaMemcpyDeviceToDevice, st));
            c->norm_nj = a->nJ;
        }
        c->norm_n = gs->n;
        c->norm_gs = gs;
    }
    if (to_host) RT(cudaEventRecord(c->ev_copy, c->copy_stream));
    if (a->async) return GDB_OK;

    const bool trace = getenv("GDB_TRACE") != nullptr;
    const auto t_enq = std::chrono::steady_clock::now();
    auto since = [&](const char *what, size_t t) {
        if (trace)
            fprintf(stderr, "[gdb_solve] %-22s tile %zu  +%.2f ms\n", what, t,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enq).count());
    };
    since("enqueued", nt);
    std::vector<unsigned long long> counters(nt * 4, 0);
    // ---- collection: convert every finished column block on the host threads ----
    if (to_host && a->out_dtype != GDB_OUT_NONE) {
        gdb_pool *pool = ctx_pool(c);
        {
            // The result arrays are usually fresh allocations: fault their pages in NOW, on all
            // host threads, while the first launch computes -- the conversion passes below then
            // run at memory bandwidth instead of at page-fault speed (192 MB of float64 for C3:
            // 47 000 faults).  One store per page; nothing has been converted yet.
            const size_t esz = a->out_dtype == GDB_OUT_F64 ? 8 : 4, page = 4096;
            char *bufs[2] = {static_cast<char *>(a->out_gram), static_cast<char *>(a->out_grad)};
            const size_t lens[2] = {(size_t)plane * esz, a->out_grad ? (size_t)plane * planes.size() * esz : 0};
            for (int b = 0; b < 2; ++b) {
                if (!bufs[b] || !lens[b]) continue;
                char *base = bufs[b];
                const size_t n_pages = (lens[b] + page - 1) / page, per = 256;
                const size_t len = lens[b];
                pool->parallel_for((n_pages + per - 1) / per, [=](size_t part) {
                    for (size_t pg = part * per; pg < std::min(n_pages, (part + 1) * per); ++pg) {
                        volatile char *q = base + std::min(pg * page, len - 1);
                        *q = 0;
                    }
                });
            }
        }
        for (size_t t = 0; t < nt; ++t) {
            const Tile &T = tiles[t];
            if (T.c1 <= T.c0) continue;
            if (t == 0) since("pages touched", t);
            RT(cudaEventSynchronize(c->tile_ev[2 * t + 1]));
            since("block on host", t);
            const uint64_t off = T.c0 * a->nX, cnt = (T.c1 - T.c0) * a->nX;
            // one parallel pass over (planes x chunks) of this column block
            const size_t chunk = 1u << 15, per_plane = (size_t)((cnt + chunk - 1) / chunk);
            const int dt = a->out_dtype;
            const float *g_src = a->gramian, *d_src = a->gradient;
            void *g_dst = a->out_gram, *d_dst = a->out_grad;
            const uint32_t *pl = planes.data();
            pool->parallel_for(per_plane * (planes.size() + 1), [=](size_t part) {
                const size_t m = part / per_plane, ch = part % per_plane;
                const uint64_t lo = ch * chunk, hi = std::min<uint64_t>(cnt, lo + chunk);
                const float *src = m == 0 ? g_src + off : d_src + pl[m - 1] * plane + off;
                const uint64_t dst_off = m == 0 ? off : (m - 1) * plane + off;
                if (dt == GDB_OUT_F64) {
                    double *dst = static_cast<double *>(m == 0 ? g_dst : d_dst) + dst_off;
                    for (uint64_t k = lo; k < hi; ++k) dst[k] = (double)src[k];
                } else {
                    float *dst = static_cast<float *>(m == 0 ? g_dst : d_dst) + dst_off;
                    memcpy(dst + lo, src + lo, (hi - lo) * 4);
                }
            });
        }
    }
    since("collected", nt);
    // (the diagnostics counters come back last: a copy into pageable memory blocks the host
    // until the stream gets there, which would serialise the collection behind the kernels)
    RT(cudaMemcpyAsync(counters.data(), c->counters.ptr, nt * 32, cudaMemcpyDeviceToHost, st));
    RT(cudaStreamSynchronize(st));
    if (to_host) RT(cudaStreamSynchronize(c->copy_stream));
    since("synchronized", nt);
    RT(cudaGetLastError());
    cudaEventElapsedTime(&a->h2d_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&a->kernel_ms, c->ev[1], c->ev[2]);
    for (size_t t = 0; t < nt; ++t) {
        a->cg_iterations += counters[4 * t + 1];
        a->matvec_products += counters[4 * t + 2];
        a->vector_elements += counters[4 * t + 3];
    }
    return GDB_OK;
}

extern "C" int gdb_last_outputs(gdb_context_t c, void **g, void **d) {
    if (!c) return gdb_fail(GDB_ERR_INVALID, "null context");
    if (g) *g = c->gram.ptr;
    if (d) *d = c->grad.ptr;
    return GDB_OK;
}
