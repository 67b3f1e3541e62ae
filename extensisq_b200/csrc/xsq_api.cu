// xsq_api.cu -- the C ABI of libxsq.so (see include/xsq.h).  Host-side
// argument validation mirrors the reference's constructors:
//   validate_tol            extensisq/common.py:30-54
//   _init_sc_control        extensisq/common.py:166-185
//   validate_first_step /   scipy/integrate/_ivp/common.py:10-23 (third party)
//   validate_max_step
// There is NO CPU fallback: every entry point either runs on the CUDA device
// or returns XSQ_ERR_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "xsq.h"
#include "xsq_launch.h"
#include "xsq_user.h"
#include "xsq_comm.h"

namespace xsq {

int rkc_solve(const xsq_rkc_args_t* A, Comm* comm, cudaStream_t st);
int rkc_stage_bench(int nx, int rows, int reps, double* ms, cudaStream_t st);
int rkc_stage_bench_tma(int nx, int rows, int reps, double* ms, double* max_abs_diff, cudaStream_t st);

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

thread_local std::string g_detail;
void set_detail(const std::string& s) { g_detail = s; }

}  // namespace xsq
#include "xsq_params.h"
namespace xsq {

static int cuda_fail(cudaError_t e, const char* what) {
    g_detail = std::string(what) + ": " + cudaGetErrorString(e);
    return XSQ_ERR_CUDA;
}
#define XSQ_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);    \
    } while (0)

// ---- device read-back of a built-in tableau ---------------------------------
template <class T>
__global__ void tableau_dump(xsq_tableau_t* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out->n_stages = T::S;
    out->order = T::ORDER;
    out->order_secondary = T::ORDER2;
    out->n_poly = T::NPOL;
    for (int i = 0; i < T::S; ++i) {
        for (int j = 0; j < T::S; ++j) out->A[i][j] = T::a(i, j);
        out->B[i] = T::b(i);
        out->C[i] = T::c(i);
    }
    for (int i = 0; i <= T::S; ++i) {
        out->E[i] = T::e(i);
        for (int k = 0; k < T::NPOL; ++k) out->P[i][k] = T::p(i, k);
    }
    out->sc_params[0] = T::SC_KB1;
    out->sc_params[1] = T::SC_KB2;
    out->sc_params[2] = T::SC_A;
    out->sc_params[3] = T::SC_G;
    out->stbrad = T::STBRAD;
    out->tanang = T::TANANG;
}

// ---- fp64 FMA peak microbenchmark -------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters,
                                                        double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b);
            x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b);
            x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] =
        ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

static int launch_method(int method, int rhs, const RkDev& P, cudaStream_t st,
                         LaunchInfo* info) {
    switch (method) {
        case XSQ_TS5: return launch_Ts5(rhs, P, st, info);
        case XSQ_BS5: return launch_BS5(rhs, P, st, info);
        case XSQ_CK5: return launch_CK5(rhs, P, st, info);
        case XSQ_ME4: return launch_Me4(rhs, P, st, info);
        case XSQ_PR7: return launch_Pr7(rhs, P, st, info);
        case XSQ_PR8: return launch_Pr8(rhs, P, st, info);
        case XSQ_PR9: return launch_Pr9(rhs, P, st, info);
        case XSQ_CFMR7OSC: return launch_CFMR7osc(rhs, P, st, info);
        case XSQ_CKDISC: return launch_CKdisc(rhs, P, st, info);
        case XSQ_FI4N: return launch_Fi4N(rhs, P, st, info);
        case XSQ_FI5N: return launch_Fi5N(rhs, P, st, info);
        case XSQ_MU5NMB: return launch_Mu5Nmb(rhs, P, st, info);
        case XSQ_MR6NN: return launch_MR6NN(rhs, P, st, info);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

// init kernel -> persistent kernel -> (stiffness diagnosis on) probe queue kernel
// Optional device timing of the three kernels of a solve (xsq_profile_enable):
// CUDA events on the launching stream around ens_init / the persistent kernel /
// stiff_queue.  Used by bench.py for the roofline of the dominant kernel.
static std::atomic<int> g_profile{0};
static constexpr int kProfRing = 8;             // the last 8 profiled solves
static cudaEvent_t g_prof_ev[kProfRing][4] = {};
static long long g_prof_count = 0;
void prof_mark(int i, cudaStream_t st) {
    if (!g_profile.load(std::memory_order_relaxed)) return;
    cudaEvent_t* ev = g_prof_ev[g_prof_count % kProfRing];
    if (!ev[i]) cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[i], st);
    if (i == 3) ++g_prof_count;
}

static int dispatch(int method, int rhs, int events, const RkDev& P, const MethodInfo& mi,
                    cudaStream_t st, LaunchInfo* info) {
    // user right-hand sides and event functions are device code compiled at
    // run time into their own kernel
    if (rhs >= XSQ_RHS_USER_BASE || events != 0)
        return user_rk_launch(method, rhs, events, P, mi.s, mi.stbrad, mi.tanang, st);
    prof_mark(0, st);
    int rc = launch_ens_init(rhs, P, st);       // f0 + h_start for all lanes
    if (rc != XSQ_OK) return rc;
    prof_mark(1, st);
    if (method == XSQ_METHOD_SWAG) {
        rc = launch_swag(rhs, P, st);
        prof_mark(2, st);
        prof_mark(3, st);
        return rc;
    }
    rc = launch_method(method, rhs, P, st, info);
    prof_mark(2, st);
    if (rc == XSQ_OK && P.stiff_q_cap > 0)
        rc = launch_stiff_queue(rhs, P, mi.s, mi.stbrad, mi.tanang, st);
    prof_mark(3, st);
    return rc;
}

// Keep freed scratch cached in the stream-ordered pool: by default the pool
// returns memory to the OS at every synchronisation, and mapping tens of
// megabytes again on each call costs milliseconds.
void keep_pool_memory() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        // bounded: the scratch of repeated solves stays cached -- mapping the tens
        // of GB of a large event queue again costs hundreds of milliseconds per
        // solve (measured) -- but never more than 40 % of the device's memory;
        // xsq_trim_memory() returns it to the driver at any time
        size_t free_b = 0, total_b = 0;
        unsigned long long keep = 6ULL << 30;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b / 5 * 2 > keep)
            keep = total_b / 5 * 2;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

// Memory a new allocation of this library can draw on: what the driver reports
// free plus what the stream-ordered pool holds without using it.
static size_t available_bytes() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long reserved = 0, used = 0;
        if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess &&
            reserved > used)
            free_b += (size_t)(reserved - used);
    }
    return free_b;
}

static int solve_device(const xsq_rk_args_t* a, cudaStream_t st,
                        LaunchInfo* info) {
    keep_pool_memory();
    RkDev P;
    MethodInfo mi;
    std::vector<double> atol;
    int rc = build_params(a, &P, &mi, &atol);
    if (rc != XSQ_OK) return rc;
    if (a->n_lanes == 0) return XSQ_OK;
    // scratch: [work counter, probe-queue counter, event-queue counter (8 B each, padded
    //          to 32)] [atol vector] [init_h N] [init_f0 n_state x N] [init_nfev N]
    const size_t N = (size_t)a->n_lanes, ns = atol.size();
    const size_t off_h = (32 + ns * sizeof(double) + 15) & ~(size_t)15;
    const size_t off_f = off_h + N * sizeof(double);
    const size_t off_n = off_f + N * ns * sizeof(double);
    const size_t bytes = off_n + N * sizeof(int);
    char* scratch = nullptr;
    XSQ_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
    XSQ_CUDA(cudaMemsetAsync(scratch, 0, 32, st));
    XSQ_CUDA(cudaMemcpyAsync(scratch + 32, atol.data(),
                             atol.size() * sizeof(double),
                             cudaMemcpyHostToDevice, st));
    // the pageable atol copy is staged by the runtime before returning
    P.queue = (unsigned long long*)scratch;
    P.atol_dev = (const double*)(scratch + 32);
    P.init_h = (double*)(scratch + off_h);
    P.init_f0 = (double*)(scratch + off_f);
    P.init_nfev = (int*)(scratch + off_n);
    P.morder = (a->method == XSQ_METHOD_SWAG) ? 1 : mi.order2;
    // stiffness probes wait in two slots per resident thread (xsq_rk_core.cuh,
    // Lane::diagnose); 2048 threads per SM is the most any geometry launches
    double* slots = nullptr;
    P.stiff_slot = nullptr;
    P.stiff_threads = 0;
    P.stiff_q = nullptr;
    P.stiff_q_cap = 0;
    P.stiff_q_count = (unsigned long long*)(scratch + 8);
    if (P.nfev_stiff_detect > 0 && a->method != XSQ_METHOD_SWAG) {
        int dev = 0, n_sm = 0;
        XSQ_CUDA(cudaGetDevice(&dev));
        XSQ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        const bool wide = a->n_state > XSQ_MAX_LANE_STATE && a->rhs != XSQ_RHS_NBODY32;
        const bool warp_rhs = a->rhs == XSQ_RHS_NBODY32 || wide;
        const size_t nl = wide ? (size_t)(a->n_state + 31) / 32
                               : (warp_rhs ? 6 : (size_t)a->n_state);
        const size_t npl = a->rhs == XSQ_RHS_NBODY32 ? 2 : (size_t)(a->n_param > 0 ? a->n_param : 1);
        size_t threads = (size_t)n_sm * 2048;
        const size_t want = ((warp_rhs ? N * 32 : N) + 255) & ~(size_t)255;
        if (want < threads) threads = want;
        P.stiff_threads = (long long)threads;
        // ... and, for thread-per-system kernels, in a queue that a separate
        // kernel works off afterwards: up to 48 records per trajectory, at most
        // a quarter of the free memory and 8 GB.  When it is full the slots
        // take over.
        const size_t rec = 5 + 4 * nl + npl;
        size_t qcap = 0;
        if (!warp_rhs) {
            size_t free_b = 0, total_b = 0;
            XSQ_CUDA(cudaMemGetInfo(&free_b, &total_b));
            // 16 records per trajectory (a probe every nfev_stiff_detect
            // evaluations plus the rare `lotsfl` ones: 14 per lane on the
            // 10^4-step Lorenz benchmark), at most an eighth of the free memory
            // and 4 GB; whatever does not fit goes through the slots.  The
            // environment override (tests) can only shrink it.
            qcap = N * 16;
            size_t budget = free_b / 8;
            if (budget > ((size_t)4 << 30)) budget = (size_t)4 << 30;
            const size_t fit = budget / (rec * sizeof(double));
            if (fit < qcap) qcap = fit;
            if (const char* e = getenv("XSQ_STIFF_QUEUE_RECORDS")) {
                const long long want_q = atoll(e);
                if (want_q >= 0 && (size_t)want_q < qcap) qcap = (size_t)want_q;
            }
        }
        cudaError_t es = cudaMallocAsync(
            (void**)&slots, (2 * threads + qcap) * rec * sizeof(double), st);
        if (es != cudaSuccess) {
            cudaFreeAsync(scratch, st);
            return cuda_fail(es, "cudaMallocAsync");
        }
        P.stiff_slot = slots;
        P.stiff_q = slots + 2 * threads * rec;
        P.stiff_q_cap = (long long)qcap;
    }
    // Event queue (xsq_rk_core.cuh after_step / event_queue_body): steps with a
    // sign change of a non-terminal event wait here for their root solve.  One
    // record per located event: at most n_events x ev_capacity per trajectory (plus
    // one partly filled chunk per CTA), at most a third of the free memory; what
    // does not fit is solved in the lane.
    // Layout: [chunk counter, 16 B] [fill per chunk] [records].
    char* evq = nullptr;
    P.evq = nullptr;
    P.evq_cap = 0;
    P.evq_count = nullptr;
    P.evq_fill = nullptr;
    P.evq_exact = 0;
    {
        const bool wide = a->n_state > XSQ_MAX_LANE_STATE && a->rhs != XSQ_RHS_NBODY32;
        if (a->events != 0 && a->method != XSQ_METHOD_SWAG && a->rhs != XSQ_RHS_NBODY32 && !wide &&
            P.n_events > 0 && P.ev_capacity > 0) {
            int dev = 0, n_sm = 0;
            XSQ_CUDA(cudaGetDevice(&dev));
            XSQ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            const size_t fields = (5 + (size_t)(mi.s + 3) * (size_t)a->n_state + 1) & ~(size_t)1;
            const size_t free_b = available_bytes();
            size_t ctas = (size_t)n_sm * 16;                       // more than any launch uses
            if (ctas > (N + 127) / 128) ctas = (N + 127) / 128;
            const size_t need = N * (size_t)P.n_events * (size_t)P.ev_capacity + ctas * kEvqChunk;
            size_t qcap = need;
            const size_t fit = (free_b / 3) / (fields * sizeof(double));
            if (fit < qcap) qcap = fit;
            if (const char* e = getenv("XSQ_EVENT_QUEUE_RECORDS")) {   // tests: shrink or switch off
                const long long want_q = atoll(e);
                if (want_q >= 0 && (size_t)want_q < qcap) qcap = (size_t)want_q;
            }
            const size_t chunks = (qcap + kEvqChunk - 1) / kEvqChunk;
            if (chunks > 0) {
                const size_t head = (16 + chunks * sizeof(unsigned) + 255) & ~(size_t)255;
                cudaError_t es = cudaMallocAsync(
                    (void**)&evq, head + chunks * kEvqChunk * fields * sizeof(double), st);
                if (es != cudaSuccess) {
                    (void)cudaGetLastError();      // no queue: every root in the lane
                    evq = nullptr;
                } else if (cudaMemsetAsync(evq, 0, head, st) != cudaSuccess) {
                    (void)cudaGetLastError();
                    cudaFreeAsync(evq, st);
                    evq = nullptr;
                } else {
                    P.evq_count = (unsigned long long*)evq;
                    P.evq_fill = (unsigned*)(evq + 16);
                    P.evq = (double*)(evq + head);
                    P.evq_cap = (long long)(chunks * kEvqChunk);
                    P.evq_exact = chunks * kEvqChunk >= need ? 1 : 0;
                }
            }
        }
    }
    rc = dispatch(a->method, a->rhs, a->events, P, mi, st, info);
    if (evq) cudaFreeAsync(evq, st);
    if (slots) cudaFreeAsync(slots, st);
    cudaError_t e = cudaFreeAsync(scratch, st);
    if (rc == XSQ_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
    return rc;
}

}  // namespace xsq

using namespace xsq;

extern "C" {

int xsq_abi_version(void) { return XSQ_ABI_VERSION; }

const char* xsq_strerror(int err) {
    switch (err) {
        case XSQ_OK: return "ok";
        case XSQ_ERR_ARG: return "invalid argument";
        case XSQ_ERR_CUDA: return "CUDA error (no usable device or launch failure)";
        case XSQ_ERR_NVRTC: return "runtime compilation of the user RHS failed";
        case XSQ_ERR_UNSUPPORTED: return "method/rhs combination not available";
        case XSQ_ERR_NOMEM: return "out of memory";
        default: return "unknown error";
    }
}

const char* xsq_last_error_detail(void) { return g_detail.c_str(); }

int xsq_device_info(int device, int32_t* n_sm, int32_t* cc_major,
                    int32_t* cc_minor) {
    int n = 0;
    XSQ_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) { g_detail = "no such device"; return XSQ_ERR_CUDA; }
    int v = 0;
    XSQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (n_sm) *n_sm = v;
    XSQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    if (cc_major) *cc_major = v;
    XSQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    if (cc_minor) *cc_minor = v;
    return XSQ_OK;
}

int xsq_tableau_get(int32_t method, xsq_tableau_t* out) {
    if (!out) return XSQ_ERR_ARG;
    xsq_tableau_t* d = nullptr;
    XSQ_CUDA(cudaMalloc((void**)&d, sizeof(xsq_tableau_t)));
    cudaMemset(d, 0, sizeof(xsq_tableau_t));
    switch (method) {
        case XSQ_TS5: tableau_dump<tab::Ts5><<<1, 32>>>(d); break;
        case XSQ_BS5: tableau_dump<tab::BS5><<<1, 32>>>(d); break;
        case XSQ_CK5: tableau_dump<tab::CK5><<<1, 32>>>(d); break;
        case XSQ_ME4: tableau_dump<tab::Me4><<<1, 32>>>(d); break;
        case XSQ_PR7: tableau_dump<tab::Pr7><<<1, 32>>>(d); break;
        case XSQ_PR8: tableau_dump<tab::Pr8><<<1, 32>>>(d); break;
        case XSQ_PR9: tableau_dump<tab::Pr9><<<1, 32>>>(d); break;
        case XSQ_CFMR7OSC: tableau_dump<tab::CFMR7osc><<<1, 32>>>(d); break;
        case XSQ_CKDISC: tableau_dump<tab::CKdisc><<<1, 32>>>(d); break;
        default: cudaFree(d); return XSQ_ERR_ARG;
    }
    count_launch();
    cudaError_t e = cudaMemcpy(out, d, sizeof(xsq_tableau_t),
                               cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "tableau_dump");
    return XSQ_OK;
}

int xsq_rhs_builtin(const char* name, int32_t* rhs_out, int32_t* n_state,
                    int32_t* n_param) {
    if (!name) return XSQ_ERR_ARG;
    for (const RhsInfo& r : kBuiltinRhs) {
        if (std::strcmp(r.name, name) == 0) {
            if (rhs_out) *rhs_out = r.id;
            if (n_state) *n_state = r.n_state;
            if (n_param) *n_param = r.n_param;
            return XSQ_OK;
        }
    }
    g_detail = std::string("unknown built-in rhs: ") + name;
    return XSQ_ERR_ARG;
}

int xsq_rk_solve(const xsq_rk_args_t* args, void* stream) {
    return solve_device(args, (cudaStream_t)stream, nullptr);
}

int xsq_swag_solve(const xsq_swag_args_t* args, int32_t k_max, void* stream) {
    if (!args) return XSQ_ERR_ARG;
    if (k_max < 1 || k_max > 12) {          // shampine.py:102-103
        g_detail = "`k_max` should be an integer between 1 and 12.";
        return XSQ_ERR_ARG;
    }
    xsq_rk_args_t a = *args;
    a.method = XSQ_METHOD_SWAG;
    a.interpolant = XSQ_INTERP_FREE;
    a.use_sc_params = 0;
    a.h_forced = nullptr;
    a.n_forced = 0;
    a.reserved0 = k_max;
    return solve_device(&a, (cudaStream_t)stream, nullptr);
}

int xsq_rk_solve_host(const xsq_rk_args_t* h, int device) {
    if (!h || h->struct_size != (int32_t)sizeof(xsq_rk_args_t)) return XSQ_ERR_ARG;
    XSQ_CUDA(cudaSetDevice(device));
    int ns = h->n_state, np = h->n_param;
    const long long N = h->n_lanes;
    if (N < 0 || ns <= 0 || np < 0) return XSQ_ERR_ARG;
    cudaStream_t st;
    XSQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    xsq_rk_args_t d = *h;
    std::vector<void*> owned;
    int rc = XSQ_OK;
    auto dalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (bytes == 0) bytes = 8;
        if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { rc = XSQ_ERR_NOMEM; return nullptr; }
        owned.push_back(p);
        return p;
    };
    auto h2d = [&](const void* src, size_t bytes) -> void* {
        void* p = dalloc(bytes);
        if (p && src && bytes)
            if (cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
                rc = XSQ_ERR_CUDA;
        return p;
    };
    const size_t nd = sizeof(double), ni = sizeof(int32_t);
    d.y0 = (const double*)h2d(h->y0, (size_t)N * ns * nd);
    d.params = np ? (const double*)h2d(h->params, (size_t)N * np * nd) : nullptr;
    d.t_eval = h->n_eval ? (const double*)h2d(h->t_eval, (size_t)h->n_eval * nd) : nullptr;
    d.h_forced = h->n_forced ? (const double*)h2d(h->h_forced, (size_t)h->n_forced * nd) : nullptr;
    d.first_step_lanes = h->first_step_lanes ? (const double*)h2d(h->first_step_lanes, (size_t)N * nd) : nullptr;
    const size_t pitch = ((size_t)h->n_eval + 3) & ~(size_t)3;
    d.y_eval = h->n_eval ? (double*)dalloc((size_t)N * ns * pitch * nd) : nullptr;
    d.t_final = (double*)dalloc((size_t)N * nd);
    d.y_final = (double*)dalloc((size_t)N * ns * nd);
    d.h_next = h->h_next ? (double*)dalloc((size_t)N * nd) : nullptr;
    d.n_accepted = (int32_t*)dalloc((size_t)N * ni);
    d.n_rejected = (int32_t*)dalloc((size_t)N * ni);
    d.nfev = (int32_t*)dalloc((size_t)N * ni);
    d.status = (int32_t*)dalloc((size_t)N * ni);
    d.n_eval_done = h->n_eval_done ? (int32_t*)dalloc((size_t)N * ni) : nullptr;
    d.stiff_flags = h->stiff_flags ? (int32_t*)dalloc((size_t)N * ni) : nullptr;
    const size_t nev = h->events ? (size_t)N * (size_t)h->n_event_fns : 0;
    const size_t nrec = nev * (size_t)(h->ev_capacity > 0 ? h->ev_capacity : 0);
    if (h->events) {
        d.t_events = (double*)dalloc(nrec * nd);
        d.y_events = (double*)dalloc(nrec * ns * nd);
        d.ev_count = (int32_t*)dalloc(nev * ni);
        // records that are never written read back as NaN
        if (rc == XSQ_OK && d.t_events && d.y_events) {
            cudaMemsetAsync(d.t_events, 0xFF, nrec * nd, st);
            cudaMemsetAsync(d.y_events, 0xFF, nrec * ns * nd, st);
        }
    }
    if (rc == XSQ_OK) rc = solve_device(&d, st, nullptr);
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (rc == XSQ_OK && dst && bytes)
            if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = XSQ_ERR_CUDA;
    };
    if (h->n_eval && rc == XSQ_OK)            // strip the row padding
        if (cudaMemcpy2DAsync(h->y_eval, (size_t)h->n_eval * nd, d.y_eval, pitch * nd,
                              (size_t)h->n_eval * nd, (size_t)N * ns,
                              cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = XSQ_ERR_CUDA;
    d2h(h->t_final, d.t_final, (size_t)N * nd);
    d2h(h->y_final, d.y_final, (size_t)N * ns * nd);
    if (h->h_next) d2h(h->h_next, d.h_next, (size_t)N * nd);
    d2h(h->n_accepted, d.n_accepted, (size_t)N * ni);
    d2h(h->n_rejected, d.n_rejected, (size_t)N * ni);
    d2h(h->nfev, d.nfev, (size_t)N * ni);
    d2h(h->status, d.status, (size_t)N * ni);
    if (h->n_eval_done) d2h(h->n_eval_done, d.n_eval_done, (size_t)N * ni);
    if (h->stiff_flags) d2h(h->stiff_flags, d.stiff_flags, (size_t)N * ni);
    if (h->events) {
        d2h(h->t_events, d.t_events, nrec * nd);
        d2h(h->y_events, d.y_events, nrec * ns * nd);
        d2h(h->ev_count, d.ev_count, nev * ni);
    }
    for (void* p : owned) cudaFreeAsync(p, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (rc == XSQ_OK && e != cudaSuccess) return cuda_fail(e, "xsq_rk_solve_host");
    return rc;
}

int xsq_comm_unique_id(char id[128]) { return comm_unique_id(id); }
int xsq_comm_create(int32_t rank, int32_t world, const char id[128], void** comm) {
    if (!comm || world < 1 || rank < 0 || rank >= world) return XSQ_ERR_ARG;
    Comm* c = nullptr;
    int rc = comm_create(rank, world, id, &c);
    *comm = c;
    return rc;
}
int xsq_comm_destroy(void* comm) {
    comm_destroy((Comm*)comm);
    return XSQ_OK;
}
int xsq_rkc_solve(const xsq_rkc_args_t* args, void* comm, void* stream) {
    return rkc_solve(args, (Comm*)comm, (cudaStream_t)stream);
}
int xsq_rkc_stage_bench(int32_t nx, int32_t rows, int32_t reps, double* ms_per_stage,
                        void* stream) {
    if (!ms_per_stage || nx < 4 || nx % 4 || rows < 1 || reps < 1) return XSQ_ERR_ARG;
    return rkc_stage_bench(nx, rows, reps, ms_per_stage, (cudaStream_t)stream);
}

int xsq_rkc_stage_bench_tma(int32_t nx, int32_t rows, int32_t reps, double* ms_per_stage,
                            double* max_abs_diff, void* stream) {
    if (!ms_per_stage || nx < 4 || nx % 4 || rows < 1 || reps < 1) return XSQ_ERR_ARG;
    return rkc_stage_bench_tma(nx, rows, reps, ms_per_stage, max_abs_diff, (cudaStream_t)stream);
}

int xsq_trim_memory(int device) {
    // scratch is cached in the device's stream-ordered pool between calls
    // (keep_pool_memory); hand it back to the driver, e.g. before another
    // library needs the memory
    XSQ_CUDA(cudaSetDevice(device));
    XSQ_CUDA(cudaDeviceSynchronize());
    cudaMemPool_t pool;
    XSQ_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    XSQ_CUDA(cudaMemPoolTrimTo(pool, 0));
    return XSQ_OK;
}

int xsq_profile_enable(int on) {
    g_profile.store(on ? 1 : 0);
    if (!on) g_prof_count = 0;
    return XSQ_OK;
}

int xsq_profile_get(int back, double* ms_init, double* ms_main, double* ms_probe) {
    if (back < 0 || back >= kProfRing || back >= g_prof_count) {
        g_detail = "no such profiled solve";
        return XSQ_ERR_ARG;
    }
    cudaEvent_t* ev = g_prof_ev[(g_prof_count - 1 - back) % kProfRing];
    XSQ_CUDA(cudaEventSynchronize(ev[3]));
    float a = 0, b = 0, c = 0;
    XSQ_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
    XSQ_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
    XSQ_CUDA(cudaEventElapsedTime(&c, ev[2], ev[3]));
    if (ms_init) *ms_init = a;
    if (ms_main) *ms_main = b;
    if (ms_probe) *ms_probe = c;
    return XSQ_OK;
}

int xsq_profile_last(double* ms_init, double* ms_main, double* ms_probe) {
    return xsq_profile_get(0, ms_init, ms_main, ms_probe);
}

int64_t xsq_launch_count(int reset) {
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

int xsq_fp64_peak(int device, int32_t iters, double* tflops) {
    if (!tflops || iters <= 0) return XSQ_ERR_ARG;
    XSQ_CUDA(cudaSetDevice(device));
    int n_sm = 0;
    XSQ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const int block = 256, grid = n_sm * 8;
    double* out = nullptr;
    XSQ_CUDA(cudaMalloc((void**)&out, sizeof(double) * grid * block));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_peak_kernel<<<grid, block>>>(out, iters / 4 + 1, 0.999999, 1e-9);  // warm-up
    count_launch();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
        count_launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(out);
    if (e != cudaSuccess) return cuda_fail(e, "fp64_peak_kernel");
    const double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * block;
    *tflops = flops / (best * 1e-3) / 1e12;
    return XSQ_OK;
}

}  // extern "C"
