// TEST INFRASTRUCTURE -- host emulation of xsq_rk_solve: the kernel bodies of
// extensisq_b200/csrc (ens_init_body, rk_fast_body / rk_persistent_body,
// stiff_queue_body) compiled by g++ (cuda_shim.h) and run by one host thread
// over all trajectories, with the parameter block built by the library's own
// build_params (xsq_params.h).  tests/test_kernel_host.py compares the result
// with the C oracle in device arithmetic: bit equality of the REAL kernel
// source against the oracle, without a GPU.
#define XSQ_HOST_EMU 1
#include "cuda_shim.h"
#ifdef XSQ_EMU_EVENTS
// Second build of this file (xsq_emu_events.so): the kernels as NVRTC compiles
// them for scipy's `events=`, with the three event functions of
// oracle/problems.py EVENT_SETS["lorenz_sections"] -- in-lane root location, the
// event queue with event_queue_body, and rk_fast with event hooks.
#define XSQ_EVENTS_N 3
namespace xsq {
static inline double user_event(int k, double t, const double* y, const double* p) {
    (void)t; (void)p;
    if (k == 0) return y[2] - 27.0;
    if (k == 1) return y[0];
    return y[0] * y[1] - 30.0;
}
}  // namespace xsq
#endif

#include <string>
#include <vector>

extern "C" const uint8_t* xsq_emu_rcp_bits = nullptr;
namespace xsq {
// oracle/xsq_devmath.h dev_rcp64h: high word of the IEEE quotient + table bit
double xsq_host_rcp64h(double x) {
    uint64_t b; std::memcpy(&b, &x, 8);
    b &= 0xffffffff00000000ULL;
    double xt; std::memcpy(&xt, &b, 8);
    const double q = 1.0 / xt;
    const uint32_t m = (uint32_t)(b >> 32) & 0xfffffu;
    uint64_t qb; std::memcpy(&qb, &q, 8);
    uint64_t hi = qb >> 32;
    hi += (xsq_emu_rcp_bits[m >> 3] >> (m & 7)) & 1u;
    hi <<= 32;
    double r; std::memcpy(&r, &hi, 8);
    return r;
}
}  // namespace xsq

#include "xsq_rk_fast.cuh"
#include "xsq_swag_core.cuh"
#include "xsq_swag_fast.cuh"
#include "xsq_rhs.cuh"
#include "xsq_user.h"

namespace xsq {
static std::string g_detail;
void set_detail(const std::string& s) { g_detail = s; }
void count_launch() {}
bool user_tableau_info(MethodInfo*) { return false; }
bool user_rhs_shape(int, int*, int*) { return false; }
#ifdef XSQ_EMU_EVENTS
int user_events_count(int) { return XSQ_EVENTS_N; }
#else
int user_events_count(int) { return -1; }
#endif
}  // namespace xsq
#include "xsq_params.h"

using namespace xsq;

template <class Tab, class R>
static int run(RkDev P, const MethodInfo& mi, bool want_fast, int* used_fast) {
    const long long N = P.n_lanes;
    gridDim = {1, 1, 1};
    blockDim = {1, 1, 1};
    threadIdx = {0, 0, 0};
    for (long long i = 0; i < N; ++i) {          // ens_init: one thread per lane
        blockIdx = {(unsigned)i, 0, 0};
        ens_init_body<R>(P);
    }
    blockIdx = {0, 0, 0};
    *used_fast = 0;
    if constexpr (Tab::VARIANT == tab::GENERIC && !R::WARP) {
#ifdef XSQ_EMU_EVENTS
        // xsq_user.cpp fast_events_variant
        bool ok = want_fast && P.evq_cap > 0 && P.evq_exact && P.n_forced == 0 && P.n_eval == 0 &&
                  P.minalpha == 0.0 && P.max_steps == 0x7fffffff;
        if (ok) {
#else
        if (want_fast && fast_eligible<Tab, R>(P)) {
#endif
            fast_prepare<Tab>(P);
            if (P.nfev_stiff_detect > 0) rk_fast_body<Tab, R, 1, true>(P);
            else rk_fast_body<Tab, R, 1, false>(P);
            *used_fast = 1;
        }
    }
    if (!*used_fast) rk_persistent_body<Tab, R>(P);
    if (P.stiff_q_cap > 0) stiff_queue_body<R>(P, mi.s, mi.stbrad, mi.tanang);
#ifdef XSQ_EMU_EVENTS
    if (P.evq_cap > 0) {
        blockDim = {128, 1, 1};                 // the queue kernel's tile size
        for (unsigned tx = 0; tx < 128; ++tx) {
            threadIdx = {tx, 0, 0};
            event_queue_body<Tab, R>(P);
        }
        blockDim = {1, 1, 1};
        threadIdx = {0, 0, 0};
    }
#endif
    return 0;
}

// Runge-Kutta-Nystrom methods: second order problems ([x, v] layouts) only
template <class Tab>
static int run_rkn(int rhs, const RkDev& P, const MethodInfo& mi, int* used) {
    if constexpr (Tab::VELOCITY_DEPENDENT) {
        switch (rhs) {
            case XSQ_RHS_VANDERPOL: return run<Tab, rhs::VanDerPol>(P, mi, false, used);
            case XSQ_RHS_ARENSTORF: return run<Tab, rhs::Arenstorf>(P, mi, false, used);
            default: return XSQ_ERR_UNSUPPORTED;
        }
    } else {
        return XSQ_ERR_UNSUPPORTED;     // MR6NN: no built-in velocity independent lane-per-system rhs
    }
}

template <class Tab>
static int run_rhs(int rhs, const RkDev& P, const MethodInfo& mi, bool fast, int* used) {
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return run<Tab, rhs::Lorenz63>(P, mi, fast, used);
        case XSQ_RHS_VANDERPOL: return run<Tab, rhs::VanDerPol>(P, mi, fast, used);
        case XSQ_RHS_ARENSTORF: return run<Tab, rhs::Arenstorf>(P, mi, fast, used);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

#ifdef XSQ_EMU_EVENTS
static long long g_evq_records = -1;      // -1: every possible record (exact), 0: no queue
extern "C" void xsq_emu_set_event_queue(long long records) { g_evq_records = records; }
#endif
extern "C" const char* xsq_emu_detail() { return g_detail.c_str(); }
extern "C" void xsq_emu_set_rcp_table(const uint8_t* bits) { xsq_emu_rcp_bits = bits; }

// All pointers of `a` are HOST pointers (same SoA layout as the device ABI).
extern "C" int xsq_emu_rk_solve(const xsq_rk_args_t* a, int want_fast, long long queue_records,
                                int* used_fast) {
    if (!xsq_emu_rcp_bits) return XSQ_ERR_ARG;
    RkDev P;
    MethodInfo mi;
    std::vector<double> atol;
    int rc = build_params(a, &P, &mi, &atol);
    if (rc != XSQ_OK) return rc;
    if (a->n_lanes == 0) return XSQ_OK;
    // scratch, as solve_device (xsq_api.cu)
    const size_t N = (size_t)a->n_lanes, ns = atol.size();
    std::vector<unsigned long long> counters(2, 0ULL);
    std::vector<double> init_h(N), init_f0(N * ns);
    std::vector<int> init_nfev(N);
    P.queue = &counters[0];
    P.stiff_q_count = &counters[1];
    P.atol_dev = atol.data();
    P.init_h = init_h.data();
    P.init_f0 = init_f0.data();
    P.init_nfev = init_nfev.data();
    P.morder = mi.order2;
    std::vector<double> slots;
    P.stiff_slot = nullptr;
    P.stiff_threads = 0;
    P.stiff_q = nullptr;
    P.stiff_q_cap = 0;
    if (P.nfev_stiff_detect > 0) {
        const size_t rec = 5 + 4 * (size_t)a->n_state + (size_t)(a->n_param > 0 ? a->n_param : 1);
        const size_t threads = 1;
        size_t qcap = queue_records >= 0 ? (size_t)queue_records : N * 16;
        slots.assign((2 * threads + qcap) * rec, 0.0);
        P.stiff_threads = (long long)threads;
        P.stiff_slot = slots.data();
        P.stiff_q = slots.data() + 2 * threads * rec;
        P.stiff_q_cap = (long long)qcap;
    }
#ifdef XSQ_EMU_EVENTS
    // the event queue, as solve_device (xsq_api.cu)
    std::vector<double> evq_mem;
    std::vector<unsigned> evq_fill;
    unsigned long long evq_count = 0ULL;
    P.evq = nullptr; P.evq_cap = 0; P.evq_count = &evq_count; P.evq_fill = nullptr; P.evq_exact = 0;
    if (a->events != 0 && P.n_events > 0 && P.ev_capacity > 0) {
        const size_t fields = (5 + (size_t)(mi.s + 3) * (size_t)a->n_state + 1) & ~(size_t)1;
        const size_t need = N * (size_t)P.n_events * (size_t)P.ev_capacity + (size_t)kEvqChunk;
        size_t qcap = need;
        if (g_evq_records >= 0 && (size_t)g_evq_records < qcap) qcap = (size_t)g_evq_records;
        const size_t chunks = (qcap + kEvqChunk - 1) / kEvqChunk;
        if (chunks > 0) {
            evq_mem.assign(chunks * kEvqChunk * fields, 0.0);
            evq_fill.assign(chunks, 0u);
            P.evq = evq_mem.data();
            P.evq_fill = evq_fill.data();
            P.evq_cap = (long long)(chunks * kEvqChunk);
            P.evq_exact = chunks * kEvqChunk >= need ? 1 : 0;
        }
    }
    if (a->rhs != XSQ_RHS_LORENZ63) return XSQ_ERR_UNSUPPORTED;
    switch (a->method) {
        case XSQ_TS5: return run<tab::Ts5, rhs::Lorenz63>(P, mi, want_fast, used_fast);
        case XSQ_BS5: return run<tab::BS5, rhs::Lorenz63>(P, mi, want_fast, used_fast);
        case XSQ_PR8: return run<tab::Pr8, rhs::Lorenz63>(P, mi, want_fast, used_fast);
        case XSQ_CKDISC: return run<tab::CKdisc, rhs::Lorenz63>(P, mi, want_fast, used_fast);
        default: return XSQ_ERR_UNSUPPORTED;
    }
#endif
    switch (a->method) {
        case XSQ_TS5: return run_rhs<tab::Ts5>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_BS5: return run_rhs<tab::BS5>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_CK5: return run_rhs<tab::CK5>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_ME4: return run_rhs<tab::Me4>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_PR7: return run_rhs<tab::Pr7>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_PR8: return run_rhs<tab::Pr8>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_PR9: return run_rhs<tab::Pr9>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_CFMR7OSC: return run_rhs<tab::CFMR7osc>(a->rhs, P, mi, want_fast, used_fast);
        case XSQ_CKDISC: return run_rhs<tab::CKdisc>(a->rhs, P, mi, want_fast, used_fast);
#ifndef XSQ_EMU_EVENTS
        case XSQ_FI4N: return run_rkn<tab::Fi4N>(a->rhs, P, mi, used_fast);
        case XSQ_FI5N: return run_rkn<tab::Fi5N>(a->rhs, P, mi, used_fast);
        case XSQ_MU5NMB: return run_rkn<tab::Mu5Nmb>(a->rhs, P, mi, used_fast);
        case XSQ_MR6NN: return run_rkn<tab::MR6NN>(a->rhs, P, mi, used_fast);
#endif
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

template <class R>
static int run_swag(RkDev P) {
    const long long N = P.n_lanes;
    gridDim = {1, 1, 1};
    blockDim = {1, 1, 1};
    threadIdx = {0, 0, 0};
    for (long long i = 0; i < N; ++i) {
        blockIdx = {(unsigned)i, 0, 0};
        ens_init_body<R>(P);
    }
    blockIdx = {0, 0, 0};
    if constexpr (!R::WARP && R::NL <= 4) {
        if (swag_fast_eligible<R>(P)) {
            swag_fast_body<R, 1>(P);
            return 0;
        }
    }
    swag_persistent_body<R>(P);
    return 0;
}

// xsq_swag_solve (xsq_api.cu) on the host
extern "C" int xsq_emu_swag_solve(const xsq_rk_args_t* args, int k_max) {
    if (!xsq_emu_rcp_bits || k_max < 1 || k_max > 12) return XSQ_ERR_ARG;
    xsq_rk_args_t a = *args;
    a.method = XSQ_METHOD_SWAG;
    a.interpolant = XSQ_INTERP_FREE;
    a.use_sc_params = 0;
    a.h_forced = nullptr;
    a.n_forced = 0;
    a.reserved0 = k_max;
    RkDev P;
    MethodInfo mi;
    std::vector<double> atol;
    int rc = build_params(&a, &P, &mi, &atol);
    if (rc != XSQ_OK) return rc;
    if (a.n_lanes == 0) return XSQ_OK;
    const size_t N = (size_t)a.n_lanes, ns = atol.size();
    std::vector<unsigned long long> counters(2, 0ULL);
    std::vector<double> init_h(N), init_f0(N * ns);
    std::vector<int> init_nfev(N);
    P.queue = &counters[0];
    P.stiff_q_count = &counters[1];
    P.atol_dev = atol.data();
    P.init_h = init_h.data();
    P.init_f0 = init_f0.data();
    P.init_nfev = init_nfev.data();
    P.morder = 1;
    switch (a.rhs) {
        case XSQ_RHS_LORENZ63: return run_swag<rhs::Lorenz63>(P);
        case XSQ_RHS_VANDERPOL: return run_swag<rhs::VanDerPol>(P);
        case XSQ_RHS_ARENSTORF: return run_swag<rhs::Arenstorf>(P);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}
