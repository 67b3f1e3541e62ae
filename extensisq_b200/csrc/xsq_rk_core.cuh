// xsq_rk_core.cuh -- the tableau-driven adaptive explicit Runge-Kutta step as
// a persistent fp64 kernel: one lane (thread) or one warp per ODE system,
// state + all stage vectors in registers, per-lane step control, finished
// lanes retired by warp ballot and refilled from a global work queue.
//
// What it computes is what these reference functions compute (file:line in
// /root/reference/extensisq):
//   RungeKutta.__init__            common.py:187-220
//   _init_min_step_parameters      common.py:123-148   (H_MIN_A, h_min_b)
//   _init_sc_control               common.py:166-185   (host side, xsq_api.cu)
//   h_start                        common.py:519-763   -> h_start_dev()
//   _reassess_stepsize             common.py:310-331   -> reassess()
//   _rk_stage                      common.py:353-356   -> stage sums below
//   _comp_sol_err/_estimate_error  common.py:333-351
//   calculate_scale, norm          common.py:57-66
//   controller in _step_impl       common.py:249-287
//   BS5._step_impl + pre-error     bogacki.py:238-346
//   CFMR7osc._step_impl + pre-err  calvo.py:152-261
//   _dense_output_impl / Horner    common.py:358-368, 766-790
//   BS5 interpolants               bogacki.py:348-393
//   _diagnose_stiffness, stiff_a-d common.py:370-516, 824-1204 -> Lane::diagnose,
//                                  stiff_probe_impl, stiff_queue (probe queue)
//   CKdisc._step_impl              cash.py:245-416     -> Lane::attempt_ckdisc
//   solve_ivp t_eval slicing       scipy/integrate/_ivp/ivp.py:711-728
//   solve_ivp events (third party) scipy/integrate/_ivp/ivp.py find_active_events,
//                                  handle_events, solve_event_equation; optimize/
//                                  Zeros/brentq.c      -> Lane::after_step, event_root
//                                  (only in kernels built with XSQ_EVENTS_N)
//   OdeSolver.step finish test     scipy/integrate/_ivp/base.py:195-210
//
// This header is also the translation unit NVRTC compiles for user-supplied
// right-hand sides, so it includes nothing from the host toolchain.
#pragma once
#include "xsq_tableaux_gen.cuh"
#include "xsq_math_tables_gen.cuh"
#include "xsq_math.cuh"
#include "xsq_intrin.cuh"

namespace xsq {

// ---- machine constants (numpy.finfo(float64)) ------------------------------
#define XSQ_SQRT_TINY 0x1.0p-511               /* sqrt(2.2250738585072014e-308) */
#define XSQ_BIG 0x1.fffffffffffffp+511         /* sqrt(DBL_MAX)                 */
#define XSQ_SMALL 0x1.0000000000001p-53        /* nextafter(epsneg, 1)          */
#define XSQ_RELPER 0x1.172b83c7d517bp-20       /* XSQ_SMALL ** 0.375  (libm)    */
#define XSQ_NAN __longlong_as_double(0x7ff8000000000000LL)
#define XSQ_INF __longlong_as_double(0x7ff0000000000000LL)

static constexpr double kMinFactor = 0.2;       // common.py:18
static constexpr double kMaxFactor = 4.0;       // common.py:19
static constexpr double kMaxFactor0 = 10.0;     // common.py:20

enum LaneStatus : int {
    LANE_FINISHED = 0, LANE_TOO_SMALL = -1, LANE_OVERFLOW = -2,
    LANE_STEP_BUDGET = -5, LANE_RUNNING = 1,
    LANE_EVQ_FULL = -6,  // the event queue ran out (cannot happen when RkDev::evq_exact)
    LANE_FLUSH = 2,     // internal: still running, stiffness probe slots are full
    LANE_EVENT = 3,     // internal: a terminal event ended the trajectory (status 1)
    LANE_EVCHECK = 4    // internal (rk_fast): the accepted step may hold a terminal event
};
#ifndef XSQ_MAX_BLOCK
#define XSQ_MAX_BLOCK 256   // largest CTA any rk_persistent geometry launches
#endif
enum Interp : int { IP_FREE = 1, IP_LOW = 2, IP_BEST = 3 };
constexpr int kEvqChunk = 512;   // records per chunk of the event queue

// Device-side parameter block; passed by value as the kernel argument so every
// field is a constant-bank operand.
struct RkDev {
    long long n_lanes;
    const double* y0;
    const double* params;
    double t0, t_bound, direction;
    double rtol;
    double atol[16];
    const double* atol_dev;       // [n_state], used by warp-per-system kernels
    double first_step, max_step;
    const double* first_h;      // per-lane first |h| (resume), or null
    double err_exp, minbeta1, minbeta2, minalpha, safety, safety_sc;
    double log2n;                 // log2(n_state), for log2(error_norm)
    // step-size controller in the log2 domain (ctl_factor below), with
    // l2 = log2(sum((err/scale)^2)) = 2 log2(error_norm) + log2 n:
    //   safety    * error_norm^err_exp                     = 2^(a1s l2 + a0s)
    //   safety_sc * error_norm^minbeta1 * err_old^minbeta2 = 2^(a1c l2 + a2c l2_old + a0c)
    CtlConst ctl;
    // rk_fast: high words that bracket "min_step < h_abs < max_step" (xsq_rk_fast.cuh)
    int fast_hi_min, fast_hi_span;
    int fast_dir_mask;            // 0 for forward integration, 0x80000000 for backward
    const double* t_eval;
    double* y_eval;
    const double* h_forced;
    double* t_final;
    double* y_final;
    double* h_next;
    int* n_acc;
    int* n_rej;
    int* nfev;
    int* status;
    int* n_eval_done;
    unsigned long long* queue;
    int n_eval, n_forced, max_steps, interpolant;
    int eval_pitch;               // row pitch of y_eval in doubles, multiple of 4
    // written by ens_init (one convergent pass over all lanes), read at refill
    double* init_f0;              // SoA [n_state][n_lanes]: f(t0, y0)
    double* init_h;               // [n_lanes]: |h| from h_start (if needed)
    int* init_nfev;               // [n_lanes]: evaluations spent so far
    int morder;                   // h_start's order argument (common.py:210-212)
    // stiffness diagnosis (common.py:150-164, 370-516); 0 = off
    int nfev_stiff_detect;
    int stiff_many_steps;         // nfev_stiff_detect // n_stages
    int* stiff_flags;             // [n_lanes] OR of STIFF_* codes, may be null
    // deferred probes: 2 slots per resident thread, SoA [slot][field][thread]
    double* stiff_slot;
    long long stiff_threads;      // thread stride of stiff_slot (>= grid size)
    // probe queue: records of StiffSlot<R>::DOUBLES doubles, appended by the
    // persistent kernel, worked off by stiff_queue after it
    double* stiff_q;
    long long stiff_q_cap;                  // records; 0 = no queue
    unsigned long long* stiff_q_count;      // may run past stiff_q_cap
    // events (scipy solve_ivp `events=`; only kernels built with XSQ_EVENTS_N)
    int n_events, ev_capacity;
    int ev_terminal[8];           // 0: never; k: stop at the k-th occurrence
    int ev_direction[8];          // -1, 0, +1
    double* t_events;             // [n_lanes][n_events][ev_capacity]
    double* y_events;             // [n_lanes][n_events][ev_capacity][n_state]
    int* ev_count;                // [n_lanes][n_events]
    // event queue: a step with a sign change that cannot end the trajectory is
    // appended here (stages + both end states) and the root is located by
    // event_queue after the persistent kernel; SoA [field][evq_cap]
    // The queue is handed out in chunks of kEvqChunk records: a CTA of the
    // persistent kernel appends through a counter in ITS shared memory -- a lane
    // that waits for a global atomic keeps its whole warp waiting (measured: the
    // largest stall of the first version) -- and takes a new chunk from the global
    // counter when its current one is full (once per kEvqChunk appends).
    double* evq;
    long long evq_cap;                      // records (multiple of kEvqChunk); 0 = no queue
    unsigned long long* evq_count;          // chunks handed out; may run past the capacity
    unsigned* evq_fill;                     // [evq_cap / kEvqChunk] records written per chunk
    int evq_exact;                          // host: the queue holds every possible record
};

// ---- reductions over one system -------------------------------------------
template <bool WARP>
__device__ __forceinline__ double sys_sum(double x) {
    if (WARP) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    return x;
}
template <bool WARP>
__device__ __forceinline__ double sys_min(double x) {
    if (WARP) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    }
    return x;
}

// Python's built-in max(a, b) / min(a, b): "b if b > a else a".  A plain
// compare-and-select (3 instructions); fmax()/fmin() cost ~8 on sm_100 because
// of their NaN rules, and these are the reference's semantics anyway.
__device__ __forceinline__ double pymax(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double pymin(double a, double b) { return b < a ? b : a; }

// ---- log2 / exp2 for the step-size controller -------------------------------
// error_norm ** x  (common.py:257, 264-266, 281) is evaluated as
// exp2(x * log2(error_norm)).  Both functions are table driven (tables in
// shared memory, one LDS.128 + one LDS.64 / one LDS.128 per call; generated by
// tools/gen_math_tables.py) followed by a short polynomial:
//   log2:  x = 2^e m,  i = top 7 mantissa bits,  r = fma(m, inv_i, -1), |r| < 2^-8,
//          log2 x = (e + L_hi) + fma(r, q(r), L_lo)              10 fp64 operations
//   exp2:  z = n + j/64 + r, |r| <= 2^-7,  2^z = 2^n (T_hi + fma(T_hi, p(r), T_lo))
//                                                                 10 fp64 operations
// max error 1.2 ulp (exp2) / 1 ulp, and 1.2e-16 absolute near x = 1 (log2),
// measured against mpmath.  The kernel is bound by instruction issue (an fp64
// instruction occupies its scheduler for two cycles, anything else for one),
// so what counts is the number of operations, not the depth of the chain.
// Every operation is an IEEE add / mul / fma on table constants: the C oracle
// (oracle/xsq_devmath.h) repeats them bit for bit.
// The *_core forms are total functions on bit patterns (finite garbage for 0,
// Inf, NaN); the callers decide those cases before they use the value.
static __constant__ __align__(16) double c_xsq_havg[2] = {0.9, 0.1};       // common.py:372

struct MathTabs {
    double lg[128 * 4];      // inv_i, L_hi, L_lo, pad
    double e2[64 * 2];       // T_hi, T_lo
};
__device__ __forceinline__ MathTabs& math_tabs() {
    __shared__ __align__(16) MathTabs s;
    return s;
}
// every kernel that calls log2_* / exp2_* runs this first (all threads)
__device__ __forceinline__ void math_tabs_init() {
    MathTabs& m = math_tabs();
    for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) m.lg[i] = c_xsq_lg_tab[i];
    for (int i = threadIdx.x; i < 64 * 2; i += blockDim.x) m.e2[i] = c_xsq_e2_tab[i];
    __syncthreads();
}

__device__ __forceinline__ double log2_core(double x) {
    const double* T = math_tabs().lg + log2_tab_offset(x);
    const double2 t01 = *reinterpret_cast<const double2*>(T);
    return log2_arith(x, t01.x, t01.y, T[2], c_xsq_lg_pol);
}

__device__ __forceinline__ double exp2_core(double z) {
    const Exp2Split s = exp2_split(z);
    const double2 T = *reinterpret_cast<const double2*>(math_tabs().e2 + ((s.N & 63) << 1));
    return exp2_arith(s, T.x, T.y, c_xsq_e2_pol);
}

// log2 with the special values: 0 -> -inf, +inf -> +inf, NaN -> NaN (a
// subnormal argument gives an irrelevant finite value: such an error norm is
// below the tiny_err threshold anyway)
__device__ __forceinline__ double log2_fast(double x) {
    double r2 = log2_core(x);
    if (x == 0.0) r2 = -XSQ_INF;
    if (!(x < XSQ_INF)) r2 = x;
    return r2;
}
// exp2 for |z| < 1000; a non-finite z yields NaN, which the caller's
// max()/min() absorb exactly like the reference's max(min_factor, nan)
__device__ __forceinline__ double exp2_fast(double z) {
    if (!(fabs(z) < 1000.0)) return XSQ_NAN;
    return exp2_core(z);
}

// 1/x to ~2^-40: seed + ONE Newton step.  Used for err/scale only: the error
// norm feeds a compare against 1 and a power with |exponent| <= 0.25, so a
// relative 1e-12 moves an accept/reject decision with probability ~1e-12 per
// step (measured: step counts unchanged on every parity lane).
__device__ __forceinline__ double rcp_scale(double x) {
    const double r = rcp64h_seed(x);
    const double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// ---- the step-size controller, common.py:249-287 ---------------------------
// factor by which h_abs is multiplied after an attempt.  l2 / l2_old are
// log2_core(ss) of this attempt and of the last accepted step (ss = sum of
// squares of err/scale); the caller has decided accept / tiny / the flags.
// Accepting and rejecting lanes share ONE exp2; a NaN / Inf error norm is
// handled by the caller (`bad`: the reference's max(0.2, nan) is 0.2).
//   z_extra: minalpha * log2(h / h_prev), or 0 (all built-in presets)
struct Exp2Default {
    __device__ __forceinline__ double operator()(double z) const { return exp2_core(z); }
};
template <bool EXTRA, class E2 = Exp2Default>
__device__ __forceinline__ double ctl_factor(const RkDev& P, double l2, double l2_old,
                                             double z_extra, bool accept, bool second,
                                             bool rej, bool tiny, double max_factor,
                                             E2 e2 = E2()) {
    return ctl_factor_arith<EXTRA>(P.ctl, l2, l2_old, z_extra, accept, second, rej, tiny,
                                   max_factor, e2);
}

// RMS norm, common.py:64-66
template <class R>
__device__ __forceinline__ double rms(const double (&x)[R::NL]) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < R::NL; ++c) s = fma(x[c], x[c], s);
    s = sys_sum<R::WARP>(s);
    return sqrt(s / (double)R::N);
}

// Warp-per-system policies whose size is not a multiple of 32 have padded
// slots (comp(k, lane) >= N): they hold zeros, are never loaded or stored, and
// add exact zeros to every norm.
template <class R>
__device__ __forceinline__ bool slot_ok(int k, int lane) {
    if constexpr (R::PADDED) return R::comp(k, lane) < R::N;
    else return true;
}
template <class R>
__device__ __forceinline__ double load_slot(const double* __restrict__ base, long long n_lanes,
                                            long long idx, int k, int lane) {
    return slot_ok<R>(k, lane) ? base[(long long)R::comp(k, lane) * n_lanes + idx] : 0.0;
}
template <class R>
__device__ __forceinline__ void store_slot(double* __restrict__ base, long long n_lanes,
                                           long long idx, int k, int lane, double v) {
    if (slot_ok<R>(k, lane)) base[(long long)R::comp(k, lane) * n_lanes + idx] = v;
}

template <class R>
__device__ __forceinline__ double atol_of(const RkDev& P, int k, int lane) {
    if (R::WARP) return slot_ok<R>(k, lane) ? P.atol_dev[R::comp(k, lane)] : 1.0;
    return P.atol[k];
}

// ---- dense-output staging --------------------------------------------------
// y_eval is [n_lanes][n_state][eval_pitch] (scipy's sol.y per lane).  A lane
// produces its t_eval points in order, so it collects four consecutive points
// of a component in shared memory and writes them as one aligned 32-byte
// sector (two 16-byte stores) instead of four scattered 8-byte stores: 4x
// fewer store instructions and only full-sector writes reach L2/HBM.
// Shared layout [component][slot][thread]: conflict-free for any mix of slots.
#ifndef XSQ_HOST_EMU
extern __shared__ double xsq_eval_stage[];
#else
static double xsq_eval_stage[4 * 64];      // host emulation: one thread per block
#endif

template <class R>
__device__ __forceinline__ void eval_put(const RkDev& P, long long sys, int lane,
                                         int i, const double (&v)[R::NL]) {
    const int slot = i & 3;
    const int tid = threadIdx.x, bs = blockDim.x;
#pragma unroll
    for (int c = 0; c < R::NL; ++c)
        xsq_eval_stage[(c * 4 + slot) * bs + tid] = v[c];
    if (slot == 3 || i == P.n_eval - 1) {
        const int i0 = i & ~3;
#pragma unroll
        for (int c = 0; c < R::NL; ++c) {
            if (!slot_ok<R>(c, lane)) continue;
            const long long row = sys * (long long)R::N + R::comp(c, lane);
            double2* dst = reinterpret_cast<double2*>(
                P.y_eval + row * P.eval_pitch + i0);
            dst[0] = make_double2(xsq_eval_stage[(c * 4 + 0) * bs + tid],
                                  xsq_eval_stage[(c * 4 + 1) * bs + tid]);
            dst[1] = make_double2(xsq_eval_stage[(c * 4 + 2) * bs + tid],
                                  xsq_eval_stage[(c * 4 + 3) * bs + tid]);
        }
    }
}

// A lane that ends early (failure, or a zero-length span): flush the staged
// partial group, then fill the remaining points with `fill` (NaN: the
// reference returns only the points reached; y0 for t0 == t_bound).
template <class R>
__device__ __forceinline__ void eval_finish(const RkDev& P, long long sys,
                                            int lane, int ieval, bool constant,
                                            const double (&y)[R::NL]) {
    const int tid = threadIdx.x, bs = blockDim.x;
    const int i0 = ieval & ~3;
#pragma unroll
    for (int c = 0; c < R::NL; ++c) {
        if (!slot_ok<R>(c, lane)) continue;
        const long long row = sys * (long long)R::N + R::comp(c, lane);
        double* dst = P.y_eval + row * P.eval_pitch;
        for (int i = i0; i < ieval; ++i)
            dst[i] = xsq_eval_stage[(c * 4 + (i & 3)) * bs + tid];
        for (int i = ieval; i < P.n_eval; ++i) dst[i] = constant ? y[c] : XSQ_NAN;
    }
}

// ---- Watts' starting step, common.py:519-763 --------------------------------
template <class R>
__device__ double h_start_dev(const RkDev& P, double a, double b,
                              const double (&y)[R::NL],
                              const double (&yprime)[R::NL],
                              const double (&prm)[R::NPL], int morder,
                              int lane, int& nfev) {
    constexpr int NL = R::NL;
    const double big = XSQ_BIG, small = XSQ_SMALL, relper = XSQ_RELPER;
    double spy[NL], pv[NL], yp[NL], sf[NL];
    const double dx = b - a;
    const double absdx = fabs(dx);
    double da = copysign(fmax(fmin(relper * fabs(a), absdx),
                              100.0 * small * fabs(a)), dx);
    if (da == 0.0) da = relper * dx;
    R::f(a + da, y, prm, sf);
    ++nfev;
#pragma unroll
    for (int c = 0; c < NL; ++c) yp[c] = sf[c] - yprime[c];
    double delf = rms<R>(yp);
    double dfdxb = big;
    if (delf < big * fabs(da)) dfdxb = delf / fabs(da);
    double fbnd = rms<R>(sf);

    double dely = relper * rms<R>(y);
    if (dely == 0.0) dely = relper;
    dely = copysign(dely, dx);
    delf = rms<R>(yprime);
    fbnd = fmax(fbnd, delf);
    if (delf != 0.0) {
#pragma unroll
        for (int c = 0; c < NL; ++c) { spy[c] = yprime[c]; yp[c] = yprime[c]; }
    } else {
#pragma unroll
        for (int c = 0; c < NL; ++c) { spy[c] = 0.0; yp[c] = slot_ok<R>(c, lane) ? 1.0 : 0.0; }
        delf = rms<R>(yp);
    }
    double dfdub = 0.0;
    const int lk = R::N + 1 < 3 ? R::N + 1 : 3;
    for (int k = 1; k <= lk; ++k) {
        const double q = dely / delf;
#pragma unroll
        for (int c = 0; c < NL; ++c) pv[c] = fma(q, yp[c], y[c]);
        if (k == 2) {
            R::f(a + da, pv, prm, yp);
#pragma unroll
            for (int c = 0; c < NL; ++c) pv[c] = yp[c] - sf[c];
        } else {
            R::f(a, pv, prm, yp);
#pragma unroll
            for (int c = 0; c < NL; ++c) pv[c] = yp[c] - yprime[c];
        }
        ++nfev;
        fbnd = fmax(fbnd, rms<R>(yp));
        delf = rms<R>(pv);
        if (delf >= big * fabs(dely)) { dfdub = big; break; }
        dfdub = fmax(dfdub, delf / fabs(dely));
        if (k == lk) break;
        if (delf == 0.0) delf = 1.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            double dy;
            if (k == 2) dy = (y[c] != 0.0) ? y[c] : dely / relper;
            else dy = (pv[c] != 0.0) ? pv[c] : delf;
            if (spy[c] == 0.0) spy[c] = yp[c];
            yp[c] = (spy[c] != 0.0) ? copysign(dy, spy[c]) : dy;
            if (!slot_ok<R>(c, lane)) yp[c] = 0.0;          // padding stays zero
        }
        delf = rms<R>(yp);
    }
    const double ydpb = fma(dfdub, fbnd, dfdxb);
    double tolsum = 0.0, tolmin = XSQ_INF;
#pragma unroll
    for (int c = 0; c < NL; ++c) {
        const double etol = fma(P.rtol, fabs(y[c]), atol_of<R>(P, c, lane));
        // log10(etol) in the reference; 10^(c log10 x) = 2^(c log2 x), and the
        // table-driven log2 / exp2 are repeated bit for bit by the C oracle
        const double te = log2_fast(etol);
        if (slot_ok<R>(c, lane)) {
            tolsum += te;
            tolmin = fmin(tolmin, te);
        }
    }
    tolsum = sys_sum<R::WARP>(tolsum);
    tolmin = fmin(sys_min<R::WARP>(tolmin), big);
    const double tolp =
        exp2_fast(0.5 * (tolsum / (double)R::N + tolmin) / (double)(morder + 1));
    double h = absdx;
    if (ydpb == 0.0 && fbnd == 0.0) {
        if (tolp < 1.0) h = absdx * tolp;
    } else if (ydpb == 0.0) {
        if (tolp < fbnd * absdx) h = tolp / fbnd;
    } else {
        const double srydpb = sqrt(0.5 * ydpb);
        if (tolp < srydpb * absdx) h = tolp / srydpb;
    }
    if (dfdub != 0.0) h = fmin(h, 1.0 / dfdub);
    h = fmax(h, 100.0 * small * fabs(a));
    if (h == 0.0) h = small * fabs(b);
    return fabs(h);
}

// ---- initialisation pass ---------------------------------------------------
// f(t0, y0) and Watts' starting step for EVERY lane, one thread (or warp) per
// lane, fully convergent.  The persistent kernel then refills a finished lane
// with a handful of loads; doing the 4-5 RHS evaluations + log10/pow of
// h_start inside the refill would stall the other 31 lanes of the warp each
// time (~15 % of the run for trajectories of a few hundred steps).
template <class R>
__device__ __forceinline__ void ens_init_body(const RkDev& P) {
    math_tabs_init();
    const int lane = threadIdx.x & 31;
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long idx = R::WARP ? gthread / 32 : gthread;
    if (idx >= P.n_lanes) return;
    double y[R::NL], f[R::NL], prm[R::NPL];
#pragma unroll
    for (int k = 0; k < R::NL; ++k)
        y[k] = load_slot<R>(P.y0, P.n_lanes, idx, k, lane);
    R::load_params(P.params, idx, P.n_lanes, lane, prm);
    int nfev = 1;
    R::f(P.t0, y, prm, f);
    double h = 0.0;
    if (P.first_h != nullptr) {
        h = fmin(P.first_h[idx], fabs(P.t_bound - P.t0));       // resume: the caller's step
    } else if (P.n_forced == 0 && !(P.first_step > 0.0)) {
        const double b = P.t0 + P.direction *
            fmin(fabs(P.t_bound - P.t0), P.max_step);
        h = h_start_dev<R>(P, P.t0, b, y, f, prm, P.morder, lane, nfev);
    }
#pragma unroll
    for (int k = 0; k < R::NL; ++k)
        store_slot<R>(P.init_f0, P.n_lanes, idx, k, lane, f[k]);
    if (!R::WARP || lane == 0) {
        P.init_h[idx] = h;
        P.init_nfev[idx] = nfev;
    }
}

template <class R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) ens_init(const RkDev P) {
    ens_init_body<R>(P);
}

// ---- stiffness diagnosis ----------------------------------------------------
// Port of the reference's _diagnose_stiffness / stiff_a..d (common.py:370-516,
// 824-1204, themselves a port of RKSuite): a nonlinear power iteration for the
// two dominant eigenvalues of havg*J, compared with the method's stability
// radius.  It never changes t, y or h; it spends a few RHS evaluations (counted
// in nfev, as in the reference) and, instead of warnings, sets per-lane flags.
// Runs every nfev_stiff_detect/n_stages accepted steps or after >= 10 failures
// in 40 steps, so it is kept out of line: only copies of the lane's vectors
// are handed to it, the hot loop's register allocation is untouched.
enum : int { STIFF_REAL = 1, STIFF_COMPLEX = 2, STIFF_OSCILLATORY = 4 };

// <a, b> = sum (a/wt)(b/wt) over one system.  Every vector the probe makes is
// divided by wt ONCE (the quotient is the same number wherever the reference
// recomputes it), so the inner products are plain FMAs: 4x fewer divisions
// and a much smaller probe -- its size matters, the hot loop of a 13-stage
// pair already fills the instruction cache.
template <class R>
__device__ __forceinline__ double wdotq(const double (&aq)[R::NL], const double (&bq)[R::NL]) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < R::NL; ++c) s = fma(aq[c], bq[c], s);
    return sys_sum<R::WARP>(s);
}

// everything stiff_d needs that does not change during one probe
template <class R>
struct StiffCtx {
    double havg, x, scale;
    double y[R::NL], fxy[R::NL], wt[R::NL], prm[R::NPL];
};

// stiff_d: z ~ havg * J * v by a difference of f; zq = z / wt; returns <z, z>.
// Out of line: called three times per iteration, and it holds the only copy
// of the right-hand side in the probe.
template <class R>
__device__ __forceinline__ double stiff_jac_times_impl(const StiffCtx<R>& C,
                                                       const double (&v)[R::NL], double vdotv,
                                                       double (&z)[R::NL], double (&zq)[R::NL]) {
    const double temp1 = C.scale / sqrt(vdotv);
    double yp[R::NL];
#pragma unroll
    for (int c = 0; c < R::NL; ++c) yp[c] = fma(temp1, v[c], C.y[c]);
    R::f(C.x, yp, C.prm, z);
    const double q = C.havg / temp1;
#pragma unroll
    for (int c = 0; c < R::NL; ++c) {
        z[c] = q * (z[c] - C.fxy[c]);
        zq[c] = z[c] / C.wt[c];
    }
    return wdotq<R>(zq, zq);
}

template <class R>
__device__ __noinline__ double stiff_jac_times_ool(const StiffCtx<R>& C, const double (&v)[R::NL],
                                                   double vdotv, double (&z)[R::NL],
                                                   double (&zq)[R::NL]) {
    return stiff_jac_times_impl<R>(C, v, vdotv, z, zq);
}

// SMALL: inside the persistent kernel, where code size is what matters
template <class R, bool SMALL>
__device__ __forceinline__ double stiff_jac_times(const StiffCtx<R>& C, const double (&v)[R::NL],
                                                  double vdotv, double (&z)[R::NL],
                                                  double (&zq)[R::NL]) {
    if constexpr (SMALL) return stiff_jac_times_ool<R>(C, v, vdotv, z, zq);
    else return stiff_jac_times_impl<R>(C, v, vdotv, z, zq);
}

// stiff_b
__device__ __forceinline__ bool stiff_dominant_real(double v1v1, double v0v1, double v0v0,
                                                    double& rold, double& rho,
                                                    double (&root1)[2], double (&root2)[2]) {
    const double r = v0v1 / v0v0;
    rho = fabs(r);
    const double det = v0v0 * v1v1 - v0v1 * v0v1;
    const double res = fabs(det / v0v0);
    const bool rootre = det == 0.0 || (res <= 1e-6 * v1v1 && fabs(r - rold) <= 0.001 * rho);
    root1[0] = rootre ? r : 0.0;
    root1[1] = root2[0] = root2[1] = 0.0;
    rold = r;
    return rootre;
}

// stiff_c
__device__ __forceinline__ void stiff_quadratic_roots(double alpha, double beta,
                                                      double (&r1)[2], double (&r2)[2]) {
    r1[0] = r1[1] = r2[0] = r2[1] = 0.0;
    const double temp = alpha / 2;
    const double disc = temp * temp - beta;
    if (disc == 0.0) { r1[0] = r2[0] = -temp; return; }
    const double sqdisc = sqrt(fabs(disc));
    if (disc < 0.0) { r1[0] = r2[0] = -temp; r1[1] = sqdisc; r2[1] = -sqdisc; }
    else { r1[0] = temp > 0.0 ? -temp - sqdisc : -temp + sqdisc; r2[0] = beta / r1[0]; }
}

// stiff_a + the classification of common.py:424-455.  All inputs come from a
// probe slot (see Lane::diagnose); returns flags | evaluations << 8.
template <class R>
struct StiffSlot {
    // x, hnow, havg, lotsfl, trajectory index, then y_new, y_old, f_new, h*K.E
    // and the parameters
    static constexpr int HEAD = 5;
    static constexpr int DOUBLES = HEAD + 4 * R::NL + R::NPL;
};

template <class R, bool SMALL>
__device__ __forceinline__ int stiff_probe_impl(const double* slot, long long stride, double xend,
                                            int maxfcn, int cost, double stbrad,
                                            double tanang) {
    constexpr int NL = R::NL;
    StiffCtx<R> C;
    C.x = slot[0];
    C.havg = slot[2 * stride];
    const double x = C.x, hnow = slot[stride], havg = C.havg;
    const bool lotsfl = slot[3 * stride] != 0.0;
    const double* q = slot + StiffSlot<R>::HEAD * stride;
    double v0[NL], v1[NL], v2[NL], v3[NL];        // the Krylov vectors ...
    double v0q[NL], v1q[NL], v2q[NL], v3q[NL];    // ... and each divided by wt
#pragma unroll
    for (int c = 0; c < NL; ++c) {
        C.y[c] = q[c * stride];
        C.fxy[c] = q[(2 * NL + c) * stride];
        v0[c] = q[(3 * NL + c) * stride];
        C.wt[c] = fmax(0.5 * (fabs(C.y[c]) + fabs(q[(NL + c) * stride])), XSQ_SQRT_TINY);
    }
#pragma unroll
    for (int c = 0; c < R::NPL; ++c) C.prm[c] = q[(4 * NL + c) * stride];
    int nfev = 0;
    const double epsneg = 0x1.0p-53;
    int stif = 0, rootre = -1;          // stif: 1 / 0 / -1 (unsure)
    bool have_root = false;
    double root1[2] = {0, 0}, root2[2] = {0, 0}, rho = 0.0;
    do {
        if (fabs(hnow / havg) > 5 || fabs(hnow / havg) < 0.2) break;
        if (cost * fabs((xend - x) / havg) <= maxfcn) break;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            v1q[c] = C.y[c] / C.wt[c];              // scratch: y / wt
            v0q[c] = v0[c] / C.wt[c];
        }
        double ynrm = sqrt(wdotq<R>(v1q, v1q));
        const double sqrrmc = sqrt(epsneg);
        C.scale = ynrm * sqrrmc;
        if (C.scale == 0.0) {
            ynrm = sqrt(wdotq<R>(v0q, v0q));
            C.scale = ynrm * sqrrmc;
            if (C.scale == 0.0) { stif = -1; break; }
        }
        double v0v0 = wdotq<R>(v0q, v0q);
        if (v0v0 == 0.0) {
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                const bool ok = slot_ok<R>(c, (int)(threadIdx.x & 31));
                v0[c] = ok ? 1.0 : 0.0;
                v0q[c] = ok ? 1.0 / C.wt[c] : 0.0;
            }
            v0v0 = wdotq<R>(v0q, v0q);
        }
        const double v0nrm = sqrt(v0v0);
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            v0[c] /= v0nrm;
            v0q[c] = v0[c] / C.wt[c];
        }
        v0v0 = 1.0;
        double rold = 0.0;
        bool converged = false, early = false;
#pragma unroll 1
        for (int ntry = 0; ntry < 8; ++ntry) {
            const double v1v1 = stiff_jac_times<R, SMALL>(C, v0, v0v0, v1, v1q);
            ++nfev;
            if (sqrt(v1v1) > 1.0e10 * sqrt(v0v0)) { stif = -1; rootre = -1; early = true; break; }
            const double v0v1 = wdotq<R>(v0q, v1q);
            if (ntry == 0) {
                rold = v0v1 / v0v0;
                if (fabs(rold) < cbrt(epsneg)) { stif = 0; rootre = -1; early = true; break; }
            } else {
                if (stiff_dominant_real(v1v1, v0v1, v0v0, rold, rho, root1, root2)) { rootre = 1; converged = true; break; }
                rootre = 0;
            }
            const double v2v2 = stiff_jac_times<R, SMALL>(C, v1, v1v1, v2, v2q);
            ++nfev;
            const double v0v2 = wdotq<R>(v0q, v2q), v1v2 = wdotq<R>(v1q, v2q);
            if (stiff_dominant_real(v2v2, v1v2, v1v1, rold, rho, root1, root2)) { rootre = 1; converged = true; break; }
            rootre = 0;
            const double det1 = v0v0 * v1v1 - v0v1 * v0v1;
            const double alpha1 = (-v0v0 * v1v2 + v0v1 * v0v2) / det1;
            const double beta1 = (v0v1 * v1v2 - v1v1 * v0v2) / det1;
            const double v3v3 = stiff_jac_times<R, SMALL>(C, v2, v2v2, v3, v3q);
            ++nfev;
            const double v1v3 = wdotq<R>(v1q, v3q), v2v3 = wdotq<R>(v2q, v3q);
            if (stiff_dominant_real(v3v3, v2v3, v2v2, rold, rho, root1, root2)) { rootre = 1; converged = true; break; }
            const double det2 = v1v1 * v2v2 - v1v2 * v1v2;
            const double alpha2 = (-v1v1 * v2v3 + v1v2 * v1v3) / det2;
            const double beta2 = (v1v2 * v2v3 - v2v2 * v1v3) / det2;
            const double res2 = fabs(v3v3 + v2v2 * (alpha2 * alpha2) + v1v1 * (beta2 * beta2) +
                                     2 * v2v3 * alpha2 + 2 * v1v3 * beta2 + 2 * v1v2 * alpha2 * beta2);
            if (res2 <= 1e-6 * v3v3) {
                double r1[2], r2[2];
                stiff_quadratic_roots(alpha1, beta1, r1, r2);
                stiff_quadratic_roots(alpha2, beta2, root1, root2);
                rho = sqrt(root1[0] * root1[0] + root1[1] * root1[1]);
                const double D1 = (root1[0] - r1[0]) * (root1[0] - r1[0]) + (root1[1] - r1[1]) * (root1[1] - r1[1]);
                const double D2 = (root1[0] - r2[0]) * (root1[0] - r2[0]) + (root1[1] - r2[1]) * (root1[1] - r2[1]);
                if (sqrt(fmin(D1, D2)) <= 0.001 * rho) { converged = true; break; }
            }
            const double v3nrm = sqrt(v3v3);
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                v0[c] = v3[c] / v3nrm;
                v0q[c] = v0[c] / C.wt[c];
            }
            v0v0 = 1.0;
        }
        if (early) break;
        if (!converged) { stif = -1; rootre = -1; break; }
        have_root = true;
        stif = -1;
    } while (false);
    int flags = 0;
    if (have_root) {                                   // common.py:424-455
        rootre = root1[1] == 0.0 ? 1 : 0;
        if (root1[0] > 0.0) {
            stif = 0;
        } else {
            const double rho2 = sqrt(root2[0] * root2[0] + root2[1] * root2[1]);
            if (rho2 >= 0.9 * rho && root2[0] > 0.0) stif = 0;
            else if (fabs(root1[1]) > fabs(root1[0]) * tanang) stif = -1;
            else stif = rho >= 0.9 * stbrad ? 1 : 0;
        }
    }
    if (stif < 0) flags = (rootre == 0 && lotsfl) ? STIFF_OSCILLATORY : 0;
    else if (stif == 1 && rootre >= 0) flags = rootre ? STIFF_REAL : STIFF_COMPLEX;
    return flags | (nfev << 8);
}

// the copy inside the persistent kernel: out of line (see rk_persistent_body)
template <class R>
__device__ __noinline__ int stiff_probe_dev(const double* slot, long long stride, double xend,
                                            int maxfcn, int cost, double stbrad,
                                            double tanang) {
    return stiff_probe_impl<R, true>(slot, stride, xend, maxfcn, cost, stbrad, tanang);
}

// scipy/optimize/Zeros/brentq.c with xtol = rtol = 4 eps and 100 iterations, as
// solve_event_equation calls it (scipy/integrate/_ivp/ivp.py).  f(x) is any
// callable; shared by the Runge-Kutta and the SWAG lanes.
template <class F>
__device__ __forceinline__ double brentq_dev(F&& f, double xa, double xb) {
    const double tol = 4.0 * 0x1.0p-52;
    double xpre = xa, xcur = xb;
    double xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = f(xpre);
    double fcur = f(xcur);
    if (fpre == 0.0) return xpre;
    if (fcur == 0.0) return xcur;
    auto neg = [](double v) { return __double2hiint(v) < 0; };
    if (neg(fpre) == neg(fcur)) return xcur;     // scipy raises; cannot happen after
                                                 // find_active_events up to rounding
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0.0 && fcur != 0.0 && neg(fpre) != neg(fcur)) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (tol + tol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0.0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < pymin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur);
    }
    return xcur;
}

template <bool ON>
struct CkExtra {
    double tw[2], q[2];
};
template <>
struct CkExtra<false> {};

// ---- one trajectory ---------------------------------------------------------
#ifdef XSQ_EVENTS_N
// ---- event queue: chunked allocation ------------------------------------------
// word = generation << 20 | records taken from the current chunk (a 32-bit word:
// native shared-memory atomics); chunk[g & 15] is the chunk of generation g (-1:
// the queue is exhausted).  At most blockDim.x appends overshoot a full chunk, so
// the count never reaches 2^20.
// Synchronisation is through `word` alone: the installer writes chunk[g + 1], fences,
// and publishes generation g + 1 with an atomic; a thread reads chunk[g] only after
// its own atomic on `word` returned generation g.  (compute-sanitizer's racecheck
// models barriers, not atomics, and reports this write / read pair as a hazard.)
// A slot is rewritten 16 generations = 8192 appends of the CTA later.
struct EvqShared {
    unsigned word;
    long long chunk[16];
};
__device__ __forceinline__ EvqShared& evq_shared() {
    __shared__ EvqShared s;
    return s;
}
// thread 0 of the CTA, before a __syncthreads(): generation 0 is "full", so the
// first append installs the first chunk
__device__ __forceinline__ void evq_cta_init() {
    EvqShared& s = evq_shared();
    s.word = (unsigned)kEvqChunk;
    for (int i = 0; i < 16; ++i) s.chunk[i] = -1;
}
// The rare part of an append: the chunk is full.  Exactly one thread (the one
// that drew i == kEvqChunk) installs the next chunk and takes its record 0; the
// others wait for the new generation and draw again.
__device__ __noinline__ long long evq_alloc_slow(unsigned* evq_fill, unsigned long long* evq_count,
                                                 long long n_chunks, unsigned old) {
    EvqShared& s = evq_shared();
    for (;;) {
        const unsigned g = old >> 20, i = old & 0xfffffu;
        if (i < (unsigned)kEvqChunk) {
            const long long c = *(volatile long long*)&s.chunk[g & 15u];
            return c < 0 ? -1 : c * kEvqChunk + (long long)i;
        }
        if (i == (unsigned)kEvqChunk) {
            const long long prev = *(volatile long long*)&s.chunk[g & 15u];
            if (prev >= 0) atomicMax(evq_fill + prev, (unsigned)kEvqChunk);
            const unsigned long long c = atomicAdd(evq_count, 1ull);
            const long long cc = c < (unsigned long long)n_chunks ? (long long)c : -1;
            *(volatile long long*)&s.chunk[(g + 1u) & 15u] = cc;
            __threadfence_block();
            atomicExch(&s.word, (((g + 1u) & 0xfffu) << 20) | (cc >= 0 ? 1u : 0u));
            return cc < 0 ? -1 : cc * kEvqChunk;
        }
        while ((*(volatile unsigned*)&s.word >> 20) == g) __nanosleep(40);
        old = atomicAdd(&s.word, 1u);
    }
}
// Index of a free record, or -1 when the queue is exhausted.
__device__ __forceinline__ long long evq_alloc(const RkDev& P) {
    EvqShared& s = evq_shared();
    const unsigned old = atomicAdd(&s.word, 1u);
    const unsigned i = old & 0xfffffu;
    if (i < (unsigned)kEvqChunk) {
        const long long c = *(volatile long long*)&s.chunk[(old >> 20) & 15u];
        if (c >= 0) return c * kEvqChunk + (long long)i;
    }
    return evq_alloc_slow(P.evq_fill, P.evq_count, P.evq_cap / kEvqChunk, old);
}
// lane 0 of every warp when it leaves the persistent loop: the counter only
// grows, so the last warp publishes the fill of the CTA's last chunk
__device__ __forceinline__ void evq_cta_publish(const RkDev& P) {
    EvqShared& s = evq_shared();
    const unsigned w = *(volatile unsigned*)&s.word;
    const long long c = *(volatile long long*)&s.chunk[(w >> 20) & 15u];
    const unsigned i = w & 0xfffffu;
    const unsigned n = i < (unsigned)kEvqChunk ? i : (unsigned)kEvqChunk;
    if (c >= 0) atomicMax(P.evq_fill + c, n);
}
// One record: NF consecutive doubles (16-byte aligned, 128-bit accesses):
//   0 trajectory   1 event k | cubic << 8 | output slot << 32   2 t_old   3 t_new   4 h
//   5.. y_old[NL], y_new[NL], K[0..S][NL]
template <int S, int NL>
struct EvqRecord {
    static constexpr int RAW = 5 + (S + 3) * NL;
    static constexpr int NF = (RAW + 1) & ~1;
};
template <int S, int NL, int KR>
__device__ __forceinline__ void evq_write(const RkDev& P, long long idx, long long sys, int k,
                                          int slot, bool cubic, double t, double t_new, double h,
                                          const double (&y)[NL], const double (&y_new)[NL],
                                          const double (&K)[KR][NL]) {
    constexpr int NF = EvqRecord<S, NL>::NF;
    double rec[NF];
    rec[0] = __longlong_as_double(sys);
    rec[1] = __longlong_as_double((long long)k | ((long long)(cubic ? 1 : 0) << 8) |
                                  ((long long)slot << 32));
    rec[2] = t;
    rec[3] = t_new;
    rec[4] = h;
#pragma unroll
    for (int c = 0; c < NL; ++c) {
        rec[5 + c] = y[c];
        rec[5 + NL + c] = y_new[c];
    }
#pragma unroll
    for (int i = 0; i <= S; ++i)
#pragma unroll
        for (int c = 0; c < NL; ++c) rec[5 + (2 + i) * NL + c] = K[i][c];
    if constexpr ((EvqRecord<S, NL>::RAW & 1) != 0) rec[NF - 1] = 0.0;
    double2* q = reinterpret_cast<double2*>(P.evq + idx * NF);
#pragma unroll
    for (int j = 0; j < NF / 2; ++j) q[j] = make_double2(rec[2 * j], rec[2 * j + 1]);
}
#endif

template <class Tab, class R>
struct Lane {
    static constexpr int S = Tab::S;
    static constexpr int NL = R::NL;
    // rows of K: S+1, plus BS5's extra stages for the low/best interpolants
    static constexpr int KROWS = (Tab::VARIANT == tab::BS5V) ? S + 4 : S + 1;

    long long sys;          // trajectory index
    double t, h_abs, h_prev, l2_old, max_factor, min_step;  // l2_old: log2_core(ss) of the last accepted step
    double y[NL], f[NL], prm[R::NPL];
    // nfev holds only the evaluations made outside the hot loop (f0, h_start,
    // BS5 extra stages); the per-attempt evaluations are added in store() from
    // the step counters, so the hot loop maintains no evaluation counter.
    int n_acc, n_rej, n_pre, nfev, ieval;
    // Stiffness diagnosis state lives in shared memory, not in registers: it is
    // touched once per accepted step, while every register of the lane is
    // needed for the stage vectors (the kernel sits at its register cap).
    // The reference's okstp equals n_acc + 1 at the time of the check (the
    // diagnosis is off for forced step sequences) and jflstp is n_rej minus
    // its value at the last reset, so rejected attempts touch nothing here:
    //   hot.cnt        accepted steps until okstp == 20 or okstp % 40 == 39
    //   hot.next_many  the next okstp with okstp % many_steps == many_steps - 1
    //   rej_base       n_rej when jflstp was last set to 0
    //   bits           [7:8] probes waiting in the slots, [9:11] STIFF_* flags
    static constexpr unsigned SB_PEND1 = 1u << 7, SB_PEND2 = 1u << 8, SB_FLAG_SHIFT = 9;
    struct alignas(16) StiffHot {
        double havg;
        int next_many, cnt;
    };
    struct StiffState {
        StiffHot hot[XSQ_MAX_BLOCK];
        int rej_base[XSQ_MAX_BLOCK];
        unsigned bits[XSQ_MAX_BLOCK];
    };
    static __device__ __forceinline__ StiffState& stiff_state() {
        __shared__ StiffState s;
        return s;
    }
    bool standard_sc, fresh, step_rejected;
    // CKdisc's twiddle / quit factors (cash.py:241-243).  An empty member for
    // every other method: the lane is copied as a whole around the stiffness
    // probes, so an unused member would still cost its registers.
    CkExtra<Tab::VARIANT == tab::CKDISCV> ck;
#ifdef XSQ_EVENTS_N
    double ev_g[XSQ_EVENTS_N];     // event function values at (t, y)  (ivp.py `g`)
    int ev_n[XSQ_EVENTS_N];        // occurrences so far               (`event_count`)
#endif

    // RungeKutta.__init__, common.py:187-220
    __device__ __forceinline__ void init(const RkDev& P, long long idx,
                                         int lane) {
        sys = idx;
        t = P.t0;
#pragma unroll
        for (int k = 0; k < NL; ++k)
            y[k] = load_slot<R>(P.y0, P.n_lanes, idx, k, lane);
        R::load_params(P.params, idx, P.n_lanes, lane, prm);
        n_acc = n_rej = n_pre = ieval = 0;
        StiffState& ss = stiff_state();
        // probes of the thread's previous trajectory may still wait in the slots
        ss.bits[threadIdx.x] &= SB_PEND1 | SB_PEND2;
        ss.hot[threadIdx.x].havg = 0.0;
        ss.hot[threadIdx.x].next_many = P.stiff_many_steps > 1 ? P.stiff_many_steps - 1 : 1;
        ss.hot[threadIdx.x].cnt = 20;
        ss.rej_base[threadIdx.x] = 0;
#pragma unroll
        for (int k = 0; k < NL; ++k)
            f[k] = load_slot<R>(P.init_f0, P.n_lanes, idx, k, lane);
        nfev = P.init_nfev[idx];
        standard_sc = true;
        fresh = true;
        step_rejected = false;
        max_factor = kMaxFactor0;
        h_prev = 0.0;
        l2_old = 0.0;
        min_step = 0.0;
        if constexpr (Tab::VARIANT == tab::CKDISCV) {
            ck.tw[0] = 1.5;
            ck.tw[1] = 1.1;
            ck.q[0] = ck.q[1] = 100.0;
        }
#ifdef XSQ_EVENTS_N
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            ev_g[k] = user_event(k, t, y, prm);        // g = [event(t0, y0)]
            ev_n[k] = 0;
        }
#endif
        // first step: forced table, first_step, or Watts' h_start (ens_init)
        if (P.n_forced > 0) h_abs = P.h_forced[0];
        else if (P.first_step > 0.0) h_abs = P.first_step;
        else h_abs = P.init_h[idx];
    }

    // _reassess_stepsize, common.py:310-331.  attempt() tests the common case
    // (min_step <= h_abs <= max_step and t_bound at least two steps away)
    // with three independent compares and only then calls this slow path, so
    // the first stage of a step does not wait for a chain of selects.
    __device__ __forceinline__ void reassess_slow(const RkDev& P, double d) {
        if (h_abs < min_step || h_abs > P.max_step) {
            h_abs = pymin(P.max_step, pymax(min_step, h_abs));
            standard_sc = true;
        }
        if (d < 2.0 * h_abs) {
            if (d > h_abs) {
                h_abs = pymax(0.5 * d, min_step);
                standard_sc = true;
            } else {
                h_abs = d;
            }
        }
    }

    // K_i = f(t + c_i h, y + h * sum_j a_ij K_j), common.py:353-356
    template <int I>
    __device__ __forceinline__ void stage(double (&K)[KROWS][NL], double h) {
        double ys[NL];
        if constexpr (Tab::VARIANT == tab::NYSTROMV) {
            // RungeKuttaNystrom._rk_stage, common.py:1279-1285.  Slots [0, NH) of a
            // thread are positions, [NH, NL) their velocities; rows of K are whole
            // derivative vectors of which only the acceleration half is read.
            constexpr int NH = NL / 2;
            const double dt = Tab::cv(I) * h, hh = h * h;
#pragma unroll
            for (int c = 0; c < NH; ++c) {
                double au = 0.0, av = 0.0;
                bool fu = true, fv = true;
#pragma unroll
                for (int j = 0; j < I; ++j) {
                    if (Tab::a(I, j) != 0.0) {
                        const double a = Tab::av(I, j);
                        au = fu ? a * K[j][NH + c] : fma(a, K[j][NH + c], au);
                        fu = false;
                    }
                    if (Tab::ap(I, j) != 0.0) {
                        const double a = Tab::apv(I, j);
                        av = fv ? a * K[j][NH + c] : fma(a, K[j][NH + c], av);
                        fv = false;
                    }
                }
                ys[c] = y[c] + (au * hh + dt * y[NH + c]);
                ys[NH + c] = y[NH + c] + av * h;
            }
            R::f(t + dt, ys, prm, K[I]);
            return;
        }
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < I; ++j) {
                if (Tab::a(I, j) != 0.0) {     // structure: compile time
                    const double a = Tab::av(I, j);   // value: constant bank
                    acc = first ? a * K[j][c] : fma(a, K[j][c], acc);
                    first = false;
                }
            }
            ys[c] = fma(h, acc, y[c]);
        }
        R::f(__dadd_rn(t, __dmul_rn(Tab::cv(I), h)), ys, prm, K[I]);
    }
    template <int I, int END>
    __device__ __forceinline__ void stages(double (&K)[KROWS][NL], double h) {
        if constexpr (I < END) {
            stage<I>(K, h);
            stages<I + 1, END>(K, h);
        }
    }

    // sum_c (err_c / scale_c)^2 with scale = atol + rtol*max(|y|,|y_ref|)
    // (common.py:57-61, 338-339).  The RMS norm is sqrt(ss / n); the kernel
    // works with ss directly:  error_norm < 1  <=>  ss < n  holds EXACTLY in
    // IEEE arithmetic (sqrt and the division by n are monotone and correctly
    // rounded, and every double below n is also below n*(1 - 2^-54)), and
    // log2(error_norm) = (log2(ss) - log2(n)) / 2 feeds the controller.
    __device__ __forceinline__ double scaled_ss(
        const RkDev& P, const double (&errv)[NL], const double (&yref)[NL],
        int lane) {
        double ss = 0.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            // max(|y|, |yref|): pick the operand, |.| is a free modifier of the fma
            const double big = fabs(yref[c]) > fabs(y[c]) ? yref[c] : y[c];
            const double scale = fma(P.rtol, fabs(big), atol_of<R>(P, c, lane));
            const double q = errv[c] * rcp_scale(scale);
            ss = fma(q, q, ss);
        }
        return sys_sum<R::WARP>(ss);
    }

    // Dense output over the step just accepted: emit every t_eval point in
    // (t_old, t_new] (ivp.py:711-728).  K holds all stages of the step.
    __device__ void emit(const RkDev& P, double (&K)[KROWS][NL], double h,
                         double t_new, const double (&y_new)[NL], int lane) {
        if (ieval >= P.n_eval) return;
        if (P.direction * (P.t_eval[ieval] - t_new) > 0.0) return;
        if constexpr (Tab::NPOL == 0) emit_cubic(P, K, t_new, y_new, lane);
        else emit_poly(P, K, h, t_new, y_new, lane);
    }

    // CubicDenseOutput, common.py:793-821 (tableaux without P)
    __device__ void emit_cubic(const RkDev& P, double (&K)[KROWS][NL],
                               double t_new, const double (&y_new)[NL],
                               int lane) {
        const double hh = t_new - t;
        double te = P.t_eval[ieval];
        do {
            const double x = (te - t) / hh;
            const double omx = 1.0 - x;
            const double h00 = (1.0 + 2.0 * x) * (omx * omx);
            const double h10 = x * (omx * omx) * hh;
            const double h01 = (x * x) * (3.0 - 2.0 * x);
            const double h11 = (x * x) * (x - 1.0) * hh;
            double out[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c)
                out[c] = ((h00 * y[c] + h10 * K[0][c]) + h01 * y_new[c]) +
                         h11 * K[S][c];
            eval_put<R>(P, sys, lane, ieval, out);
            ++ieval;
            if (ieval >= P.n_eval) break;
            te = P.t_eval[ieval];
        } while (P.direction * (te - t_new) <= 0.0);
    }

    __device__ void emit_poly(const RkDev& P, double (&K)[KROWS][NL], double h,
                              double t_new, const double (&y_new)[NL],
                              int lane) {
        double te = P.t_eval[ieval];
        constexpr int NQ = (Tab::VARIANT == tab::BS5V) ? 6
                         : (Tab::NPOL > 0 ? Tab::NPOL : 1);
        double Q[NQ][NL];
        int npol = Tab::NPOL;
        double t_anchor = t, h_anchor = t_new - t;
        bool anchor_end = false;
        if constexpr (Tab::VARIANT == tab::BS5V) {
            if (P.interpolant == IP_FREE) {
                form_q_free(K, Q);
            } else if (P.interpolant == IP_LOW) {
                bs5_low(K, Q, h);
                npol = Tab::NPOL_LOW;
            } else {
                bs5_best(K, Q, h, y_new);
                npol = Tab::NPOL_BEST;
                anchor_end = true;
                t_anchor = t_new;
                h_anchor = (t_new + h) - t_new;
            }
        } else {
            form_q_free(K, Q);
        }
#pragma unroll
        for (int k = 0; k < NQ; ++k)
#pragma unroll
            for (int c = 0; c < NL; ++c) Q[k][c] *= h_anchor;
        do {
            const double x = (te - t_anchor) / h_anchor;
            double out[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double v = 0.0;
                // Horner, common.py:781-785: y = Q[-1]*x; (y += q; y *= x)...
#pragma unroll
                for (int k = NQ - 1; k >= 0; --k) {
                    if (k < npol) {
                        v = (k == npol - 1) ? Q[k][c] * x : (v + Q[k][c]) * x;
                    }
                }
                out[c] = v + (anchor_end ? y_new[c] : y[c]);
            }
            eval_put<R>(P, sys, lane, ieval, out);
            ++ieval;
            if (ieval >= P.n_eval) break;
            te = P.t_eval[ieval];
        } while (P.direction * (te - t_new) <= 0.0);
    }

    // Q = K.T @ P, common.py:363
    template <int NQ>
    __device__ __forceinline__ void form_q_free(double (&K)[KROWS][NL],
                                                double (&Q)[NQ][NL]) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double acc = 0.0;
                if (k < Tab::NPOL) {
#pragma unroll
                    for (int i = 0; i <= S; ++i) {
                        if (Tab::p(i, k) != 0.0)
                            acc = fma(Tab::pv(i, k), K[i][c], acc);
                    }
                }
                Q[k][c] = acc;
            }
        }
    }

    // BS5 extra stage r (row S+1+R_), bogacki.py:356-361, 366-369
    template <int R_>
    __device__ __forceinline__ void bs5_extra_stage(double (&K)[KROWS][NL],
                                                    double h) {
        if constexpr (Tab::VARIANT == tab::BS5V) {
            constexpr int ROW = S + 1 + R_;
            double ys[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < ROW; ++j) {
                    if (Tab::a_extra(R_, j) != 0.0)
                        acc = fma(Tab::a_extrav(R_, j), K[j][c], acc);
                }
                ys[c] = fma(acc, h, y[c]);
            }
            R::f(__dadd_rn(t, __dmul_rn(Tab::c_extrav(R_), h)), ys, prm,
                 K[ROW]);
            ++nfev;
        }
    }
    template <int NQ>
    __device__ void bs5_low(double (&K)[KROWS][NL], double (&Q)[NQ][NL],
                            double h) {
        if constexpr (Tab::VARIANT == tab::BS5V) {
            bs5_extra_stage<0>(K, h);
#pragma unroll
            for (int k = 0; k < NQ; ++k)
#pragma unroll
                for (int c = 0; c < NL; ++c) {
                    double acc = 0.0;
                    if (k < Tab::NPOL_LOW) {
#pragma unroll
                        for (int i = 0; i <= S + 1; ++i) {
                            if (Tab::plow(i, k) != 0.0)
                                acc = fma(Tab::plowv(i, k), K[i][c], acc);
                        }
                    }
                    Q[k][c] = acc;
                }
        }
    }
    // RKSuite's grouped summation, bogacki.py:372-388
    template <int NQ>
    __device__ void bs5_best(double (&K)[KROWS][NL], double (&Q)[NQ][NL],
                             double h, const double (&y_new)[NL]) {
        if constexpr (Tab::VARIANT == tab::BS5V) {
            bs5_extra_stage<0>(K, h);
            bs5_extra_stage<1>(K, h);
            bs5_extra_stage<2>(K, h);
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double kp[11];
                Q[0][c] = K[7][c];
#define XSQ_KP(col)                                                      \
    _Pragma("unroll") for (int i = 0; i < 11; ++i)                       \
        kp[i] = __dmul_rn(K[i][c], Tab::pbestv(i, col));
#define A2(a, b) __dadd_rn(a, b)
                XSQ_KP(1)
                Q[1][c] = A2(A2(A2(kp[4], A2(A2(kp[5], kp[7]), kp[0])),
                                A2(A2(kp[2], kp[8]), kp[9])),
                             A2(A2(kp[3], kp[10]), kp[6]));
                XSQ_KP(2)
                Q[2][c] = A2(A2(A2(kp[4], kp[5]),
                                A2(A2(A2(kp[2], kp[8]), A2(kp[9], kp[7])),
                                   kp[0])),
                             A2(A2(kp[3], kp[10]), kp[6]));
                XSQ_KP(3)
                Q[3][c] = A2(A2(A2(A2(kp[3], kp[7]), A2(kp[6], kp[5])), kp[4]),
                             A2(A2(A2(kp[9], kp[8]), A2(kp[2], kp[10])),
                                kp[0]));
                XSQ_KP(4)
                Q[4][c] = A2(A2(A2(kp[9], kp[8]), A2(A2(kp[6], kp[5]), kp[4])),
                             A2(A2(A2(kp[3], kp[7]), A2(kp[2], kp[10])),
                                kp[0]));
                XSQ_KP(5)
                Q[5][c] = A2(A2(kp[4], A2(A2(kp[9], kp[7]), A2(kp[6], kp[5]))),
                             A2(A2(A2(kp[3], kp[8]), A2(kp[2], kp[10])),
                                kp[0]));
#undef A2
#undef XSQ_KP
            }
        }
    }

#ifdef XSQ_EVENTS_N
    // ---- events: scipy's solve_ivp loop (ivp.py) on the device ---------------
    // The dense output of the step just accepted ("sol = solver.dense_output()"),
    // formed once per step and evaluated wherever the root finder asks.
    struct Dense {
        static constexpr int NQ = (Tab::VARIANT == tab::BS5V) ? 6
                                : (Tab::NPOL > 0 ? Tab::NPOL : 1);
        double Q[NQ][NL];
        double t_anchor, h_anchor;
        int npol;
        bool anchor_end, cubic, built;
    };
    template <class Cfg>
    __device__ void dense_build(const Cfg& P, Dense& D, double (&K)[KROWS][NL], double h,
                                double t_new, const double (&y_new)[NL], bool cubic) {
        D.built = true;
        D.cubic = cubic || Tab::NPOL == 0;
        D.npol = Tab::NPOL;
        D.t_anchor = t;
        D.h_anchor = t_new - t;
        D.anchor_end = false;
        if (D.cubic) return;
        if constexpr (Tab::VARIANT == tab::BS5V) {
            if (P.interpolant == IP_FREE) {
                form_q_free(K, D.Q);
            } else if (P.interpolant == IP_LOW) {
                bs5_low(K, D.Q, h);
                D.npol = Tab::NPOL_LOW;
            } else {
                bs5_best(K, D.Q, h, y_new);
                D.npol = Tab::NPOL_BEST;
                D.anchor_end = true;
                D.t_anchor = t_new;
                D.h_anchor = (t_new + h) - t_new;
            }
        } else {
            form_q_free(K, D.Q);
        }
#pragma unroll
        for (int k = 0; k < Dense::NQ; ++k)
#pragma unroll
            for (int c = 0; c < NL; ++c) D.Q[k][c] *= D.h_anchor;
    }
    // sol(te): HornerDenseOutput / CubicDenseOutput, common.py:766-821
    __device__ void dense_eval(const Dense& D, double (&K)[KROWS][NL], double t_new,
                               const double (&y_new)[NL], double te, double (&out)[NL]) {
        if (D.cubic) {
            const double hh = t_new - t;
            const double x = (te - t) / hh;
            const double omx = 1.0 - x;
            const double h00 = (1.0 + 2.0 * x) * (omx * omx);
            const double h10 = x * (omx * omx) * hh;
            const double h01 = (x * x) * (3.0 - 2.0 * x);
            const double h11 = (x * x) * (x - 1.0) * hh;
#pragma unroll
            for (int c = 0; c < NL; ++c)
                out[c] = ((h00 * y[c] + h10 * K[0][c]) + h01 * y_new[c]) + h11 * K[S][c];
            return;
        }
        const double x = (te - D.t_anchor) / D.h_anchor;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = Dense::NQ - 1; k >= 0; --k) {
                if (k < D.npol) v = (k == D.npol - 1) ? D.Q[k][c] * x : (v + D.Q[k][c]) * x;
            }
            out[c] = v + (D.anchor_end ? y_new[c] : y[c]);
        }
    }
    // the root of  tt -> event(k, tt, sol(tt))  in the step (ivp.py solve_event_equation)
    __device__ double event_root(const Dense& D, double (&K)[KROWS][NL], double t_new,
                                 const double (&y_new)[NL], int k) {
        return brentq_dev([&](double tt) {
            double ytmp[NL];
            dense_eval(D, K, t_new, y_new, tt, ytmp);
            return user_event(k, tt, ytmp, prm);
        }, t, t_new);
    }
    // The roots of the active events of this step, located now (scipy's
    // handle_events, ivp.py): counts, terminal decision, event records.
    struct EvCfg {
        int ev_terminal[XSQ_EVENTS_N];
        int ev_capacity, interpolant;
        double direction;
        double* t_events;
        double* y_events;
    };
    template <class Cfg>
    __device__ __forceinline__ void events_now(const Cfg& P, Dense& D, double (&K)[KROWS][NL],
                                               double h, double t_new, const double (&y_new)[NL],
                                               bool cubic, unsigned active, bool& terminate,
                                               double& t_stop) {
        double root[XSQ_EVENTS_N];
        {
            dense_build(P, D, K, h, t_new, y_new, cubic);
            bool any_term = false;
            double r_star = 0.0;
            // Each lane works through ITS OWN active events, so lanes that
            // solve for different event functions iterate together (the index k
            // is lane data, not a uniform loop variable): the warp pays for the
            // longest per-lane sequence, not for one root solve per function.
            for (unsigned todo = active; todo; todo &= todo - 1u) {
                const int k = __ffs(todo) - 1;
                const double r = event_root(D, K, t_new, y_new, k);
#pragma unroll
                for (int kk = 0; kk < XSQ_EVENTS_N; ++kk)
                    if (kk == k) root[kk] = r;
            }
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k) {
                if (!(active >> k & 1u)) continue;
                ++ev_n[k];
                if (P.ev_terminal[k] > 0 && ev_n[k] >= P.ev_terminal[k]) {
                    // handle_events: the first terminal root in time order
                    if (!any_term || P.direction * (root[k] - r_star) < 0.0) r_star = root[k];
                    any_term = true;
                }
            }
            terminate = any_term;
            if (terminate) t_stop = r_star;
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k) {
                if (!(active >> k & 1u)) continue;
                if (terminate && P.direction * (root[k] - r_star) > 0.0) continue;   // after the stop
                const int slot = ev_n[k] - 1;
                if (slot < P.ev_capacity) {
                    const long long base = (sys * XSQ_EVENTS_N + k) * P.ev_capacity + slot;
                    double ye[NL];
                    dense_eval(D, K, t_new, y_new, root[k], ye);
                    P.t_events[base] = root[k];
#pragma unroll
                    for (int c = 0; c < NL; ++c) P.y_events[base * NL + c] = ye[c];
                }
            }
        }
    }
    struct EvSlow {
        double K[S + 1][NL];
        double y[NL], y_new[NL], prm[R::NPL];
        double y_stop[NL];          // sol(t_stop) when a terminal event ends the trajectory
        double t, t_new, h, t_stop;
        long long sys;
        EvCfg cfg;
        int ev_n[XSQ_EVENTS_N];
        unsigned active;
        bool cubic, terminate;
    };
    template <int KR>
    static __device__ __forceinline__ void evslow_fill(EvSlow& a, const RkDev& P,
                                                       const double (&K)[KR][NL],
                                                       const double (&y)[NL],
                                                       const double (&y_new)[NL],
                                                       const double (&prm)[R::NPL],
                                                       const int (&ev_n)[XSQ_EVENTS_N], double t,
                                                       double t_new, double h, long long sys,
                                                       unsigned active, bool cubic) {
#pragma unroll
        for (int i = 0; i <= S; ++i)
#pragma unroll
            for (int c = 0; c < NL; ++c) a.K[i][c] = K[i][c];
#pragma unroll
        for (int c = 0; c < NL; ++c) { a.y[c] = y[c]; a.y_new[c] = y_new[c]; }
#pragma unroll
        for (int c = 0; c < R::NPL; ++c) a.prm[c] = prm[c];
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            a.ev_n[k] = ev_n[k];
            a.cfg.ev_terminal[k] = P.ev_terminal[k];
        }
        a.cfg.ev_capacity = P.ev_capacity;
        a.cfg.interpolant = P.interpolant;
        a.cfg.direction = P.direction;
        a.cfg.t_events = P.t_events;
        a.cfg.y_events = P.y_events;
        a.t = t; a.t_new = t_new; a.h = h; a.sys = sys;
        a.active = active; a.cubic = cubic;
    }
    static __device__ __noinline__ void events_slow(EvSlow& a) {
        Lane L;
        L.t = a.t;
        L.sys = a.sys;
        double K[KROWS][NL], y_new[NL];
#pragma unroll
        for (int i = 0; i <= S; ++i)
#pragma unroll
            for (int c = 0; c < NL; ++c) K[i][c] = a.K[i][c];
#pragma unroll
        for (int c = 0; c < NL; ++c) { L.y[c] = a.y[c]; y_new[c] = a.y_new[c]; }
#pragma unroll
        for (int c = 0; c < R::NPL; ++c) L.prm[c] = a.prm[c];
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) L.ev_n[k] = a.ev_n[k];
        Dense D;
        bool terminate = false;
        double t_stop = a.t_new;
        L.events_now(a.cfg, D, K, a.h, a.t_new, y_new, a.cubic, a.active, terminate, t_stop);
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) a.ev_n[k] = L.ev_n[k];
        a.terminate = terminate;
        a.t_stop = t_stop;
        if (terminate) {                       // y = sol(t) at the event (ivp.py)
            double ys[NL];
            L.dense_eval(D, K, a.t_new, y_new, t_stop, ys);
#pragma unroll
            for (int c = 0; c < NL; ++c) a.y_stop[c] = ys[c];
        }
    }

    static constexpr int EVQ_FIELDS = EvqRecord<S, NL>::NF;
    // The root of one queued step: what after_step does inside the lane.
    __device__ void evq_solve(const RkDev& P, long long idx) {
        double rec[EVQ_FIELDS];
        const double2* q = reinterpret_cast<const double2*>(P.evq + idx * EVQ_FIELDS);
#pragma unroll
        for (int j = 0; j < EVQ_FIELDS / 2; ++j) {
            const double2 v = q[j];
            rec[2 * j] = v.x;
            rec[2 * j + 1] = v.y;
        }
        sys = __double_as_longlong(rec[0]);
        const long long w = __double_as_longlong(rec[1]);
        const int k = (int)(w & 0xff), slot = (int)(w >> 32);
        const bool cubic = (w >> 8) & 1;
        t = rec[2];
        double t_new = rec[3];
        const double h = rec[4];
        double K[KROWS][NL], y_new[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            y[c] = rec[5 + c];
            y_new[c] = rec[5 + NL + c];
        }
#pragma unroll
        for (int i = 0; i <= S; ++i)
#pragma unroll
            for (int c = 0; c < NL; ++c) K[i][c] = rec[5 + (2 + i) * NL + c];
        R::load_params(P.params, sys, P.n_lanes, 0, prm);
        Dense D;
        dense_build(P, D, K, h, t_new, y_new, cubic);
        const double r = event_root(D, K, t_new, y_new, k);
        double ye[NL];
        dense_eval(D, K, t_new, y_new, r, ye);
        const long long base = (sys * XSQ_EVENTS_N + k) * P.ev_capacity + slot;
        P.t_events[base] = r;
#pragma unroll
        for (int c = 0; c < NL; ++c) P.y_events[base * NL + c] = ye[c];
    }

    // Everything solve_ivp does after solver.step() returned (ivp.py): events,
    // then the t_eval points of the step.  Returns true when a terminal event
    // ends the trajectory; t_new / y_new are then the event point.
    __device__ bool after_step(const RkDev& P, double (&K)[KROWS][NL], double h, double& t_new,
                               double (&y_new)[NL], int lane, bool cubic) {
        Dense D;
        D.built = false;
        double g_new[XSQ_EVENTS_N];
        unsigned active = 0;
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            g_new[k] = user_event(k, t_new, y_new, prm);
            const bool up = ev_g[k] <= 0.0 && g_new[k] >= 0.0;      // find_active_events
            const bool down = ev_g[k] >= 0.0 && g_new[k] <= 0.0;
            const int d = P.ev_direction[k];
            if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) active |= 1u << k;
        }
        bool terminate = false;
        double t_stop = t_new;
        // Deferred root location.  A root solve inside the lane runs while the
        // other 31 lanes of the warp wait (4 % of the steps carry an event on a
        // Poincare-section workload, so 3 of 4 warp iterations used to pay for
        // one).  When no active event of this step can be terminal at this
        // occurrence, nothing the lane does next depends on the root: it
        // reserves the output slot, appends the step (stages, end states) to
        // the event queue and goes on; event_queue_body locates the roots of all
        // queued steps afterwards, one thread per record, with the same
        // dense_build / event_root / dense_eval -- the same bits.  What does
        // not fit in the queue is solved here.
        if (active && P.evq_cap > 0 &&
            (Tab::VARIANT != tab::BS5V || P.interpolant == IP_FREE)) {
            bool may_end = false;
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k)
                if ((active >> k & 1u) && P.ev_terminal[k] > 0 && ev_n[k] + 1 >= P.ev_terminal[k])
                    may_end = true;
            if (!may_end) {
#pragma unroll
                for (int k = 0; k < XSQ_EVENTS_N; ++k) {
                    if (!(active >> k & 1u)) continue;
                    if (ev_n[k] >= P.ev_capacity) {        // no room in the output: count only
                        ++ev_n[k];
                        active &= ~(1u << k);
                        continue;
                    }
                    const long long idx = evq_alloc(P);
                    if (idx >= 0) {
                        evq_write<S, NL, KROWS>(P, idx, sys, k, ev_n[k], cubic, t, t_new, h, y,
                                                y_new, K);
                        ++ev_n[k];
                        active &= ~(1u << k);
                    }
                }
            }
        }
        if (active) {
            if constexpr (Tab::VARIANT == tab::BS5V) {
                // BS5's low / best interpolants evaluate extra stages of this lane
                events_now(P, D, K, h, t_new, y_new, cubic, active, terminate, t_stop);
            } else {
                // Rare (a terminal occurrence, or the queue is full): out of line, so
                // that the root finder's registers and code are not the hot loop's.
                // The stages are copied into a local-memory block; K itself stays in
                // registers.
                EvSlow a;
                evslow_fill(a, P, K, y, y_new, prm, ev_n, t, t_new, h, sys, active, cubic);
                events_slow(a);
#pragma unroll
                for (int k = 0; k < XSQ_EVENTS_N; ++k) ev_n[k] = a.ev_n[k];
                terminate = a.terminate;
                t_stop = a.t_stop;
            }
        }
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) ev_g[k] = g_new[k];
        // the t_eval points of (t_old, t] -- up to the event when it is terminal
        if (P.n_eval > 0 && ieval < P.n_eval &&
            P.direction * (P.t_eval[ieval] - t_stop) <= 0.0) {
            if (!D.built) dense_build(P, D, K, h, t_new, y_new, cubic);
            double te = P.t_eval[ieval];
            do {
                double out[NL];
                dense_eval(D, K, t_new, y_new, te, out);
                eval_put<R>(P, sys, lane, ieval, out);
                ++ieval;
                if (ieval >= P.n_eval) break;
                te = P.t_eval[ieval];
            } while (P.direction * (te - t_stop) <= 0.0);
        }
        if (terminate) {
            if (!D.built) dense_build(P, D, K, h, t_new, y_new, cubic);
            double ys[NL];
            dense_eval(D, K, t_new, y_new, t_stop, ys);          // y = sol(t)
#pragma unroll
            for (int c = 0; c < NL; ++c) y_new[c] = ys[c];
            t_new = t_stop;
        }
        return terminate;
    }
    __device__ void store_event_counts(const RkDev& P) {
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) P.ev_count[sys * XSQ_EVENTS_N + k] = ev_n[k];
    }
#endif

    // One ATTEMPT of a step (the body of `while not step_accepted`,
    // common.py:232-287).  Returns the lane status: LANE_RUNNING, or a final
    // code when the trajectory ends here.
    // FAST: adaptive stepping without t_eval output (the ensemble benchmark
    // path); the forced-step and dense-output tests on kernel parameters
    // vanish at compile time.
    template <bool FAST>
    __device__ __forceinline__ int attempt(const RkDev& P, int lane) {
        if constexpr (Tab::VARIANT == tab::CKDISCV) return attempt_ckdisc<FAST>(P, lane);
        else return attempt_rk<FAST>(P, lane);
    }

    // ---- CKdisc, cash.py:245-388 ------------------------------------------
    // y + h * sum_{i<NS} w_i K_i and h * sum_{i<NS} e_i K_i for one of the
    // embedded solutions (cash.py:390-404), then sum((err/tol)^2).
    // SEL: 0/1 assessment pair, 2/3 fallback pair, 4 the fifth order pair.
    template <int SEL>
    __device__ __forceinline__ double ck_solution(const RkDev& P, double (&K)[KROWS][NL],
                                                  double h, double (&sol)[NL], int lane) {
        if constexpr (Tab::VARIANT == tab::CKDISCV) {
            constexpr int NS = (SEL == 0 || SEL == 2) ? 2 : (SEL == 4 ? 6 : 4);
            double errv[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double sb = 0.0, se = 0.0;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    if constexpr (SEL < 2) {
                        if (Tab::b_assess(SEL, i) != 0.0) sb = fma(Tab::b_assessv(SEL, i), K[i][c], sb);
                        if (Tab::e_assess(SEL, i) != 0.0) se = fma(Tab::e_assessv(SEL, i), K[i][c], se);
                    } else if constexpr (SEL < 4) {
                        if (Tab::b_fallback(SEL - 2, i) != 0.0) sb = fma(Tab::b_fallbackv(SEL - 2, i), K[i][c], sb);
                        if (Tab::e_fallback(SEL - 2, i) != 0.0) se = fma(Tab::e_fallbackv(SEL - 2, i), K[i][c], se);
                    } else {
                        if (Tab::b(i) != 0.0) sb = fma(Tab::bv(i), K[i][c], sb);
                        if (Tab::e(i) != 0.0) se = fma(Tab::ev(i), K[i][c], se);
                    }
                }
                sol[c] = fma(h, sb, y[c]);
                errv[c] = h * se;
            }
            return scaled_ss(P, errv, sol, lane);
        } else {
            return 0.0;
        }
    }
    // norm(err/tol) ** (1/p) with norm = sqrt(ss/n): (ss/n) ** (1/(2p))
    static __device__ __forceinline__ double ck_root(const RkDev& P, double ss, double inv2p) {
        if (ss == 0.0) return 0.0;
        if (!(ss < XSQ_INF)) return ss;                      // inf / nan pass through
        return exp2_fast(inv2p * (log2_fast(ss) - P.log2n));
    }

    // One pass of `while not order_accepted` (cash.py:255-367).
    template <bool FAST>
    __device__ __forceinline__ int attempt_ckdisc(const RkDev& P, int lane) {
        if (fresh) {
            min_step = pymax(Tab::H_MIN_A * (fabs(t) + h_abs), XSQ_SQRT_TINY);
            const double d = fabs(P.t_bound - t);
            if ((h_abs < min_step) | (h_abs > P.max_step) | (d < 2.0 * h_abs)) reassess_slow(P, d);
            step_rejected = false;
            fresh = false;
        }
        if (h_abs < min_step) return LANE_TOO_SMALL;          // cash.py:257-258
        constexpr double NTOT = (double)R::N;
        double h = h_abs * P.direction;
        double K[KROWS][NL];
        double y_new[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) K[0][c] = f[c];
        stage<1>(K, h);
        ++nfev;
        // first order error, second order solution (cash.py:266-270)
        const double E1 = ck_root(P, ck_solution<0>(P, K, h, y_new, lane), 0.25);
        double esttol = E1 / ck.q[0];
        int accepted = 0;
        bool retried = false;
        if (E1 < ck.tw[0] * ck.q[0]) {
            stage<2>(K, h);
            stage<3>(K, h);
            nfev += 2;
            // second order error, third order solution (cash.py:278-282)
            const double E2 = ck_root(P, ck_solution<1>(P, K, h, y_new, lane), 1.0 / 6.0);
            esttol = E2 / ck.q[1];
            if (E2 < ck.tw[1] * ck.q[1]) {
                stage<4>(K, h);
                stage<5>(K, h);
                nfev += 2;
                double E4 = ck_root(P, ck_solution<4>(P, K, h, y_new, lane), 0.1);
                if (E4 == 0.0) E4 = 1e-160;                   // cash.py:293
                esttol = E4;
                if (E4 < 1.0) {
                    accepted = 4;                             // cash.py:297-317
                    double factor = pymin(Tab::MAX_FACTOR, Tab::SAFETY / E4);
                    if (step_rejected) factor = pymin(1.0, factor);
                    h_abs *= factor;
                    const double e12[2] = {E1, E2};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        double q = e12[j] / E4;
                        if (q > ck.q[j]) q = pymin(q, 10 * ck.q[j]);
                        else q = pymax(q, 2.0 / 3.0 * ck.q[j]);
                        ck.q[j] = pymax(1.0, pymin(10000.0, q));
                    }
                } else {
                    if (!(E4 < XSQ_INF)) return LANE_OVERFLOW;    // cash.py:320-321
                    const double e12[2] = {E1, E2};               // cash.py:324-328
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double EQ = e12[i] / ck.q[i];
                        if (EQ < ck.tw[i]) ck.tw[i] = pymax(1.1, EQ);
                    }
                    // third order fallback over 3/5 of the step (cash.py:331-341)
                    if (E2 < 1.0 && ck_solution<3>(P, K, h, y_new, lane) < NTOT) {
                        accepted = 2;
                        h_abs *= Tab::c_fallback(1);
                        h = h_abs * P.direction;
                    }
                }
            }
            // second order fallback over 1/5 of the step (cash.py:344-361)
            if (accepted == 0 && E1 < 1.0) {
                if (ck_solution<2>(P, K, h, y_new, lane) < NTOT) {
                    accepted = 1;
                    h_abs *= Tab::c_fallback(0);
                    h = h_abs * P.direction;
                } else {                                      // non-smooth: retry with h/5
                    step_rejected = true;
                    h_abs *= Tab::c_fallback(0);
                    ++n_rej;
                    retried = true;
                }
            }
        }
        if (accepted == 0) {
            if (!retried) {                                   // cash.py:363-367
                step_rejected = true;
                h_abs *= pymax(Tab::MIN_FACTOR, Tab::SAFETY / esttol);
                ++n_rej;
            }
            return (n_acc + n_rej >= P.max_steps) ? LANE_STEP_BUDGET : LANE_RUNNING;
        }
        // the accepted solution's derivative: first stage of the next step and
        // end slope of the interpolants (cash.py:371-375)
        const double t_new = t + h;
        R::f(t_new, y_new, prm, K[S]);
        ++nfev;
#ifdef XSQ_EVENTS_N
        double t_end = t_new;
        const bool ev_stop = after_step(P, K, h, t_end, y_new, lane, accepted != 4);
#else
        const double t_end = t_new;
        const bool ev_stop = false;
        if (!FAST && P.n_eval > 0 && ieval < P.n_eval &&
            P.direction * (P.t_eval[ieval] - t_new) <= 0.0) {
            // cash.py:406-416: Horner for the fifth order solution, else cubic
            if (accepted == 4) emit_poly(P, K, h, t_new, y_new, lane);
            else emit_cubic(P, K, t_new, y_new, lane);
        }
#endif
        t = t_end;
#pragma unroll
        for (int c = 0; c < NL; ++c) { y[c] = y_new[c]; f[c] = K[S][c]; }
        ++n_acc;
        fresh = true;
        if (ev_stop) return LANE_EVENT;
        if (P.direction * (t - P.t_bound) >= 0.0) return LANE_FINISHED;
        return (n_acc + n_rej >= P.max_steps) ? LANE_STEP_BUDGET : LANE_RUNNING;
    }

    template <bool FAST>
    __device__ __forceinline__ int attempt_rk(const RkDev& P, int lane) {
        // Forced-step mode shares every instruction of the adaptive path: h
        // comes from the table instead of reassess(), every step is accepted,
        // and the controller's h update is overwritten at the next step.
        const bool forced = !FAST && P.n_forced > 0;
        if (fresh) {
            if (forced) {
                h_abs = P.h_forced[n_acc];
            } else {
                min_step = pymax(Tab::H_MIN_A * (fabs(t) + h_abs), XSQ_SQRT_TINY);
                const double d = fabs(P.t_bound - t);
                if ((h_abs < min_step) | (h_abs > P.max_step) | (d < 2.0 * h_abs)) {
                    reassess_slow(P, d);
                    if (h_abs < min_step) return LANE_TOO_SMALL;   // common.py:234
                }
            }
            step_rejected = false;
            fresh = false;
        }
        const double h = h_abs * P.direction;
        const double t_new = t + h;
        double K[KROWS][NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) K[0][c] = f[c];

        constexpr bool EARLY = Tab::VARIANT == tab::BS5V || Tab::VARIANT == tab::CFMRV;
        constexpr bool NYSTROM = Tab::VARIANT == tab::NYSTROMV;
        constexpr int NFIRST = EARLY ? S - 1 : S;
        stages<1, NFIRST>(K, h);

        constexpr double NTOT = (double)R::N;
        double y_new[NL], errv[NL];
        double ss;                      // sum of squares of err/scale
        bool pre_reject = false;
        if constexpr (EARLY) {
            // pre-error from the first S-1 stages: bogacki.py:340-346 uses
            // (B_scale_pre, E_pre), calvo.py:255-261 uses (A[8,:8], E[:8])
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double sb = 0.0, se = 0.0;
#pragma unroll
                for (int i = 0; i < S - 1; ++i) {
                    if constexpr (Tab::VARIANT == tab::BS5V) {
                        if (Tab::b_scale_pre(i) != 0.0)
                            sb = fma(Tab::b_scale_prev(i), K[i][c], sb);
                        if (Tab::e_pre(i) != 0.0)
                            se = fma(Tab::e_prev(i), K[i][c], se);
                    } else {
                        if (Tab::a(S - 1, i) != 0.0)
                            sb = fma(Tab::av(S - 1, i), K[i][c], sb);
                        if (Tab::e(i) != 0.0)
                            se = fma(Tab::ev(i), K[i][c], se);
                    }
                }
                y_new[c] = fma(h, sb, y[c]);
                errv[c] = h * se;
            }
            ss = scaled_ss(P, errv, y_new, lane);
            // error_norm_pre > 1: certainly when ss exceeds n by a few ulp,
            // decided with the reference's own expression inside that band
            pre_reject = !forced && ss > NTOT &&
                (ss >= NTOT * (1.0 + 0x1.0p-48) || sqrt(ss / NTOT) > 1.0);
        }
        if (!pre_reject) {
            if constexpr (EARLY) stage<S - 1>(K, h);
            if constexpr (NYSTROM) {
                // RungeKuttaNystrom._comp_sol_err / _estimate_error, common.py:1287-1309
                constexpr int NH = NL / 2;
                const double hh = h * h;
#pragma unroll
                for (int c = 0; c < NH; ++c) {
                    double sb = 0.0, sp = 0.0;
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        if (Tab::b(i) != 0.0) sb = fma(Tab::bv(i), K[i][NH + c], sb);
                        if (Tab::bp(i) != 0.0) sp = fma(Tab::bpv(i), K[i][NH + c], sp);
                    }
                    y_new[c] = y[c] + (sb * hh + h * y[NH + c]);
                    y_new[NH + c] = y[NH + c] + sp * h;
                }
                if constexpr (Tab::FSAL) R::f(t_new, y_new, prm, K[S]);
#pragma unroll
                for (int c = 0; c < NH; ++c) {
                    double se = 0.0, sp = 0.0;
#pragma unroll
                    for (int i = 0; i < S + Tab::FSAL; ++i) {
                        if (Tab::e(i) != 0.0) se = fma(Tab::ev(i), K[i][NH + c], se);
                        if (Tab::ep(i) != 0.0) sp = fma(Tab::epv(i), K[i][NH + c], sp);
                    }
                    errv[c] = se * hh;
                    errv[NH + c] = sp * h;
                }
            } else {
            // _comp_sol_err, common.py:341-351
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double sb = 0.0;
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    if (Tab::b(i) != 0.0) sb = fma(Tab::bv(i), K[i][c], sb);
                }
                y_new[c] = fma(h, sb, y[c]);
            }
            if constexpr (Tab::FSAL) R::f(t_new, y_new, prm, K[S]);
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double se = 0.0;
#pragma unroll
                for (int i = 0; i < S + Tab::FSAL; ++i) {
                    if (Tab::e(i) != 0.0) se = fma(Tab::ev(i), K[i][c], se);
                }
                errv[c] = h * se;
            }
            }
            ss = scaled_ss(P, errv, y_new, lane);
        }

        // ---- controller, common.py:249-287 ---------------------------------
        // error_norm**x is exp2(x * log2(error_norm)).  Accepting and rejecting
        // lanes share ONE log2/exp2 (a warp almost always holds both), and
        // everything that does not depend on log2(ss) -- which formula, which
        // bounds -- is decided first, so the serial chain after the error norm
        // is log2 -> fma -> exp2 -> 2 selects -> h update.
        const bool accept = forced || (!pre_reject && ss < NTOT);
        const bool bad = !forced && !pre_reject && !(ss < XSQ_INF);  // NaN/Inf
        if (Tab::VARIANT == tab::BS5V && bad) {          // bogacki.py:314-315
            nfev += S - 1 + Tab::FSAL;        // this attempt is in no counter
            return LANE_OVERFLOW;
        }
        // Accepted lanes: f(t+h, y_new) of non-FSAL pairs (common.py:289-291)
        // and the stiffness bookkeeping (common.py:306) come BEFORE the
        // controller arithmetic, while the error vector is still live.
        bool flush = false;
        if (accept) {
            if constexpr (!Tab::FSAL)
                R::f(t_new, y_new, prm, K[S]);
            if (P.nfev_stiff_detect > 0 && !forced)
                flush = diagnose(P, K, errv, y_new, t_new, h);
        }
        const bool tiny_err = ss < NTOT * 0x1.0p-1022;       // err < sqrt(tiny)
        const bool second = accept && !standard_sc;          // 2nd-order SC
        const double l2ss = log2_core(ss);
        double factor;
        if (P.minalpha != 0.0) {          // user sc_params with an alpha term (uniform branch)
            const double z_extra = second ? P.minalpha * log2_fast(h / h_prev) : 0.0;
            factor = ctl_factor<true>(P, l2ss, l2_old, z_extra, accept, second, step_rejected,
                                      tiny_err, max_factor);
        } else {
            factor = ctl_factor<false>(P, l2ss, l2_old, 0.0, accept, second, step_rejected,
                                       tiny_err, max_factor);
        }
        if (bad || (pre_reject && !(ss < XSQ_INF))) factor = kMinFactor;   // max(0.2, nan)
        h_abs *= factor;
        if (!accept) {
            step_rejected = true;
            ++n_rej;
            if (EARLY && pre_reject) ++n_pre;
            // ++jflstp (common.py:284) is ++n_rej, see StiffState
            if (bad) return LANE_OVERFLOW;                   // common.py:286
            if (h_abs < min_step) return LANE_TOO_SMALL;     // common.py:234
            if (n_acc + n_rej >= P.max_steps) return LANE_STEP_BUDGET;
            return LANE_RUNNING;
        }
        standard_sc = tiny_err;
        if (factor < kMaxFactor) max_factor = kMaxFactor;
#ifdef XSQ_EVENTS_N
        double t_end = t_new;
        const bool ev_stop = after_step(P, K, h, t_end, y_new, lane, false);
#else
        const double t_end = t_new;
        const bool ev_stop = false;
        if (!FAST && P.n_eval > 0) emit(P, K, h, t_new, y_new, lane);
#endif
        // common.py:294-303
        h_prev = h;
        l2_old = l2ss;
        t = t_end;
#pragma unroll
        for (int c = 0; c < NL; ++c) { y[c] = y_new[c]; f[c] = K[S][c]; }
        ++n_acc;
        fresh = true;
        if (ev_stop) return LANE_EVENT;
        // OdeSolver.step, base.py:207-208
        const bool done = forced ? (n_acc >= P.n_forced)
                                 : (P.direction * (t - P.t_bound) >= 0.0);
        if (done) return LANE_FINISHED;
        if (n_acc + n_rej >= P.max_steps) return LANE_STEP_BUDGET;
        return flush ? LANE_FLUSH : LANE_RUNNING;
    }

    // _diagnose_stiffness, common.py:370-516, bookkeeping part.  Called after an
    // accepted step, before the state moves on: y is still the old state,
    // y_new / K[S] / errv belong to the new one.  The probe itself changes
    // nothing but nfev and the flags, so it is DEFERRED: its inputs go to a
    // slot in global memory and the persistent loop runs the probes of all
    // lanes of the warp together (flush_probes) instead of one lane at a time
    // with 31 lanes idle.
    // Returns true when both slots are taken: the warp must flush before this
    // lane's next accepted step.
    __device__ __forceinline__ bool diagnose(const RkDev& P, double (&K)[KROWS][NL],
                                             const double (&errv)[NL],
                                             const double (&y_new)[NL],
                                             double t_new, double h) {
        StiffState& ss = stiff_state();
        const double2 hot = *reinterpret_cast<const double2*>(&ss.hot[threadIdx.x]);
        double havg = c_xsq_havg[0] * hot.x + c_xsq_havg[1] * h;
        int next_many = __double2loint(hot.y), cnt = __double2hiint(hot.y) - 1;
        const bool toomch = n_acc + 1 == next_many;
        bool lotsfl = false;
        if (cnt == 0) {                       // okstp == 20 or okstp % 40 == 39
            if (n_acc == 19) {
                havg = h;
                cnt = 19;
            } else {
                lotsfl = n_rej - ss.rej_base[threadIdx.x] >= 10;
                cnt = 40;
            }
            ss.rej_base[threadIdx.x] = n_rej;                // jflstp = 0
        }
        if (toomch) next_many += P.stiff_many_steps;
        *reinterpret_cast<double2*>(&ss.hot[threadIdx.x]) =
            make_double2(havg, __hiloint2double(cnt, next_many));
        if (!(toomch || lotsfl)) return false;
        // ---- rare from here ----
        // Where the probe's inputs go: a record of the probe queue, worked off
        // by stiff_queue after this kernel (thread-per-system policies, while
        // the queue has room), else one of the thread's two slots.
        using SL = StiffSlot<R>;
        double* s = nullptr;
        long long stride = 1;
        if (!R::WARP && P.stiff_q_cap > 0) {
            const unsigned long long qi = atomicAdd(P.stiff_q_count, 1ULL);
            if (qi < (unsigned long long)P.stiff_q_cap) s = P.stiff_q + qi * SL::DOUBLES;
        }
        bool urgent = false;
        if (s == nullptr) {
            unsigned sbits = ss.bits[threadIdx.x];
            stride = P.stiff_threads;
            const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const int islot = (sbits & SB_PEND1) ? 1 : 0;
            s = P.stiff_slot + (long long)islot * SL::DOUBLES * stride + gtid;
            sbits += SB_PEND1;                               // 0 -> 1 -> 2
            urgent = (sbits & SB_PEND2) != 0u;
            ss.bits[threadIdx.x] = sbits;
        }
        s[0] = t_new;
        s[stride] = h;
        s[2 * stride] = havg;
        s[3 * stride] = lotsfl ? 1.0 : 0.0;
        s[4 * stride] = __longlong_as_double(sys);
        s += SL::HEAD * stride;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            s[c * stride] = y_new[c];
            s[(NL + c) * stride] = y[c];
            s[(2 * NL + c) * stride] = K[S][c];
            s[(3 * NL + c) * stride] = errv[c];
        }
#pragma unroll
        for (int c = 0; c < R::NPL; ++c) s[(4 * NL + c) * stride] = prm[c];
        return urgent;
    }

    static __device__ __forceinline__ bool probes_urgent() {
        return (stiff_state().bits[threadIdx.x] & SB_PEND2) != 0u;
    }
    static __device__ __forceinline__ bool probes_pending() {
        return (stiff_state().bits[threadIdx.x] & (SB_PEND1 | SB_PEND2)) != 0u;
    }

    // Runs the waiting probes of this thread.  A probe only adds to nfev and to
    // the flags of its trajectory: for the trajectory the thread is on now
    // (`cur`) they are returned / kept in shared memory, for one that has
    // already been stored they are applied to its results in global memory.
    // Touches no Lane member (see rk_persistent_body).
    static __device__ __forceinline__ int flush_probes(const RkDev& P, long long cur, int lane) {
        using SL = StiffSlot<R>;
        const long long stride = P.stiff_threads;
        const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned sbits = stiff_state().bits[threadIdx.x];
        int evals = 0;
        while (sbits & (SB_PEND1 | SB_PEND2)) {
            sbits -= SB_PEND1;
            const int islot = (sbits & SB_PEND1) ? 1 : 0;
            const double* slot = P.stiff_slot + (long long)islot * SL::DOUBLES * stride + gtid;
            const int r = stiff_probe_dev<R>(slot, stride, P.t_bound, P.nfev_stiff_detect, S,
                                             Tab::STBRAD, Tab::TANANG);
            const long long owner = __double_as_longlong(slot[4 * stride]);
            if (owner == cur) {
                evals += r >> 8;
                sbits |= (unsigned)(r & 7) << SB_FLAG_SHIFT;
            } else if (!R::WARP || lane == 0) {
                P.nfev[owner] += r >> 8;
                if (P.stiff_flags) P.stiff_flags[owner] |= r & 7;
            }
        }
        stiff_state().bits[threadIdx.x] = sbits;
        return evals;
    }

    // RHS evaluations made by attempt(): S-1 stages (+1 if FSAL) per full
    // attempt, one fewer stage and no FSAL evaluation for a pre-rejected
    // attempt (bogacki.py:266-275, calvo.py:178-187), and f(t+h, y_new) per
    // ACCEPTED step for non-FSAL pairs (common.py:289-291).
    __device__ __forceinline__ int evals_in_loop() const {
        if (Tab::VARIANT == tab::CKDISCV) return 0;   // counted where they are made
        const int attempts = n_acc + n_rej;
        return (S - 1 + Tab::FSAL) * (attempts - n_pre) + (S - 2) * n_pre +
               (Tab::FSAL ? 0 : n_acc);
    }

    // `constant`: zero-length span, every t_eval point equals t0 and scipy
    // returns y0 there (ConstantDenseOutput, base.py:224-226).
    __device__ __forceinline__ void store(const RkDev& P, int st, int lane,
                                          bool constant = false) {
#pragma unroll
        for (int k = 0; k < NL; ++k)
            store_slot<R>(P.y_final, P.n_lanes, sys, k, lane, y[k]);
        if (P.n_eval > 0 && ieval < P.n_eval) {
            eval_finish<R>(P, sys, lane, ieval, constant, y);
            if (constant) ieval = P.n_eval;
        }
        if (!R::WARP || lane == 0) {
            P.t_final[sys] = t;
            if (P.h_next) P.h_next[sys] = h_abs;
            P.n_acc[sys] = n_acc;
            P.n_rej[sys] = n_rej;
            P.nfev[sys] = nfev + evals_in_loop();
            P.status[sys] = st == LANE_EVENT ? 1 : st;    // 1: a termination event occurred
#ifdef XSQ_EVENTS_N
            store_event_counts(P);
#endif
            if (P.n_eval_done) P.n_eval_done[sys] = ieval;
            if (P.stiff_flags) P.stiff_flags[sys] =
                    (int)((stiff_state().bits[threadIdx.x] >> SB_FLAG_SHIFT) & 7u);
        }
    }
};

// ---- the persistent kernel ---------------------------------------------------
// Lane-per-system: every thread owns one trajectory; when it ends the thread
// stores it and claims the next index from the global queue (warp-aggregated
// atomicAdd), so divergence in step counts does not idle the warp.
// Warp-per-system: the claim is made once per warp and broadcast.
template <class Tab, class R>
__device__ __forceinline__ void rk_persistent_body(const RkDev& P) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    Lane<Tab, R> L;
    bool live = false;
    bool exhausted = false;
    const bool fast = P.n_forced == 0 && P.n_eval == 0;
#ifdef XSQ_EVENTS_N
    if (threadIdx.x == 0) evq_cta_init();                      // before the barrier below
#endif
    math_tabs_init();
    Lane<Tab, R>::stiff_state().bits[threadIdx.x] = 0u;
    // Stiffness probes wait in their slots until kProbeWindow attempts have
    // passed, so that one pass serves many lanes of the warp (a trajectory may
    // end, and the thread start the next one, while its probes wait).  The
    // probe is an out-of-line call in a kernel that sits at its register cap:
    // anything live across it would be spilled for the whole kernel, hot loop
    // included.  So the lane is parked in local memory by hand for the
    // duration of the (rare) pass and nothing but `live` crosses it.
    constexpr int kProbeWindow = 640;
    int it = 0;
    auto flush = [&](long long cur) {
        if (!__any_sync(full, Lane<Tab, R>::probes_pending())) return;
        if (Lane<Tab, R>::probes_pending()) {
            Lane<Tab, R> parked = L;
            memory_fence_for(&parked);
            const int evals = Lane<Tab, R>::flush_probes(P, cur, lane);
            memory_fence_for(&parked);
            L = parked;
            L.nfev += evals;
        }
        __syncwarp(full);
    };
    for (;;) {
        // ---- refill ----
        if (R::WARP) {
            if (!live && !exhausted) {
                unsigned long long idx = 0;
                if (lane == 0) idx = atomicAdd(P.queue, 1ULL);
                idx = __shfl_sync(full, idx, 0);
                if ((long long)idx < P.n_lanes) {
                    L.init(P, (long long)idx, lane);
                    live = true;
                    if (P.n_forced == 0 && P.t0 == P.t_bound) {
                        L.store(P, LANE_FINISHED, lane, true);
                        live = false;
                    }
                } else {
                    exhausted = true;
                }
            }
        } else {
            const unsigned need = __ballot_sync(full, !live && !exhausted);
            if (need) {
                unsigned long long base = 0;
                const int leader = __ffs(need) - 1;
                if (lane == leader)
                    base = atomicAdd(P.queue, (unsigned long long)__popc(need));
                base = __shfl_sync(full, base, leader);
                if (!live && !exhausted) {
                    const long long idx =
                        (long long)base + __popc(need & ((1u << lane) - 1u));
                    if (idx < P.n_lanes) {
                        L.init(P, idx, lane);
                        live = true;
                        // t0 == t_bound / zero-length span: scipy base.py:197
                        if (P.n_forced == 0 && P.t0 == P.t_bound) {
                            L.store(P, LANE_FINISHED, lane, true);
                            live = false;
                        }
                    } else {
                        exhausted = true;
                    }
                }
            }
            __syncwarp(full);
        }
        if (__all_sync(full, !live)) {
            // nothing to step: done when the queue is exhausted, else refill (lanes
            // that ended at once -- zero-length span -- must not end the warp)
            if (__all_sync(full, exhausted)) break;
            continue;
        }
        // ---- attempts, until some lane of the warp ends its trajectory ----
        // Leaves the loop when a lane ends or has both probe slots taken
        // (LANE_FLUSH), and after kProbeWindow attempts (see below).
        int st = LANE_RUNNING;
        if (fast) {
            do {
                if (live) st = L.template attempt<true>(P, lane);
            } while (!__any_sync(full, st != LANE_RUNNING) && ++it < kProbeWindow);
        } else {
            do {
                if (live) st = L.template attempt<false>(P, lane);
            } while (!__any_sync(full, st != LANE_RUNNING) && ++it < kProbeWindow);
        }
        const bool expired = it >= kProbeWindow;
        if (expired) it = 0;
        // both slots of some thread taken: run the probes now.  Asked of the slot
        // state, not of `st`: a trajectory may END on the very step that filled
        // its thread's second slot (st is then LANE_FINISHED), and the thread's
        // next trajectory must not find both slots occupied.
        if (P.nfev_stiff_detect > 0 &&
            (expired || __any_sync(full, Lane<Tab, R>::probes_urgent())))
            flush(live ? L.sys : -1);
        if (st == LANE_FLUSH) st = LANE_RUNNING;
        if (st != LANE_RUNNING) {
            L.store(P, st, lane);
            live = false;
        }
        __syncwarp(full);
    }
    if (P.nfev_stiff_detect > 0) flush(-1);      // probes of stored trajectories
#ifdef XSQ_EVENTS_N
    if (P.evq_cap > 0 && lane == 0) evq_cta_publish(P);
#endif
}

// Works off the probe queue: one thread per record, grid-stride (the record
// count is only known on the device).  A probe adds to nfev and to the flags of
// the trajectory that queued it; several records may belong to one trajectory.
template <class R>
__device__ __forceinline__ void stiff_queue_body(const RkDev& P, int cost, double stbrad,
                                                 double tanang) {
    if constexpr (!R::WARP) {
        unsigned long long n = *P.stiff_q_count;
        if (n > (unsigned long long)P.stiff_q_cap) n = (unsigned long long)P.stiff_q_cap;
        const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
             i < n; i += step) {
            const double* rec = P.stiff_q + i * StiffSlot<R>::DOUBLES;
            const int r = stiff_probe_impl<R, false>(rec, 1, P.t_bound, P.nfev_stiff_detect,
                                                     cost, stbrad, tanang);
            const long long owner = __double_as_longlong(rec[4]);
            if (r >> 8) atomicAdd(&P.nfev[owner], r >> 8);
            if ((r & 7) && P.stiff_flags) atomicOr(&P.stiff_flags[owner], r & 7);
        }
    }
}

#ifdef XSQ_EVENTS_N
// Works off the event queue: one thread per queued (step, event) pair.
template <class Tab, class R>
__device__ __forceinline__ void event_queue_body(const RkDev& P) {
    if constexpr (!R::WARP) {
        // chunks handed out, walked as tiles of blockDim.x records
        unsigned long long used = *P.evq_count;
        const unsigned long long n_chunks = (unsigned long long)(P.evq_cap / kEvqChunk);
        if (used > n_chunks) used = n_chunks;
        const long long tpc = kEvqChunk / (int)blockDim.x;          // launched with 128 threads
        for (long long tile = blockIdx.x; tile < (long long)used * tpc; tile += gridDim.x) {
            const long long c = tile / tpc;
            unsigned n = P.evq_fill[c];
            if (n > (unsigned)kEvqChunk) n = (unsigned)kEvqChunk;
            const unsigned i = (unsigned)(tile % tpc) * blockDim.x + threadIdx.x;
            if (i < n) {
                Lane<Tab, R> L;
                L.evq_solve(P, c * kEvqChunk + (long long)i);
            }
        }
    }
}
#endif

template <class R>
__global__ void __launch_bounds__(128, 6) stiff_queue(const RkDev P, int cost, double stbrad,
                                                   double tanang) {
    stiff_queue_body<R>(P, cost, stbrad, tanang);
}

template <class Tab, class R, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) rk_persistent(const RkDev P) {
    rk_persistent_body<Tab, R>(P);
}

#ifdef XSQ_TUNE
// profiling builds only: exact register cap instead of a CTAs-per-SM hint
template <class Tab, class R, int BLOCK, int MAXREG>
__global__ void __launch_bounds__(BLOCK) __maxnreg__(MAXREG)
    rk_persistent_mr(const RkDev P) {
    rk_persistent_body<Tab, R>(P);
}
#endif

}  // namespace xsq
