// xsq_rk_fast.cuh -- the ensemble hot path: adaptive stepping of a generic
// embedded pair (Ts5, CK5, Me4, Pr7, Pr8, Pr9), one trajectory per thread, final
// state only (no t_eval, no events, no forced steps, default step budget,
// controller presets without the alpha term).  Same arithmetic, operation by
// operation, as rk_persistent<Tab, R> in xsq_rk_core.cuh (the results of the
// two kernels are bit-identical; tests/test_gpu_rk.py asserts it) -- what
// differs is how the instructions are spent.
//
// Measured on B200 (profiles/r01_*): the persistent kernel is bound by
// instruction ISSUE, not by the depth of its dependency chains: an fp64
// instruction occupies its scheduler for two cycles and any other instruction
// for one, and   2 * (fp64 instructions) + (other instructions)   per attempted
// step reproduces the measured time within 2 % (749 predicted / 747 measured
// cycles per warp and attempt for Ts5/Lorenz).  So this kernel minimises
// instructions:
//   * tableau coefficients are streamed from shared memory in the order of
//     use, two per LDS.128; none lives in a uniform register, so the 40-odd
//     MOV.SPILL / R2UR.FILL / UMOV per attempt of the generic kernel are gone;
//   * accept / tiny / overflow are integer compares on the high word of the
//     sum of squares (exact: it is non-negative and the thresholds have a zero
//     low word);
//   * _reassess_stepsize (common.py:310-331) runs once per ACCEPTED step behind
//     three integer compares that prove "nothing to do" (h far from min_step,
//     max_step and t_bound), falling back to the exact code otherwise;
//     min_step itself is only formed when h_abs gets within reach of it;
//   * one exit per attempt: the status is a value, not a control-flow path, so
//     the loop-carried state is moved once;
//   * lane flags (standard_sc, step_rejected, max_factor == 4) are bits of one
//     register; the trajectory index is 32 bit.
// Reference (file:line in /root/reference/extensisq): common.py:222-356
// (_step_impl, _reassess_stepsize, _comp_sol_err, _rk_stage), :370-516
// (_diagnose_stiffness bookkeeping).
#pragma once
#ifndef __CUDACC_RTC__
#include <cmath>
#include <cstdlib>
#include <cstring>
#endif
#include "xsq_rk_core.cuh"

// build-time variants (measured on B200)
#ifndef XSQ_V_PREFETCH
#define XSQ_V_PREFETCH 0     // request the next coefficient row one stage ahead
#endif
#ifndef XSQ_V_INTDONE
#define XSQ_V_INTDONE 0      // "t reached t_bound" by sign bits instead of DMUL + DSETP
#endif

namespace xsq {

// compile-time loop: f(IC<B>{}), ..., f(IC<E-1>{})
template <int V>
struct IC { static constexpr int value = V; };
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- coefficient stream ------------------------------------------------------
// Rows A[1..S-1], B, E with their structural zeros removed, each row padded to
// an even number of doubles: row I starts at coef_base<Tab>(I), the k-th
// nonzero of a row sits at base + k.
template <class Tab>
struct CoefLayout {
    static constexpr int S = Tab::S;
    static constexpr int ROW_B = S, ROW_E = S + 1, ROW_C = S + 2, NROWS = S + 3;
    XSQ_HD static constexpr int row_len(int row) { return row < S ? row : (row == ROW_E ? S + Tab::FSAL : S); }
    XSQ_HD static constexpr double value(int row, int j) {
        return row < S ? Tab::a(row, j)
             : row == ROW_B ? Tab::b(j)
             : row == ROW_E ? Tab::e(j) : Tab::c(j);
    }
    XSQ_HD static constexpr int nnz(int row) {
        int n = 0;
        for (int j = 0; j < row_len(row); ++j) n += value(row, j) != 0.0 ? 1 : 0;
        return n;
    }
    XSQ_HD static constexpr int base(int row) {
        int b = 0;
        for (int r = 1; r < row; ++r) b += (nnz(r) + 1) & ~1;
        return b;
    }
    XSQ_HD static constexpr int index(int row, int j) {     // position of (row, j) in its row
        int k = 0;
        for (int jj = 0; jj < j; ++jj) k += value(row, jj) != 0.0 ? 1 : 0;
        return k;
    }
    XSQ_HD static constexpr int nnz_range(int r0, int r1) {
        int n = 0;
        for (int r = r0; r < r1; ++r) n += nnz(r);
        return n;
    }
    static constexpr int TOTAL = base(NROWS);
};

template <class Tab>
struct FastShared {
    double coef[CoefLayout<Tab>::TOTAL + 2];
};

// The loop's constants other than the tableau, one copy per CTA in shared
// memory (filled from the kernel parameters): read with LDS into vector
// registers where they are used, which keeps them out of the uniform registers.
struct LoopConsts {
    double lg_pol[6];        // log2 polynomial        (xsq_math_tables_gen.cuh)
    double e2_pol[6];        // exp2 polynomial, padded
    double ctl[6];           // a1s, a0s, a1c, a2c, a0c, pad (CtlConst)
    double havg[2];          // 0.9, 0.1               (common.py:372)
    double atol[16];
};

// 32-bit shared-memory addresses, formed once per thread and kept opaque so
// that they stay in registers (re-deriving one costs three instructions)
struct SmemAddr {
    SAddr coef, lg, e2;
    SAddr h0;      // this thread's "h_abs at the start of the step" word
    SAddr lc;      // LoopConsts
};
template <int N, int OFF_DOUBLES>
__device__ __forceinline__ void lc_load(SAddr lc, double (&v)[N]) {
    static_for<0, N / 2>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        lds2_stream<(OFF_DOUBLES + 2 * q) * 8>(lc, v[2 * q], v[2 * q + 1]);
    });
    if constexpr (N & 1) lds1_stream<(OFF_DOUBLES + N - 1) * 8>(lc, v[N - 1]);
}
__device__ __forceinline__ double log2_core_s(double x, SAddr lg, SAddr lc) {
    const SAddr a = lg + (unsigned)log2_tab_offset(x) * 8u;
    double inv, l_hi, l_lo, pol[6];
    lds2(a, inv, l_hi);
    lds1_at<16>(a, l_lo);
    lc_load<6, 0>(lc, pol);
    return log2_arith(x, inv, l_hi, l_lo, pol);
}
struct Exp2Shared {
    SAddr e2, lc;
    __device__ __forceinline__ double operator()(double z) const {
        const Exp2Split s = exp2_split(z);
        double t_hi, t_lo, pol[6];
        lds2(e2 + ((unsigned)(s.N & 63) << 4), t_hi, t_lo);
        lc_load<6, 6>(lc, pol);
        return exp2_arith(s, t_hi, t_lo, pol);
    }
};

// two consecutive doubles of the stream (OFF even: one LDS.128), or one
template <int OFF>
__device__ __forceinline__ void coef_ld2(SAddr base, double& a, double& b) {
    lds2_stream<OFF * 8>(base, a, b);
}
template <int OFF>
__device__ __forceinline__ void coef_ld1(SAddr base, double& a) {
    lds1_stream<OFF * 8>(base, a);
}

// Where a pair's coefficients are taken from.  An fp64 instruction costs
// max(2, distinct source REGISTERS) issue cycles (a 64-bit operand reads an
// even and an odd register, the file delivers one of each per cycle; measured:
// tools/sass_cost.py reproduces the kernel time within 4 %), so
//   acc = fma(coef, K, acc)   is 3 cycles with the coefficient in a register,
//                                2 cycles with it in a UNIFORM register.
// About 36 uniform doubles fit before ptxas starts to spill them through vector
// registers (measured with a 24..44-constant test kernel): a pair whose A, B, E
// nonzeros fit -- Ts5, CK5, Me4 -- reads them straight from the constant bank
// (hoisted into uniform registers once, no load in the loop), and the loop's
// other constants make room by living in shared memory (LoopConsts).  Larger
// pairs stream their rows from shared memory into registers, two per LDS.128.
template <class Tab>
struct CoefMode {
    using L = CoefLayout<Tab>;
    static constexpr int NNZ_ALL = L::nnz_range(1, L::ROW_C);
    static constexpr bool UNIFORM = NNZ_ALL <= 30;
};

template <class Tab, int ROW, bool UNIFORM = CoefMode<Tab>::UNIFORM>
struct CoefRow;

template <class Tab, int ROW>
struct CoefRow<Tab, ROW, true> {
    using L = CoefLayout<Tab>;
    __device__ __forceinline__ explicit CoefRow(SAddr) {}
    template <int J>
    __device__ __forceinline__ double at() const {
        if constexpr (ROW < L::S) return Tab::av(ROW, J);
        else if constexpr (ROW == L::ROW_B) return Tab::bv(J);
        else return Tab::ev(J);
    }
};

template <class Tab, int ROW>
struct CoefRow<Tab, ROW, false> {
    using L = CoefLayout<Tab>;
    static constexpr int NNZ = L::nnz(ROW);
    static constexpr int BASE = L::base(ROW);
    double w[NNZ > 0 ? ((NNZ + 1) & ~1) : 2];
    template <int Q>
    __device__ __forceinline__ void load_from(SAddr base) {
        if constexpr (2 * Q + 1 < NNZ) {
            coef_ld2<BASE + 2 * Q>(base, w[2 * Q], w[2 * Q + 1]);
            load_from<Q + 1>(base);
        } else if constexpr (2 * Q < NNZ) {
            coef_ld1<BASE + 2 * Q>(base, w[2 * Q]);
        }
    }
    __device__ __forceinline__ explicit CoefRow(SAddr base) { load_from<0>(base); }
    // coefficient of column J (a structural nonzero)
    template <int J>
    __device__ __forceinline__ double at() const { return w[L::index(ROW, J)]; }
};

template <class Tab, int ROW>
__device__ __forceinline__ void coef_fill_row(double* coef) {
    using L = CoefLayout<Tab>;
    if constexpr (ROW < L::NROWS) {
        static_for<0, L::row_len(ROW)>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (L::value(ROW, j) != 0.0) {
                double v;
                if constexpr (ROW < L::S) v = Tab::av(ROW, j);
                else if constexpr (ROW == L::ROW_B) v = Tab::bv(j);
                else if constexpr (ROW == L::ROW_E) v = Tab::ev(j);
                else v = Tab::cv(j);
                coef[L::base(ROW) + L::index(ROW, j)] = v;
            }
        });
        coef_fill_row<Tab, ROW + 1>(coef);
    }
}

// high word of a non-negative double whose low word is zero
__host__ __device__ constexpr unsigned hi_word_of_small_int(int n) {
    // n = m * 2^e with 1 <= m < 2:  (1023 + e) << 20 | top 20 mantissa bits
    int e = 0;
    while ((n >> (e + 1)) != 0) ++e;
    const unsigned frac = ((unsigned)n << (20 - e)) & 0xfffffu;   // n < 2^20
    return ((unsigned)(1023 + e) << 20) | frac;
}

// FL_SLOW: this step began on the exact path of _reassess_stepsize
enum : unsigned { FL_STD = 1u, FL_REJ = 2u, FL_MF4 = 4u, FL_SLOW = 8u };

template <class Tab, class R>
struct FastLane {
    static constexpr int S = Tab::S;
    static constexpr int NL = R::NL;
    using L = CoefLayout<Tab>;
    using SS = typename Lane<Tab, R>::StiffState;
    static_assert(Tab::VARIANT == tab::GENERIC && !R::WARP, "fast kernel: generic pairs, lane per system");

    double t, h_abs, l2_old;
    double y[NL], f[NL], prm[R::NPL];
    int sys, n_acc, n_rej, nfev0;
    unsigned fl;
    // _diagnose_stiffness bookkeeping (common.py:370-400), in registers: the
    // running mean of h, the okstp at which the 40-step window closes / at
    // which `toomch` fires next, and n_rej at the last window close
    // (jflstp = n_rej - rej_base)
    double havg;
    int next_cnt, next_many, rej_base;
#ifdef XSQ_EVENTS_N
    // scipy's `events=` with no terminal event (kernels compiled at run time with
    // the user's event functions): occurrences so far.  The previous values of
    // the event functions are not kept -- g(t, y) of the state the step starts
    // from is the same number -- and every root is located by event_queue_body.
    int ev_n[XSQ_EVENTS_N];
#endif

    // RungeKutta.__init__ (common.py:187-220) from what ens_init left
    __device__ __forceinline__ void init(const RkDev& P, int idx, SAddr h0) {
        sys = idx;
        t = P.t0;
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            y[k] = P.y0[(long long)k * P.n_lanes + idx];
            f[k] = P.init_f0[(long long)k * P.n_lanes + idx];
        }
        R::load_params(P.params, idx, P.n_lanes, 0, prm);
        n_acc = n_rej = 0;
        nfev0 = P.init_nfev[idx];
        fl = FL_STD;
        l2_old = 0.0;
        h_abs = P.first_step > 0.0 ? P.first_step : P.init_h[idx];
        SS& ss = Lane<Tab, R>::stiff_state();
        ss.bits[threadIdx.x] &= Lane<Tab, R>::SB_PEND1 | Lane<Tab, R>::SB_PEND2;
        havg = 0.0;
        next_many = P.stiff_many_steps > 1 ? P.stiff_many_steps - 1 : 1;
        next_cnt = 20;
        rej_base = 0;
#ifdef XSQ_EVENTS_N
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) ev_n[k] = 0;
#endif
        sts1(h0, h_abs);
    }

    // _reassess_stepsize, common.py:310-331, exact; d = |t_bound - t|.  Returns
    // false when the step is too small (common.py:234).
    __device__ __forceinline__ static bool reassess_exact(const RkDev& P, double t, double d,
                                                       double& h_abs, unsigned& fl) {
        const double min_step = pymax(Tab::H_MIN_A * (fabs(t) + h_abs), XSQ_SQRT_TINY);
        if (h_abs < min_step || h_abs > P.max_step) {
            h_abs = pymin(P.max_step, pymax(min_step, h_abs));
            fl |= FL_STD;
        }
        if (d < 2.0 * h_abs) {
            if (d > h_abs) {
                h_abs = pymax(0.5 * d, min_step);
                fl |= FL_STD;
            } else {
                h_abs = d;
            }
        }
        return !(h_abs < min_step);
    }

    // Stage I with its row of A already in registers.
    template <int I>
    __device__ __forceinline__ void stage(const CoefRow<Tab, I>& a, double (&K)[S + 1][NL], double h) {
        double ys[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            double acc = 0.0;
            static_for<0, I>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if constexpr (Tab::a(I, j) != 0.0) {
                    if constexpr (L::index(I, j) == 0) acc = a.template at<j>() * K[j][c];
                    else acc = fma(a.template at<j>(), K[j][c], acc);
                }
            });
            ys[c] = fma(h, acc, y[c]);
        }
        // right-hand sides of the ensemble path are autonomous or take t from
        // the tableau image; t + c_i h is dead code for the built-in ones
        R::f(__dadd_rn(t, __dmul_rn(Tab::cv(I), h)), ys, prm, K[I]);
    }
    // Stages I..S-1, then y_new = y + h K^T B (common.py:341-351).  The NEXT row
    // of coefficients is requested before the current stage is computed, so the
    // shared-memory latency (~30 cycles) hides behind a whole stage instead of
    // standing at the head of each one.
    template <int I>
    __device__ __forceinline__ void stages(SAddr cb, const CoefRow<Tab, I>& a,
                                           double (&K)[S + 1][NL], double h,
                                           double (&y_new)[NL]) {
        if constexpr (I + 1 < S) {
#if XSQ_V_PREFETCH
            const CoefRow<Tab, I + 1> next(cb);
            stage<I>(a, K, h);
#else
            stage<I>(a, K, h);
            const CoefRow<Tab, I + 1> next(cb);
#endif
            stages<I + 1>(cb, next, K, h, y_new);
        } else {
#if XSQ_V_PREFETCH
            const CoefRow<Tab, L::ROW_B> b(cb);
            stage<I>(a, K, h);
#else
            stage<I>(a, K, h);
            const CoefRow<Tab, L::ROW_B> b(cb);
#endif
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double sb = 0.0;
                static_for<0, S>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    if constexpr (Tab::b(i) != 0.0) sb = fma(b.template at<i>(), K[i][c], sb);
                });
                y_new[c] = fma(h, sb, y[c]);
            }
        }
    }

    // One attempt of a step; returns the lane status.  `cb`: shared-memory
    // address of the coefficient stream, `h0`: this thread's word of the
    // "h_abs at the start of the step" array.
#ifdef XSQ_EVENTS_N
    using EvStash = typename Lane<Tab, R>::EvSlow;
#else
    struct EvStash {};
#endif
    // `stash`: where a step that may hold a terminal event is put for
    // rk_fast_body to look at outside the stepping loop (LANE_EVCHECK)
    template <bool STIFF>
    __device__ __forceinline__ int attempt(const RkDev& P, const SmemAddr& sa, EvStash& stash) {
        const SAddr cb = sa.coef;
        constexpr unsigned HI_N = hi_word_of_small_int(R::N);          // (double)N
        constexpr unsigned HI_TINY = HI_N - (1022u << 20);             // N * 2^-1022
        const double h = h_abs * P.direction;
        double K[S + 1][NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) K[0][c] = f[c];
        double y_new[NL], errv[NL];
        {
            const CoefRow<Tab, 1> a1(cb);
            stages<1>(cb, a1, K, h, y_new);
        }
        const double t_new = t + h;
        double ss = 0.0;
        double atol[NL];
        lc_load<NL, 20>(sa.lc, atol);
        {
#if XSQ_V_PREFETCH
            const CoefRow<Tab, L::ROW_E> e(cb);      // requested before the FSAL evaluation
            if constexpr (Tab::FSAL) R::f(t_new, y_new, prm, K[S]);
#else
            if constexpr (Tab::FSAL) R::f(t_new, y_new, prm, K[S]);
            const CoefRow<Tab, L::ROW_E> e(cb);
#endif
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                double se = 0.0;
                static_for<0, S + Tab::FSAL>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    if constexpr (Tab::e(i) != 0.0) se = fma(e.template at<i>(), K[i][c], se);
                });
                errv[c] = h * se;
                // max(|y|, |y_new|): pick the operand, |.| is a free modifier of the fma
                const double big = fabs(y_new[c]) > fabs(y[c]) ? y_new[c] : y[c];
                const double scale = fma(P.rtol, fabs(big), atol[c]);
                const double q = errv[c] * rcp_scale(scale);
                ss = fma(q, q, ss);
            }
        }
        // ss >= 0 or NaN; N and N * 2^-1022 have a zero low word, so these are
        // exactly  ss < N,  !(ss < inf),  ss < N * 2^-1022  of the generic kernel
        const unsigned sh = (unsigned)__double2hiint(ss);
        const bool accept = sh < HI_N;
        const bool bad = sh >= 0x7ff00000u;
        const bool tiny = sh < HI_TINY;
        const bool rej = (fl & FL_REJ) != 0u;
        const bool second = accept && !(fl & FL_STD);
        // f(t+h, y_new) of non-FSAL pairs, accepted lanes only (common.py:289-291)
        if constexpr (!Tab::FSAL) {
            if (accept) R::f(t_new, y_new, prm, K[S]);
        }
#ifdef XSQ_EVENTS_N
        // find_active_events (ivp.py) on the accepted step
        unsigned ev_active = 0u;
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            const double g0 = user_event(k, t, y, prm);
            const double g1 = user_event(k, t_new, y_new, prm);
            const bool up = g0 <= 0.0 && g1 >= 0.0;
            const bool down = g0 >= 0.0 && g1 <= 0.0;
            const int d = P.ev_direction[k];
            if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) ev_active |= 1u << k;
        }
        if (!accept) ev_active = 0u;
#else
        constexpr unsigned ev_active = 0u;
#endif
        // ---- controller (common.py:249-287) -----------------------------------
        const double l2 = log2_core_s(ss, sa.lg, sa.lc);
        double cc[6];
        lc_load<6, 12>(sa.lc, cc);
        const CtlConst C{cc[0], cc[1], cc[2], cc[3], cc[4]};
        const double factor = ctl_factor_arith<false>(C, l2, l2_old, 0.0, accept, second, rej, tiny,
                                                      (fl & FL_MF4) ? kMaxFactor : kMaxFactor0,
                                                      Exp2Shared{sa.e2, sa.lc});
        const double h_abs_new = h_abs * factor;
        // Everything below is straight-line code with selects: ONE basic block,
        // so that the loads, the controller and the bookkeeping overlap, and one
        // rarely taken branch (`slow`) for the exact forms.
        // ---- stiffness bookkeeping (common.py:370-400), accepted lanes ---------
        const int okstp = n_acc + 1;
        bool probe = false, lotsfl = false;
        double havg_new = havg;
        if constexpr (STIFF) {
            double hc[2];
            lc_load<2, 18>(sa.lc, hc);
            havg_new = hc[0] * havg + hc[1] * h;
            const bool close = accept && okstp == next_cnt;     // okstp == 20 or okstp % 40 == 39
            const bool first = okstp == 20;
            lotsfl = close && !first && (n_rej - rej_base >= 10);
            if (close && first) havg_new = h;
            if (close) {
                rej_base = n_rej;                                // jflstp = 0
                next_cnt = first ? 39 : next_cnt + 40;
            }
            const bool toomch = accept && okstp == next_many;
            if (toomch) next_many += P.stiff_many_steps;
            probe = toomch || lotsfl;
            havg = accept ? havg_new : havg;
        }
        // ---- state (common.py:294-303) -----------------------------------------
        const unsigned fl_acc = (tiny ? FL_STD : 0u) | (fl & FL_MF4) |
            ((unsigned)__double2hiint(factor) < 0x40100000u ? FL_MF4 : 0u);   // factor < 4
        fl = accept ? fl_acc : (fl | FL_REJ);
        n_acc += accept ? 1 : 0;
        n_rej += accept ? 0 : 1;
        l2_old = accept ? l2 : l2_old;
        const double t_next = accept ? t_new : t;
        // ---- what the next attempt needs ---------------------------------------
        // OdeSolver.step (base.py:207-208): done?  Then the next step's
        // _reassess_stepsize: nothing to do when  min_step < h_abs < max_step  and
        // 2 h_abs < |t_bound - t|, each proven on high words; after a rejection
        // h_abs is "certainly above min_step" by the same bound (min_step <=
        // max(H_MIN_A (|t| + h0), sqrt(tiny)) and h0 <= span / 2 whenever the step
        // started on the short path)
        const double s = t_next - P.t_bound;
        const unsigned hh = (unsigned)__double2hiint(h_abs_new);
        const unsigned dh = (unsigned)__double2hiint(s) & 0x7fffffffu;
#if XSQ_V_INTDONE
        // direction * s >= 0 (s is finite): s == 0, or s has the sign of direction
        const bool done = accept & (((dh | (unsigned)__double2loint(s)) == 0u) |
                                    ((int)((unsigned)__double2hiint(s) ^ (unsigned)P.fast_dir_mask) >= 0));
#else
        const bool done = accept & (P.direction * s >= 0.0);
#endif
        const bool near_min = hh <= (unsigned)P.fast_hi_min;
        const bool out_of_range = hh - (unsigned)P.fast_hi_min - 1u >= (unsigned)P.fast_hi_span;
        const bool near_end = (int)(dh - hh) <= 0x100000;
        // (bitwise, not short-circuit: no branches)
        const bool slow_acc = probe | (ev_active != 0u) | (!done & (out_of_range | near_end));
        const bool slow_rej = bad | near_min | ((fl & FL_SLOW) != 0u);
        const bool slow = accept ? slow_acc : slow_rej;
        int st = done ? LANE_FINISHED : LANE_RUNNING;
        if (accept & !done) sts1(sa.h0, h_abs_new);
        if (slow) {
            if (accept) {
                if (STIFF && probe) {
                    if (diagnose_record(P, K, errv, y_new, t_new, h, havg_new, lotsfl)) {
                        if (st == LANE_RUNNING) st = LANE_FLUSH;
                    }
                }
                bool ev_check = false;
#ifdef XSQ_EVENTS_N
                if (ev_active != 0u) {                        // after the probe: it records y_new
                    const int er = handle_events(P, ev_active, K, y_new, t_new, h, stash);
                    if (er < 0) st = LANE_EVQ_FULL;
                    if (er > 0) {                // maybe terminal: decided outside the loop
                        ev_check = true;
                        st = LANE_EVCHECK;
                    }
                }
#endif
                if (!ev_check && !done && (out_of_range || near_end)) {
                    h_abs = h_abs_new;
                    if (!reassess_exact(P, t_next, fabs(s), h_abs, fl)) st = LANE_TOO_SMALL;
                    fl |= FL_SLOW;
                } else {
                    h_abs = h_abs_new;
                }
            } else {
                h_abs = h_abs_new;
                if (bad) {                                    // common.py:280-287
                    h_abs = kMinFactor * fabs(h);             // max(0.2, nan) * h_abs
                    st = LANE_OVERFLOW;
                } else {
                    const double min_step =
                        pymax(Tab::H_MIN_A * (fabs(t) + lds1_volatile(sa.h0)), XSQ_SQRT_TINY);
                    if (h_abs < min_step) st = LANE_TOO_SMALL;               // common.py:234
                }
            }
        } else {
            h_abs = h_abs_new;
        }
        t = t_next;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            y[c] = accept ? y_new[c] : y[c];
            f[c] = accept ? K[S][c] : f[c];
        }
        return st;
    }

#ifdef XSQ_EVENTS_N
    // handle_events (ivp.py) for the accepted step.  No active event at a terminal
    // occurrence: the roots are left to the event queue (0; -1 = the queue is
    // exhausted).  Otherwise (1) the step is copied to `stash` and the lane leaves
    // the stepping loop with LANE_EVCHECK: rk_fast_body locates every root of the
    // step with the general kernel's own code (Lane::events_slow) -- an
    // out-of-line call INSIDE this function would have the registers that live
    // across it spilled in the hot loop (measured: 61 -> 81 ms).
    __device__ __forceinline__ int handle_events(const RkDev& P, unsigned active,
                                                 const double (&K)[S + 1][NL],
                                                 const double (&y_new)[NL], double t_new, double h,
                                                 EvStash& stash) {
#ifdef XSQ_EVENTS_NO_TERMINAL
        // kernels compiled for event sets without a terminal event (the host knows):
        // the stash and the check are not in the code at all (61 instead of 70 ms)
        (void)stash;
        return push_events(P, active, K, y_new, t_new, h) ? 0 : -1;
#else
        bool may_end = false;
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k)
            if ((active >> k & 1u) && P.ev_terminal[k] > 0 && ev_n[k] + 1 >= P.ev_terminal[k])
                may_end = true;
        if (!may_end) return push_events(P, active, K, y_new, t_new, h) ? 0 : -1;
        Lane<Tab, R>::evslow_fill(stash, P, K, y, y_new, prm, ev_n, t, t_new, h, (long long)sys,
                                  active, false);
        return 1;
#endif
    }
    // LANE_EVCHECK, outside the stepping loop: the lane already holds the accepted
    // step (t, y, f, h_abs from the controller; _reassess_stepsize postponed).
    // Returns the lane's status.
    __device__ __forceinline__ int finish_event_check(const RkDev& P, const EvStash& stash) {
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) ev_n[k] = stash.ev_n[k];
        if (stash.terminate) {                     // t, y at the event (ivp.py)
            t = stash.t_stop;
#pragma unroll
            for (int c = 0; c < NL; ++c) y[c] = stash.y_stop[c];
            return LANE_EVENT;
        }
        if (P.direction * (t - P.t_bound) >= 0.0) return LANE_FINISHED;
        if (!reassess_exact(P, t, fabs(P.t_bound - t), h_abs, fl)) return LANE_TOO_SMALL;
        fl |= FL_SLOW;
        return LANE_RUNNING;
    }
    // The step (stages, both end states) of every active event goes to the event
    // queue; t and y still hold the start of the step.  False: the queue is
    // exhausted (the host only selects this kernel when it cannot be).
    __device__ __forceinline__ bool push_events(const RkDev& P, unsigned active,
                                                const double (&K)[S + 1][NL],
                                                const double (&y_new)[NL], double t_new, double h) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            if (!(active >> k & 1u)) continue;
            if (ev_n[k] < P.ev_capacity) {
                const long long idx = evq_alloc(P);
                if (idx >= 0)
                    evq_write<S, NL, S + 1>(P, idx, (long long)sys, k, ev_n[k], false, t, t_new, h, y,
                                            y_new, K);
                else
                    ok = false;
            }
            ++ev_n[k];
        }
        return ok;
    }
#endif

    // The probe of _diagnose_stiffness (common.py:401-516) is deferred exactly as
    // in the generic kernel (Lane::diagnose): its inputs go to a record of the
    // probe queue, or -- queue full -- to one of the thread's two slots.  Returns
    // true when both slots are now taken.
    __device__ __forceinline__ bool diagnose_record(const RkDev& P, double (&K)[S + 1][NL],
                                                    const double (&errv)[NL],
                                                    const double (&y_new)[NL], double t_new,
                                                    double h, double havg_new, bool lotsfl) {
        SS& ss = Lane<Tab, R>::stiff_state();
        using SL = StiffSlot<R>;
        double* s = nullptr;
        long long stride = 1;
        if (P.stiff_q_cap > 0) {
            const unsigned long long qi = atomicAdd(P.stiff_q_count, 1ULL);
            if (qi < (unsigned long long)P.stiff_q_cap) s = P.stiff_q + qi * SL::DOUBLES;
        }
        bool urgent = false;
        if (s == nullptr) {                       // queue full: the thread's slots
            unsigned sbits = ss.bits[threadIdx.x];
            stride = P.stiff_threads;
            const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const int islot = (sbits & Lane<Tab, R>::SB_PEND1) ? 1 : 0;
            s = P.stiff_slot + (long long)islot * SL::DOUBLES * stride + gtid;
            sbits += Lane<Tab, R>::SB_PEND1;
            urgent = (sbits & Lane<Tab, R>::SB_PEND2) != 0u;
            ss.bits[threadIdx.x] = sbits;
        }
        s[0] = t_new;
        s[stride] = h;
        s[2 * stride] = havg_new;
        s[3 * stride] = lotsfl ? 1.0 : 0.0;
        s[4 * stride] = __longlong_as_double((long long)sys);
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

    __device__ __forceinline__ void store(const RkDev& P, int st) {
#pragma unroll
        for (int k = 0; k < NL; ++k) P.y_final[(long long)k * P.n_lanes + sys] = y[k];
        P.t_final[sys] = t;
        if (P.h_next) P.h_next[sys] = h_abs;
        P.n_acc[sys] = n_acc;
        P.n_rej[sys] = n_rej;
        P.nfev[sys] = nfev0 + (S - 1 + Tab::FSAL) * (n_acc + n_rej) + (Tab::FSAL ? 0 : n_acc);
        P.status[sys] = st == LANE_EVENT ? 1 : st;    // 1: a termination event occurred
        if (P.n_eval_done) P.n_eval_done[sys] = 0;
        if (P.stiff_flags)
            P.stiff_flags[sys] = (int)((Lane<Tab, R>::stiff_state().bits[threadIdx.x] >>
                                        Lane<Tab, R>::SB_FLAG_SHIFT) & 7u);
#ifdef XSQ_EVENTS_N
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k)
            P.ev_count[(long long)sys * XSQ_EVENTS_N + k] = ev_n[k];
#endif
    }
};

template <class Tab, class R, int BLOCK, bool STIFF>
__device__ __forceinline__ void rk_fast_body(const RkDev& P) {
    using FL = FastLane<Tab, R>;
    using LN = Lane<Tab, R>;
    __shared__ __align__(16) FastShared<Tab> fs;
    __shared__ double h0[BLOCK];
    __shared__ __align__(16) LoopConsts lcs;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    math_tabs_init();
    if (threadIdx.x == 0) {
        coef_fill_row<Tab, 1>(fs.coef);
        for (int i = 0; i < 6; ++i) {
            lcs.lg_pol[i] = c_xsq_lg_pol[i];
            lcs.e2_pol[i] = c_xsq_e2_pol[i];
        }
        lcs.ctl[0] = P.ctl.a1s; lcs.ctl[1] = P.ctl.a0s; lcs.ctl[2] = P.ctl.a1c;
        lcs.ctl[3] = P.ctl.a2c; lcs.ctl[4] = P.ctl.a0c; lcs.ctl[5] = 0.0;
        lcs.havg[0] = c_xsq_havg[0]; lcs.havg[1] = c_xsq_havg[1];
        for (int i = 0; i < 16; ++i) lcs.atol[i] = P.atol[i];
        memory_fence_for(&lcs);
#ifdef XSQ_EVENTS_N
        evq_cta_init();
#endif
    }
    LN::stiff_state().bits[threadIdx.x] = 0u;
    __syncthreads();
    SmemAddr sa;
    sa.coef = saddr_of(fs.coef);
    sa.lg = saddr_of(math_tabs().lg);
    sa.e2 = saddr_of(math_tabs().e2);
    sa.h0 = saddr_of(&h0[threadIdx.x]);
    sa.lc = saddr_of(&lcs);
    keep_in_register(sa.lc);
    keep_in_register(sa.coef);
    keep_in_register(sa.lg);
    keep_in_register(sa.e2);
    keep_in_register(sa.h0);
    FL L;
    typename FL::EvStash stash;
    bool live = false, exhausted = false;
    auto flush = [&](long long cur) {
        if (!__any_sync(full, LN::probes_pending())) return;
        if (LN::probes_pending()) {
            FL parked = L;
            memory_fence_for(&parked);
            const int evals = LN::flush_probes(P, cur, lane);
            memory_fence_for(&parked);
            L = parked;
            L.nfev0 += evals;
        }
        __syncwarp(full);
    };
    for (;;) {
        // ---- refill: finished threads claim the next trajectories ----
        const unsigned need = __ballot_sync(full, !live && !exhausted);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(P.queue, (unsigned long long)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (!live && !exhausted) {
                const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (idx < P.n_lanes) {
                    L.init(P, (int)idx, sa.h0);
                    live = true;
                    int st = LANE_RUNNING;
                    if (P.t0 == P.t_bound) {                 // scipy base.py:197
                        st = LANE_FINISHED;
                    } else {                                 // the first step's _reassess_stepsize
                        if (!FL::reassess_exact(P, L.t, fabs(P.t_bound - L.t), L.h_abs, L.fl))
                            st = LANE_TOO_SMALL;
                        L.fl |= FL_SLOW;
                    }
                    if (st != LANE_RUNNING) {
                        L.store(P, st);
                        live = false;
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        __syncwarp(full);
        if (__all_sync(full, !live)) {
            // nothing to step: done when the queue is exhausted, else refill (lanes
            // that ended at once -- zero-length span -- must not end the warp)
            if (__all_sync(full, exhausted)) break;
            continue;
        }
        // ---- attempts, until some lane of the warp ends its trajectory ----
        int st = LANE_RUNNING;
        do {
            if (live) st = L.template attempt<STIFF>(P, sa, stash);
        } while (!__any_sync(full, st != LANE_RUNNING));
        // both probe slots of some thread taken (queue full): run them now
        if (STIFF && __any_sync(full, LN::probes_urgent())) flush(live ? L.sys : -1);
#if defined(XSQ_EVENTS_N) && !defined(XSQ_EVENTS_NO_TERMINAL)
        // a step that may hold a terminal event: every root of the step now, with the
        // lane parked in local memory around the out-of-line call (as for the probes)
        if (__any_sync(full, st == LANE_EVCHECK)) {
            if (st == LANE_EVCHECK) {
                FL parked = L;
                memory_fence_for(&parked);
                LN::events_slow(stash);
                memory_fence_for(&parked);
                L = parked;
                st = L.finish_event_check(P, stash);
            }
            __syncwarp(full);
        }
#endif
        if (st == LANE_FLUSH) st = LANE_RUNNING;
        if (st != LANE_RUNNING) {
            L.store(P, st);
            live = false;
        }
        __syncwarp(full);
    }
    if (STIFF) flush(-1);
#ifdef XSQ_EVENTS_N
    if (P.evq_cap > 0 && lane == 0) evq_cta_publish(P);
#endif
}

#ifndef __CUDACC_RTC__
// ---- host side -----------------------------------------------------------------
// The ensemble hot path: adaptive, final state only, default step budget,
// controller without the alpha term.  XSQ_NO_FAST=1 forces the generic kernel
// (tests compare the two bit for bit).
template <class Tab, class R>
inline bool fast_eligible(const RkDev& P) {
    if constexpr (Tab::VARIANT != tab::GENERIC || R::WARP) {
        return false;
    } else {
        if (P.n_forced != 0 || P.n_eval != 0 || P.n_events != 0) return false;
        if (P.minalpha != 0.0 || P.max_steps != 0x7fffffff) return false;
        if (P.n_lanes >= (1LL << 31)) return false;
        const char* e = getenv("XSQ_NO_FAST");
        return !(e && e[0] == '1');
    }
}
// high words that bracket "min_step < h_abs < max_step":
// min_step = max(H_MIN_A (|t| + h0), sqrt(tiny)) <= M for every step that starts
// with 2 h0 < |t_bound - t| (common.py:123-148, 310-331)
inline void fast_prepare_h(RkDev& P, double h_min_a) {
    auto hi_word = [](double x) {
        unsigned long long b;
        memcpy(&b, &x, 8);
        return (long long)(unsigned)(b >> 32);
    };
    const double tmax = fmax(fabs(P.t0), fabs(P.t_bound));
    const double span = fabs(P.t_bound - P.t0);
    const double M = fmax(h_min_a * (tmax + 0.5 * span), 0x1.0p-511) * (1.0 + 0x1.0p-30);
    const long long lo = hi_word(M), hi = hi_word(P.max_step);
    P.fast_hi_min = (int)lo;
    P.fast_hi_span = (int)(hi - lo - 1 > 0 ? hi - lo - 1 : 0);
    P.fast_dir_mask = P.direction < 0.0 ? (int)0x80000000u : 0;
}
template <class Tab>
inline void fast_prepare(RkDev& P) { fast_prepare_h(P, Tab::H_MIN_A); }
#endif  // __CUDACC_RTC__

template <class Tab, class R, int BLOCK, int MINB, bool STIFF>
__global__ void __launch_bounds__(BLOCK, MINB) rk_fast(const RkDev P) {
    rk_fast_body<Tab, R, BLOCK, STIFF>(P);
}

}  // namespace xsq
