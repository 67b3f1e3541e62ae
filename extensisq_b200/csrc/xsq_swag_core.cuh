// xsq_swag_core.cuh -- SWAG: Shampine-Gordon-Watts variable-order (1..12)
// Adams-Bashforth-Moulton PECE, one lane or one warp per ODE system, in the
// same persistent / work-queue kernel structure as the Runge-Kutta path.
//
// Computes what these reference functions compute (/root/reference/extensisq):
//   SWAG.__init__        shampine.py:99-178   -> SwagLane::init()
//   SWAG._step_impl      shampine.py:180-480  -> SwagLane::attempt()
//        block 1 (:247-316) coefficient recurrences psi/alpha/beta/sig, v/w -> g
//        block 2 (:326-364) predict, evaluate, error estimates at orders k-2..k
//        block 3 (:376-396) failed step: restore phi/psi, halve or optimal h
//        block 4 (:407-468) correct, evaluate, update differences, order/step
//   SwagDenseOutput      shampine.py:498-587  -> SwagLane::interp()
//   LinearDenseOutput    shampine.py:590-612
//   h_start (morder = 1) common.py:519-763    -> h_start_dev()
//
// Per-system state: the modified divided differences phi (n x 14) and ~100
// doubles of coefficient arrays.  They are indexed by the (per-lane, varying)
// order k, so they live in per-thread local memory (interleaved across the
// warp by the hardware, L1-resident); y, yp, wt and the predictor stay in
// registers.  One loop iteration of the persistent kernel is ONE attempted
// step, so a lane that fails a step does not stall its warp.
#pragma once
#include "xsq_rk_core.cuh"

namespace xsq {

static constexpr int SWAG_KMAX = 12;
static constexpr int SWAG_NCOL = SWAG_KMAX + 2;

static __constant__ __align__(16) double c_swag_two[13] = {
    2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0, 1024.0, 2048.0,
    4096.0, 8192.0};                                   // shampine.py:125-126
static __constant__ __align__(16) double c_swag_gstr[13] = {
    0.5, 0.0833, 0.0417, 0.0264, 0.0188, 0.0143, 0.0114, 0.00936, 0.00789,
    0.00679, 0.00592, 0.00524, 0.00468};               // shampine.py:127-128

enum : int { LANE_TOL_TOO_TIGHT = -3 };

template <class R>
struct SwagLane {
    static constexpr int NL = R::NL;
    long long sys;
    double t, h, hold, t_old, min_step;
    double y[NL], y_old[NL], yp[NL], wt[NL], prm[R::NPL];
    double phi[SWAG_NCOL][NL];
    double psi[SWAG_KMAX], alpha[SWAG_KMAX], beta[SWAG_KMAX],
        sig[SWAG_KMAX + 1], v[SWAG_KMAX], w[SWAG_KMAX + 1], g[SWAG_KMAX + 1],
        gi[SWAG_KMAX];
    int iv[SWAG_KMAX];
    int k, kold, kprev, ns, ivc, kgi, ifail, k_max;
    int n_acc, n_fail, nfev, ieval;
    bool phase1, fresh;
#ifdef XSQ_EVENTS_N
    double ev_g[XSQ_EVENTS_N];     // event function values at (t, y)  (ivp.py `g`)
    int ev_n[XSQ_EVENTS_N];        // occurrences so far               (`event_count`)
#endif

    __device__ __forceinline__ static double iqq(int i) {
        return 1.0 / ((double)(i + 1) * ((double)(i + 1) + 1.0));
    }

    // shampine.py:99-178
    __device__ void init(const RkDev& P, long long idx, int lane) {
        sys = idx;
        t = P.t0;
        k_max = P.interpolant;          // k_max travels in this field
#pragma unroll
        for (int c = 0; c < NL; ++c)
            y[c] = load_slot<R>(P.y0, P.n_lanes, idx, c, lane);
        R::load_params(P.params, idx, P.n_lanes, lane, prm);
#pragma unroll
        for (int c = 0; c < NL; ++c)
            yp[c] = load_slot<R>(P.init_f0, P.n_lanes, idx, c, lane);
        nfev = P.init_nfev[idx];
        const double b = P.t0 + copysign(
            fmin(fabs(P.t_bound - P.t0), P.max_step), P.direction);
        if (P.first_step > 0.0) h = copysign(P.first_step, P.direction);
        else h = copysign(P.init_h[idx], b - P.t0);      // ens_init, morder = 1
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            const double yb = y[c] - h * yp[c];
            wt[c] = fma(P.rtol, fmax(fabs(y[c]), fabs(yb)),
                        atol_of<R>(P, c, lane));
            phi[0][c] = yp[c];
            phi[1][c] = 0.0;
        }
        sig[0] = 1.0;
        g[0] = 1.0;
        g[1] = 0.5;
        hold = 0.0;
        k = 1;
        kold = kprev = 0;
        phase1 = true;
        ivc = kgi = ns = ifail = 0;
        n_acc = n_fail = ieval = 0;
        fresh = true;
        t_old = t;
        min_step = 0.0;
#ifdef XSQ_EVENTS_N
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            ev_g[k] = user_event(k, t, y, prm);
            ev_n[k] = 0;
        }
#endif
    }

    __device__ __forceinline__ double rmsq(const double (&q)[NL]) {
        return rms<R>(q);
    }

    // SwagDenseOutput / LinearDenseOutput at one output point
    __device__ void interp(double xout, double (&yout)[NL]) {
        const double x = t, ox = t_old;
        if (kold == 0) {
            const double xi = (xout - ox) / (x - ox);
#pragma unroll
            for (int c = 0; c < NL; ++c)
                yout[c] = xi * (y[c] - y_old[c]) + y_old[c];
            return;
        }
        double gdi;
        if (kold <= kgi) {
            gdi = gi[kold - 1];
        } else {
            int m;
            if (ivc == 0) { gdi = iqq(kold); m = 1; }
            else { const int iw = iv[ivc - 1]; gdi = w[iw - 1]; m = kold - iw + 2; }
            for (int i = m; i < kold; ++i) gdi = fma(gdi, -alpha[i], w[kold - i]);
        }
        double gl[SWAG_KMAX + 2], wl[SWAG_KMAX + 2];
        const double hi = xout - ox, hh = x - ox, xi = hi / hh, xim1 = xi - 1.0;
        double pw = 1.0;
        for (int i = 0; i <= kold; ++i) { pw *= xi; wl[i] = xi * (pw * iqq(i)); }
        gl[0] = xi;
        gl[1] = 0.5 * xi * xi;
        for (int i = 0; i < kold - 1; ++i) {
            const double alp = alpha[i + 1];
            const int lim = kold - i;
            const double gamma = 1.0 + xim1 * alp;
            for (int j = 0; j < lim; ++j) wl[j] = gamma * wl[j] - alp * wl[j + 1];
            gl[i + 2] = wl[0];
        }
        const double sigma = (wl[1] - xim1 * wl[0]) / gdi;
        for (int i = kold; i >= 1; --i) gl[i] -= gl[i - 1];
        double acc[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) acc[c] = 0.0;
        for (int i = 0; i <= kold; ++i) {
            const double gd = (i == 0) ? g[0] : g[i] - g[i - 1];
            const double cf = gl[i] - sigma * gd;
#pragma unroll
            for (int c = 0; c < NL; ++c) acc[c] = fma(phi[i][c], cf, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < NL; ++c)
            yout[c] = hh * acc[c] + (sigma * y[c] + (1.0 - sigma) * y_old[c]);
    }

    // t_eval points of (t_old, tlim]; tlim < t after a terminal event
    __device__ void emit(const RkDev& P, int lane, double tlim) {
        while (ieval < P.n_eval &&
               P.direction * (P.t_eval[ieval] - tlim) <= 0.0) {
            double yout[NL];
            interp(P.t_eval[ieval], yout);
            eval_put<R>(P, sys, lane, ieval, yout);
            ++ieval;
        }
    }

    // One attempted step.  Returns LANE_RUNNING or a final status.
    __device__ int attempt(const RkDev& P, int lane) {
        const double fouru = 4.0 * XSQ_SMALL, twou = 2.0 * XSQ_SMALL;
        if (fresh) {                                   // shampine.py:196-240
            fresh = false;
            ifail = 0;
            t_old = t;
#pragma unroll
            for (int c = 0; c < NL; ++c) y_old[c] = y[c];
            min_step = fouru * fabs(t);
            const double d = P.t_bound - t;
            if (fabs(d) <= min_step) {                 // extrapolate onto t_bound
                kold = 0;
#pragma unroll
                for (int c = 0; c < NL; ++c) y[c] = fma(d, yp[c], y[c]);
                t = P.t_bound;
                ++n_acc;
                return finish_step(P, lane) ? LANE_EVENT : LANE_FINISHED;
            }
            if (P.direction * (h - d) > 0.0) h = d;
            if (P.max_step != XSQ_INF)
                h = copysign(fmin(P.max_step, fabs(h)), P.direction);
            if (fabs(h) < min_step) return LANE_TOO_SMALL;
            double q[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) q[c] = y[c] / wt[c];
            if (0.5 < twou * rmsq(q)) return LANE_TOL_TOO_TIGHT;
        }
        if (n_acc + n_fail >= P.max_steps) return LANE_STEP_BUDGET;
        const int kp1 = k + 1, km1 = k - 1, km2 = k - 2;
        // ---- block 1: coefficients that change with h or k -----------------
        if (h != hold) ns = 0;
        if (ns <= kold) ns += 1;
        if (k >= ns) {
            const int nsm1 = ns - 1;
            double psi_old[SWAG_KMAX];
            for (int i = nsm1; i < km1; ++i) psi_old[i - nsm1] = psi[i];
            psi[nsm1] = h * ns;
            alpha[nsm1] = 1.0 / ns;
            beta[nsm1] = 1.0;
            double bprod = 1.0;
            for (int i = ns; i < k; ++i) {
                psi[i] = h + psi_old[i - ns];
                alpha[i] = h / psi[i];
                const double ratio = psi[i - 1] / psi_old[i - ns];
                bprod = (i == ns) ? ratio : bprod * ratio;
                beta[i] = bprod;
            }
            double sprod = 1.0;
            for (int i = ns; i <= k; ++i) {
                const double term = (double)i * alpha[i - 1];
                sprod = (i == ns) ? term : sprod * term;
                sig[i] = sprod;
            }
            if (ns == 1) {
                for (int i = 0; i < k; ++i) w[i] = v[i] = iqq(i);
                ivc = kgi = 0;
                if (k != 1) { kgi = 1; gi[0] = w[1]; }
            } else {
                if (k > kprev) {
                    int jv;
                    if (ivc != 0) {
                        ivc -= 1;
                        jv = kp1 - iv[ivc];
                    } else {
                        jv = 1;
                        w[km1] = v[km1] = iqq(km1);
                        if (k == 2) { kgi = 1; gi[0] = w[1]; }
                    }
                    for (int j = jv; j < nsm1; ++j) {
                        const int i = km1 - j;
                        v[i] = fma(-alpha[j], v[i + 1], v[i]);
                        w[i] = v[i];
                    }
                    if (k == ns && jv < nsm1) { kgi = nsm1; gi[kgi - 1] = w[1]; }
                }
                const int limit1 = kp1 - ns;
                for (int i = 0; i < limit1; ++i)
                    v[i] = fma(-alpha[nsm1], v[i + 1], v[i]);
                for (int i = 0; i <= limit1; ++i) w[i] = v[i];
                g[ns] = w[0];
                if (limit1 != 1) { kgi = ns; gi[nsm1] = w[1]; }
                if (k < kold) { iv[ivc] = limit1 + 2; ivc += 1; }
            }
            kprev = k;
            for (int i = ns; i < k; ++i) {
                const int limit2 = k - i;
                for (int j = 0; j < limit2; ++j)
                    w[j] = fma(-alpha[i], w[j + 1], w[j]);
                g[i + 1] = w[0];
            }
        }
        // ---- block 2: predict, evaluate, estimate errors --------------------
        for (int i = ns; i < k; ++i) {
            const double b = beta[i];
#pragma unroll
            for (int c = 0; c < NL; ++c) phi[i][c] *= b;
        }
        double p[NL];
        {
            double acc[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                phi[kp1][c] = phi[k][c];
                phi[k][c] = 0.0;
                acc[c] = 0.0;
            }
            for (int i = 0; i < k; ++i) {
                const double gi_ = g[i];
#pragma unroll
                for (int c = 0; c < NL; ++c) acc[c] = fma(phi[i][c], gi_, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < NL; ++c) p[c] = fma(h, acc[c], y[c]);
        }
        for (int i = km2; i >= 0; --i) {
#pragma unroll
            for (int c = 0; c < NL; ++c) phi[i][c] += phi[i + 1][c];
        }
        const double x = t + h;
        const double absh = fabs(h);
        R::f(x, p, prm, yp);
        ++nfev;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            wt[c] = fma(P.rtol, 0.5 * (fabs(p[c]) + fabs(y[c])),
                        atol_of<R>(P, c, lane));
            const double t3 = 1.0 / wt[c], t4 = yp[c] - phi[0][c];
            if (k > 2) { const double q = (phi[km2][c] + t4) * t3; s2 = fma(q, q, s2); }
            if (k > 1) { const double q = (phi[km1][c] + t4) * t3; s1 = fma(q, q, s1); }
            const double q = t4 * t3;
            s0 = fma(q, q, s0);
        }
        s0 = sys_sum<R::WARP>(s0);
        s1 = sys_sum<R::WARP>(s1);
        s2 = sys_sum<R::WARP>(s2);
        double erk, erkm1 = 0.0, erkm2 = 0.0;
        if (k > 2) { erkm2 = absh * sqrt(s2 / (double)R::N); erkm2 *= sig[km2] * c_swag_gstr[km2 - 1]; }
        if (k > 1) { erkm1 = absh * sqrt(s1 / (double)R::N); erkm1 *= sig[km1] * c_swag_gstr[km2]; }
        erk = absh * sqrt(s0 / (double)R::N);
        const double err = erk * (g[km1] - g[k]);
        erk *= sig[k] * c_swag_gstr[km1];
        int knew = k;
        if (k > 2 && fmax(erkm1, erkm2) < erk) knew = km1;
        else if (k == 2 && erkm1 < 0.5 * erk) knew = km1;

        if (!(err <= 1.0)) {
            // ---- block 3: failed step, restore ------------------------------
            phase1 = false;
            for (int i = 0; i < k; ++i) {
                const double b = beta[i];
#pragma unroll
                for (int c = 0; c < NL; ++c)
                    phi[i][c] = (phi[i][c] - phi[i + 1][c]) / b;
            }
            for (int i = 0; i < km1; ++i) psi[i] = psi[i + 1] - h;
            ++n_fail;
            ++ifail;
            double temp2 = 0.5;
            if (ifail >= 4 && 0.5 < 0.25 * erk) temp2 = sqrt(0.5 / erk);
            if (ifail >= 3) knew = 1;
            h *= temp2;
            k = knew;
            ns = 0;
            if (!(fabs(h) >= min_step)) return LANE_TOO_SMALL;   // also NaN
            return LANE_RUNNING;
        }
        // ---- block 4: correct, evaluate, update differences -----------------
        kold = k;
        hold = h;
        const double hg = h * g[k];
#pragma unroll
        for (int c = 0; c < NL; ++c) y[c] = fma(hg, yp[c] - phi[0][c], p[c]);
        R::f(x, y, prm, yp);
        ++nfev;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            phi[k][c] = yp[c] - phi[0][c];
            phi[kp1][c] = phi[k][c] - phi[kp1][c];
        }
        for (int i = 0; i < k; ++i) {
#pragma unroll
            for (int c = 0; c < NL; ++c) phi[i][c] += phi[k][c];
        }
        if (knew == km1 || k == k_max) phase1 = false;
        double erkp1 = 0.0;
        if (phase1) {
            k = kp1;
            erk = erkp1;
        } else if (knew == km1) {
            k = km1;
            erk = erkm1;
        } else if (k < ns) {
            double q[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) q[c] = phi[kp1][c] / wt[c];
            erkp1 = c_swag_gstr[k] * absh * rmsq(q);
            if (k == 1) {
                if (erkp1 < 0.5 * erk && k < k_max) { k = kp1; erk = erkp1; }
            } else if (erkm1 <= fmin(erk, erkp1)) {
                k = km1;
                erk = erkm1;
            } else if (!(erkp1 > erk || k == k_max)) {
                k = kp1;
                erk = erkp1;
            }
        }
        double hnew;
        if (phase1 || 0.5 >= erk * c_swag_two[k]) {
            hnew = h + h;
        } else if (0.5 >= erk) {
            hnew = h;
        } else {
            // (0.5 / erk) ** (1 / (k + 1)), shampine.py:465, with the controller's
            // table-driven log2 / exp2 (repeated bit for bit by the C oracle)
            const double r = exp2_fast(log2_fast(0.5 / erk) / (double)(k + 1));
            hnew = absh * fmax(0.5, fmin(0.9, r));
            hnew = copysign(fmax(hnew, min_step), h);
        }
        h = hnew;
        t = x;
        ++n_acc;
        fresh = true;
        if (finish_step(P, lane)) return LANE_EVENT;
        return (P.direction * (t - P.t_bound) >= 0.0) ? LANE_FINISHED
                                                      : LANE_RUNNING;
    }

    // What solve_ivp does after solver.step() returned (ivp.py): events on the
    // step's interpolant (SwagDenseOutput / LinearDenseOutput), then the t_eval
    // points.  Returns true when a terminal event ends the trajectory.
    __device__ bool finish_step(const RkDev& P, int lane) {
        double t_stop = t;
        bool terminate = false;
#ifdef XSQ_EVENTS_N
        double g_new[XSQ_EVENTS_N], root[XSQ_EVENTS_N];
        unsigned active = 0;
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) {
            g_new[k] = user_event(k, t, y, prm);
            const bool up = ev_g[k] <= 0.0 && g_new[k] >= 0.0;      // find_active_events
            const bool down = ev_g[k] >= 0.0 && g_new[k] <= 0.0;
            const int d = P.ev_direction[k];
            if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) active |= 1u << k;
        }
        if (active) {
            for (unsigned todo = active; todo; todo &= todo - 1u) {
                const int k = __ffs(todo) - 1;
                const double r = brentq_dev([&](double tt) {
                    double ytmp[NL];
                    interp(tt, ytmp);
                    return user_event(k, tt, ytmp, prm);
                }, t_old, t);
#pragma unroll
                for (int kk = 0; kk < XSQ_EVENTS_N; ++kk)
                    if (kk == k) root[kk] = r;
            }
            double r_star = 0.0;
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k) {
                if (!(active >> k & 1u)) continue;
                ++ev_n[k];
                if (P.ev_terminal[k] > 0 && ev_n[k] >= P.ev_terminal[k]) {
                    if (!terminate || P.direction * (root[k] - r_star) < 0.0) r_star = root[k];
                    terminate = true;
                }
            }
            if (terminate) t_stop = r_star;
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k) {
                if (!(active >> k & 1u)) continue;
                if (terminate && P.direction * (root[k] - r_star) > 0.0) continue;
                const int slot = ev_n[k] - 1;
                if (slot < P.ev_capacity) {
                    const long long base = (sys * XSQ_EVENTS_N + k) * P.ev_capacity + slot;
                    double ye[NL];
                    interp(root[k], ye);
                    P.t_events[base] = root[k];
#pragma unroll
                    for (int c = 0; c < NL; ++c) P.y_events[base * NL + c] = ye[c];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < XSQ_EVENTS_N; ++k) ev_g[k] = g_new[k];
#endif
        if (P.n_eval > 0) emit(P, lane, t_stop);
#ifdef XSQ_EVENTS_N
        if (terminate) {
            double ys[NL];
            interp(t_stop, ys);                                     // y = sol(t)
#pragma unroll
            for (int c = 0; c < NL; ++c) y[c] = ys[c];
            t = t_stop;
        }
#endif
        return terminate;
    }

    __device__ void store(const RkDev& P, int st, int lane,
                          bool constant = false) {
#pragma unroll
        for (int c = 0; c < NL; ++c)
            store_slot<R>(P.y_final, P.n_lanes, sys, c, lane, y[c]);
        if (P.n_eval > 0 && ieval < P.n_eval) {
            eval_finish<R>(P, sys, lane, ieval, constant, y);
            if (constant) ieval = P.n_eval;
        }
        if (!R::WARP || lane == 0) {
            P.t_final[sys] = t;
            if (P.h_next) P.h_next[sys] = h;
            P.n_acc[sys] = n_acc;
            P.n_rej[sys] = n_fail;
            P.nfev[sys] = nfev;
            P.status[sys] = st == LANE_EVENT ? 1 : st;    // 1: a termination event occurred
#ifdef XSQ_EVENTS_N
#pragma unroll
            for (int k = 0; k < XSQ_EVENTS_N; ++k) P.ev_count[sys * XSQ_EVENTS_N + k] = ev_n[k];
#endif
            if (P.n_eval_done) P.n_eval_done[sys] = ieval;
        }
    }
};

template <class R>
__device__ __forceinline__ void swag_persistent_body(const RkDev& P) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    SwagLane<R> L;
    bool live = false, exhausted = false;
    math_tabs_init();
    for (;;) {
        if (R::WARP) {
            if (!live && !exhausted) {
                unsigned long long idx = 0;
                if (lane == 0) idx = atomicAdd(P.queue, 1ULL);
                idx = __shfl_sync(full, idx, 0);
                if ((long long)idx < P.n_lanes) {
                    L.init(P, (long long)idx, lane);
                    live = true;
                    if (P.t0 == P.t_bound) {
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
                        if (P.t0 == P.t_bound) {
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
        int st = LANE_RUNNING;
        do {
            if (live) st = L.attempt(P, lane);
        } while (!__any_sync(full, st != LANE_RUNNING));
        if (st != LANE_RUNNING) {
            L.store(P, st, lane);
            live = false;
        }
        __syncwarp(full);
    }
}

template <class R, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) swag_persistent(const RkDev P) {
    swag_persistent_body<R>(P);
}

}  // namespace xsq
