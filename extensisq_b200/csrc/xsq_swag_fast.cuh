// xsq_swag_fast.cuh -- SWAG (Shampine-Gordon-Watts variable order Adams PECE,
// shampine.py:180-480) for ensembles of SMALL systems, final state only: the
// same arithmetic, operation by operation, as swag_persistent in
// xsq_swag_core.cuh (results are bit identical; tests/test_kernel_host.py and
// tests/test_gpu_swag.py assert it), organised for the machine:
//
//   * The modified divided differences phi (14 x n doubles) live in REGISTERS.
//     In the generic kernel they are indexed by the per-lane order k and so
//     sit in local memory: 1.3 KB per lane, more than the L1 holds for the
//     resident lanes, i.e. dependent L2 round trips (37 000 cycles per warp
//     and attempted step measured on the Arenstorf ensemble).  Here every
//     loop over phi is unrolled to the compile-time bound and predicated with
//     the lane's own k / ns, so every index is a constant.
//   * The coefficient arrays (psi, alpha, beta, sig, v, w, g: 87 doubles per
//     lane) live in SHARED memory as [index][thread]: a lane-varying index is
//     still conflict free (the bank is the thread), one LDS / STS per access.
//   * Lanes of a warp are at different orders; loops over coefficients run to
//     the lane's own bound (short, scalar), loops over phi to KMAX.
//   * One attempt per loop iteration, retire / refill as in the Runge-Kutta
//     kernels.
// Dense output (t_eval) and events stay with the generic kernel, as does the
// warp-per-system right-hand side.
#pragma once
#include "xsq_swag_core.cuh"
#include "xsq_rk_fast.cuh"     // static_for

namespace xsq {

template <int BLOCK>
struct SwagCoefs {
    double psi[SWAG_KMAX][BLOCK], alpha[SWAG_KMAX][BLOCK], beta[SWAG_KMAX][BLOCK];
    double sig[SWAG_KMAX + 1][BLOCK], v[SWAG_KMAX][BLOCK];
    double g[SWAG_KMAX + 1][BLOCK];
    int iv[SWAG_KMAX][BLOCK];
};

template <class R, int BLOCK>
struct SwagFastLane {
    static constexpr int NL = R::NL;
    static constexpr int KMAX = SWAG_KMAX;
    static_assert(!R::WARP, "fast SWAG kernel: lane per system");
    int sys;
    double t, h, hold;
    double y[NL], yp[NL], wt[NL], prm[R::NPL];
    double phi[SWAG_NCOL][NL];
    int k, kold, kprev, ns, ivc, ifail, k_max;
    int n_acc, n_fail, nfev;
    bool phase1, fresh;
    double min_step;

    __device__ __forceinline__ static double iqq(int i) {
        return 1.0 / ((double)(i + 1) * ((double)(i + 1) + 1.0));
    }

    // shampine.py:99-178
    __device__ __forceinline__ void init(const RkDev& P, int idx, SwagCoefs<BLOCK>& C) {
        const int tid = threadIdx.x;
        sys = idx;
        t = P.t0;
        k_max = P.interpolant;          // k_max travels in this field
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            y[c] = P.y0[(long long)c * P.n_lanes + idx];
            yp[c] = P.init_f0[(long long)c * P.n_lanes + idx];
        }
        R::load_params(P.params, idx, P.n_lanes, 0, prm);
        nfev = P.init_nfev[idx];
        const double b = P.t0 + copysign(fmin(fabs(P.t_bound - P.t0), P.max_step), P.direction);
        if (P.first_step > 0.0) h = copysign(P.first_step, P.direction);
        else h = copysign(P.init_h[idx], b - P.t0);      // ens_init, morder = 1
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            const double yb = y[c] - h * yp[c];
            wt[c] = fma(P.rtol, fmax(fabs(y[c]), fabs(yb)), P.atol[c]);
        }
#pragma unroll
        for (int i = 0; i < SWAG_NCOL; ++i)
#pragma unroll
            for (int c = 0; c < NL; ++c) phi[i][c] = i == 0 ? yp[c] : 0.0;
        C.sig[0][tid] = 1.0;
        C.g[0][tid] = 1.0;
        C.g[1][tid] = 0.5;
        hold = 0.0;
        k = 1;
        kold = kprev = 0;
        phase1 = true;
        ivc = ns = ifail = 0;
        n_acc = n_fail = 0;
        fresh = true;
        min_step = 0.0;
    }

    // One attempted step (the body of dsteps' loop); LANE_RUNNING or a final status.
    __device__ __forceinline__ int attempt(const RkDev& P, SwagCoefs<BLOCK>& C) {
        const int tid = threadIdx.x;
        const double fouru = 4.0 * XSQ_SMALL, twou = 2.0 * XSQ_SMALL;
        if (fresh) {                                   // shampine.py:196-240
            fresh = false;
            ifail = 0;
            min_step = fouru * fabs(t);
            const double d = P.t_bound - t;
            if (fabs(d) <= min_step) {                 // extrapolate onto t_bound
                kold = 0;
#pragma unroll
                for (int c = 0; c < NL; ++c) y[c] = fma(d, yp[c], y[c]);
                t = P.t_bound;
                ++n_acc;
                return LANE_FINISHED;
            }
            if (P.direction * (h - d) > 0.0) h = d;
            if (P.max_step != XSQ_INF) h = copysign(fmin(P.max_step, fabs(h)), P.direction);
            if (fabs(h) < min_step) return LANE_TOO_SMALL;
            double q[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) q[c] = y[c] / wt[c];
            if (0.5 < twou * rms<R>(q)) return LANE_TOL_TOO_TIGHT;
        }
        if (n_acc + n_fail >= P.max_steps) return LANE_STEP_BUDGET;
        const int kp1 = k + 1, km1 = k - 1, km2 = k - 2;
        // ---- block 1: coefficients that change with h or k (shampine.py:247-316)
        if (h != hold) ns = 0;
        if (ns <= kold) ns += 1;
        if (k >= ns) {
            // Written for the warp: every loop runs to the compile-time bound with
            // the lane's own range as a predicate, so the lanes of a warp that are
            // at different (ns, k) walk ONE instruction stream, and the scratch
            // row w of the reference lives in registers.  Same operations on the
            // same operands, in the same order per element, as shampine.py:247-316.
            const int nsm1 = ns - 1;
            const double inv_ns = 1.0 / ns;
            // psi_old[i - nsm1] of the reference is psi[i] before this update; the
            // recurrences below read psi[i - 1] (old) just before writing psi[i]
            double psi_prev = C.psi[nsm1][tid];        // old psi[nsm1]
            double psi_im1_new = h * ns;
            C.psi[nsm1][tid] = psi_im1_new;
            C.alpha[nsm1][tid] = inv_ns;
            C.beta[nsm1][tid] = 1.0;
            if (ns != 1 && k > kprev) {                // order was raised at constant h
                int jv;
                if (ivc != 0) {
                    ivc -= 1;
                    jv = kp1 - C.iv[ivc][tid];
                } else {
                    jv = 1;
                    C.v[km1][tid] = iqq(km1);
                }
                for (int j = jv; j < nsm1; ++j) {
                    const int i = km1 - j;
                    C.v[i][tid] = fma(-C.alpha[j][tid], C.v[i + 1][tid], C.v[i][tid]);
                }
            }
            // v (kept between steps) and w (scratch: every element the g recurrence
            // reads below is written here first, so it never needs to be stored)
            double w[KMAX + 1];
            {
                const int limit1 = kp1 - ns;
                const bool first = ns == 1;
                double vr[KMAX + 1];
                static_for<0, KMAX>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    vr[i] = C.v[i][tid];
                });
                vr[KMAX] = 0.0;
                static_for<0, KMAX>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const double q = 1.0 / ((double)(i + 1) * ((double)(i + 1) + 1.0));
                    const double stepped = fma(-inv_ns, vr[i + 1], vr[i]);
                    double nv = vr[i];
                    if (first ? i < k : i < limit1) nv = first ? q : stepped;
                    C.v[i][tid] = nv;
                    w[i] = nv;
                });
                w[KMAX] = 0.0;
                if (first) {
                    ivc = 0;
                } else {
                    C.g[ns][tid] = w[0];
                    if (k < kold) { C.iv[ivc][tid] = limit1 + 2; ivc += 1; }
                }
            }
            kprev = k;
            // psi, alpha, beta, sig and g for ns <= i < k.  Elements of w beyond the
            // lane's k - i are computed too and never read (the recurrence at i
            // reads w[0 .. k-i], all valid after round i-1).
            double bprod = 1.0, sprod = 1.0, a_prev = inv_ns;
            static_for<1, KMAX>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (i >= ns && i < k) {
                    sprod *= (double)i * a_prev;       // the first factor times 1.0 is exact
                    C.sig[i][tid] = sprod;
                    const double psi_old_im1 = psi_prev;   // psi_old[i - ns] = old psi[i - 1]
                    psi_prev = C.psi[i][tid];              // old psi[i], for the next round
                    const double psi_i = h + psi_old_im1;
                    C.psi[i][tid] = psi_i;
                    const double a_i = div_by(h, div_rcp(psi_i));
                    C.alpha[i][tid] = a_i;
                    bprod *= div_by(psi_im1_new, div_rcp(psi_old_im1));
                    C.beta[i][tid] = bprod;
                    psi_im1_new = psi_i;
                    a_prev = a_i;
                    static_for<0, KMAX - i>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        w[j] = fma(-a_i, w[j + 1], w[j]);
                    });
                    C.g[i + 1][tid] = w[0];
                }
            });
            sprod *= (double)k * a_prev;
            C.sig[k][tid] = sprod;
        }
        // ---- block 2: predict, evaluate, estimate errors (shampine.py:326-364) --
        static_for<1, KMAX>([&](auto ic) {             // phi[i] *= beta[i], ns <= i < k
            constexpr int i = decltype(ic)::value;
            if (i >= ns && i < k) {
                const double b = C.beta[i][tid];
#pragma unroll
                for (int c = 0; c < NL; ++c) phi[i][c] *= b;
            }
        });
        // phi[k+1] = phi[k]; phi[k] = 0
        static_for<1, KMAX + 1>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if (i == k) {
#pragma unroll
                for (int c = 0; c < NL; ++c) {
                    phi[i + 1][c] = phi[i][c];
                    phi[i][c] = 0.0;
                }
            }
        });
        double p[NL];
        {
            double acc[NL];
#pragma unroll
            for (int c = 0; c < NL; ++c) acc[c] = 0.0;
            static_for<0, KMAX>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (i < k) {
                    const double gi_ = C.g[i][tid];
#pragma unroll
                    for (int c = 0; c < NL; ++c) acc[c] = fma(phi[i][c], gi_, acc[c]);
                }
            });
#pragma unroll
            for (int c = 0; c < NL; ++c) p[c] = fma(h, acc[c], y[c]);
        }
        // phi[i] += phi[i+1] for i = k-2 .. 0
        static_for<0, KMAX - 1>([&](auto jc) {
            constexpr int i = KMAX - 2 - decltype(jc)::value;
            if (i <= km2) {
#pragma unroll
                for (int c = 0; c < NL; ++c) phi[i][c] += phi[i + 1][c];
            }
        });
        const double x = t + h;
        const double absh = fabs(h);
        double y_keep[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) y_keep[c] = y[c];
        R::f(x, p, prm, yp);
        ++nfev;
        // phi[k-1], phi[k-2] of this lane (register arrays: select by comparison)
        double phi_km1[NL], phi_km2[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) phi_km1[c] = phi_km2[c] = 0.0;
        static_for<0, KMAX>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if (i == km1) {
#pragma unroll
                for (int c = 0; c < NL; ++c) phi_km1[c] = phi[i][c];
            }
            if (i == km2) {
#pragma unroll
                for (int c = 0; c < NL; ++c) phi_km2[c] = phi[i][c];
            }
        });
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            wt[c] = fma(P.rtol, 0.5 * (fabs(p[c]) + fabs(y[c])), P.atol[c]);
            const double t3 = 1.0 / wt[c], t4 = yp[c] - phi[0][c];
            if (k > 2) { const double q = (phi_km2[c] + t4) * t3; s2 = fma(q, q, s2); }
            if (k > 1) { const double q = (phi_km1[c] + t4) * t3; s1 = fma(q, q, s1); }
            const double q = t4 * t3;
            s0 = fma(q, q, s0);
        }
        double erk, erkm1 = 0.0, erkm2 = 0.0;
        if (k > 2) { erkm2 = absh * sqrt(s2 / (double)R::N); erkm2 *= C.sig[km2][tid] * c_swag_gstr[km2 - 1]; }
        if (k > 1) { erkm1 = absh * sqrt(s1 / (double)R::N); erkm1 *= C.sig[km1][tid] * c_swag_gstr[km2]; }
        erk = absh * sqrt(s0 / (double)R::N);
        const double g_k = C.g[k][tid];
        const double err = erk * (C.g[km1][tid] - g_k);
        erk *= C.sig[k][tid] * c_swag_gstr[km1];
        int knew = k;
        if (k > 2 && fmax(erkm1, erkm2) < erk) knew = km1;
        else if (k == 2 && erkm1 < 0.5 * erk) knew = km1;

        if (!(err <= 1.0)) {
            // ---- block 3: failed step, restore (shampine.py:376-396) ------------
            phase1 = false;
            static_for<0, KMAX>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (i < k) {
                    const DivRcp b = div_rcp(C.beta[i][tid]);
#pragma unroll
                    for (int c = 0; c < NL; ++c) phi[i][c] = div_by(phi[i][c] - phi[i + 1][c], b);
                }
            });
            for (int i = 0; i < km1; ++i) C.psi[i][tid] = C.psi[i + 1][tid] - h;
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
        // ---- block 4: correct, evaluate, update differences (shampine.py:407-468)
        kold = k;
        hold = h;
        const double hg = h * g_k;
#pragma unroll
        for (int c = 0; c < NL; ++c) y[c] = fma(hg, yp[c] - phi[0][c], p[c]);
        R::f(x, y, prm, yp);
        ++nfev;
        // phi[k] = yp - phi[0];  phi[k+1] = phi[k] - phi[k+1];  phi[i] += phi[k], i < k
        double phik[NL], phikp1[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) phik[c] = yp[c] - phi[0][c];
        static_for<1, KMAX + 1>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if (i == k) {
#pragma unroll
                for (int c = 0; c < NL; ++c) {
                    phi[i][c] = phik[c];
                    phi[i + 1][c] = phik[c] - phi[i + 1][c];
                    phikp1[c] = phi[i + 1][c];
                }
            }
        });
        static_for<0, KMAX>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if (i < k) {
#pragma unroll
                for (int c = 0; c < NL; ++c) phi[i][c] += phik[c];
            }
        });
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
            for (int c = 0; c < NL; ++c) q[c] = phikp1[c] / wt[c];
            erkp1 = c_swag_gstr[k] * absh * rms<R>(q);
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
        (void)y_keep;
        return (P.direction * (t - P.t_bound) >= 0.0) ? LANE_FINISHED : LANE_RUNNING;
    }

    __device__ __forceinline__ void store(const RkDev& P, int st) {
#pragma unroll
        for (int c = 0; c < NL; ++c) P.y_final[(long long)c * P.n_lanes + sys] = y[c];
        P.t_final[sys] = t;
        if (P.h_next) P.h_next[sys] = h;
        P.n_acc[sys] = n_acc;
        P.n_rej[sys] = n_fail;
        P.nfev[sys] = nfev;
        P.status[sys] = st;
        if (P.n_eval_done) P.n_eval_done[sys] = 0;
    }
};

template <class R, int BLOCK>
__device__ __forceinline__ void swag_fast_body(const RkDev& P) {
    // 47.6 KB for 64 threads: dynamic shared memory (the launcher opts in above
    // the 48 KB static limit); this kernel stages no dense output, so the
    // dynamic segment is all its own
#ifndef XSQ_HOST_EMU
    SwagCoefs<BLOCK>& coefs = *reinterpret_cast<SwagCoefs<BLOCK>*>(xsq_eval_stage);
#else
    static SwagCoefs<BLOCK> coefs;
#endif
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    SwagFastLane<R, BLOCK> L;
    bool live = false, exhausted = false;
    math_tabs_init();
    for (;;) {
        const unsigned need = __ballot_sync(full, !live && !exhausted);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(P.queue, (unsigned long long)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (!live && !exhausted) {
                const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (idx < P.n_lanes) {
                    L.init(P, (int)idx, coefs);
                    live = true;
                    if (P.t0 == P.t_bound) {
                        L.store(P, LANE_FINISHED);
                        live = false;
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        __syncwarp(full);
        if (__all_sync(full, !live)) {
            if (__all_sync(full, exhausted)) break;
            continue;
        }
        int st = LANE_RUNNING;
        do {
            if (live) st = L.attempt(P, coefs);
        } while (!__any_sync(full, st != LANE_RUNNING));
        if (st != LANE_RUNNING) {
            L.store(P, st);
            live = false;
        }
        __syncwarp(full);
    }
}

// eligible: lane-per-system right-hand side, final state only, no events
template <class R>
inline bool swag_fast_eligible(const RkDev& P) {
    if constexpr (R::WARP || R::NL > 4) {
        return false;
    } else {
        if (P.n_eval != 0 || P.n_events != 0 || P.n_lanes >= (1LL << 31)) return false;
        const char* e = getenv("XSQ_NO_FAST");
        return !(e && e[0] == '1');
    }
}

template <class R, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) swag_fast(const RkDev P) {
    swag_fast_body<R, BLOCK>(P);
}
#ifdef XSQ_SWAG_GEOMETRY_SWEEP
template <class R, int BLOCK, int MAXREG>
__global__ void __maxnreg__(MAXREG) swag_fast_maxreg(const RkDev P) {
    swag_fast_body<R, BLOCK>(P);
}
#endif

}  // namespace xsq
