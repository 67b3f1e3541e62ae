/*
 * TEST INFRASTRUCTURE -- plain-C restatement of extensisq's SWAG solver
 * (Shampine-Gordon-Watts variable-order Adams PECE; the reference's own port of
 * SLATEC ddeabm/dsteps/dintp).  NOT part of the product.
 *
 * Follows /root/reference/extensisq/shampine.py:
 *   SWAG.__init__            :99-178    -> swag_init()
 *   SWAG._step_impl          :180-480   -> swag_step()  (blocks 1-4 of dsteps)
 *   SwagDenseOutput          :498-587   -> swag_interp()
 *   LinearDenseOutput        :590-612
 * and scipy's solve_ivp loop / t_eval slicing (ivp.py:659-731).
 *
 * Pinned by tests/test_oracle_golden.py against golden vectors of the
 * unmodified reference (tools/gen_golden.py -> tests/golden/swag_golden.npz):
 * accepted / failed / nfev counts equal, states to 1e-9 relative.
 * Sums are accumulated in index order with fma() (the reference uses BLAS for
 * phi @ g), the same order as the CUDA kernel.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "xsq_devmath.h"     /* the kernels' own log2 / exp2 (device arithmetic mode) */
int xsq_oracle_device_math(void);

#define SMALL 0x1.0000000000001p-53
#define MAXN 192
#define KMAX 12
#define NCOL (KMAX + 2)

typedef void (*rhs_fn)(double t, const double* y, const double* p, double* dy);
rhs_fn xsq_oracle_builtin_rhs(int id);
double xsq_oracle_h_start(rhs_fn f, const double* prm, int n, double a, double b,
                          const double* y, const double* yprime, int morder,
                          double rtol, const double* atol, int* nfev);

enum { ST_RUNNING = 1, ST_FINISHED = 0, ST_TOO_SMALL = -1, ST_TOL = -3, ST_BUDGET = -5,
       ST_EVENT = 2 /* internal; reported as status 1 */ };

/* scipy's `events=` on SWAG's interpolant (xsq_swag_core.cuh SwagLane::finish_step),
 * device arithmetic; brentq and the built-in event sets are shared with xsq_oracle.c */
typedef double (*event_fn)(int k, double t, const double* y, const double* p);
double xsq_oracle_brentq(double (*fn)(const void*, double), const void* c, double xa, double xb);
event_fn xsq_oracle_builtin_events(int id);
#define MAXEV 8
typedef struct {
    event_fn g;
    int n_events, capacity;
    int terminal[MAXEV], direction[MAXEV];
    double* t_events;   /* [n_events][capacity] */
    double* y_events;   /* [n_events][capacity][n] */
    int32_t* counts;    /* [n_events] */
    double g_old[MAXEV];
} swag_ev_t;

static const double TWO[13] = {2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0,
                               1024.0, 2048.0, 4096.0, 8192.0};
static const double GSTR[13] = {0.5, 0.0833, 0.0417, 0.0264, 0.0188, 0.0143, 0.0114,
                                0.00936, 0.00789, 0.00679, 0.00592, 0.00524, 0.00468};

typedef struct {
    rhs_fn f;
    const double* prm;
    int n, k_max;
    double rtol;
    const double* atol;
    double t_bound, direction, max_step;
    double iqq[KMAX + 1];
    /* state */
    double t, h, hold;
    double y[MAXN], y_old[MAXN], yp[MAXN], wt[MAXN];
    double phi[NCOL][MAXN];
    double psi[KMAX], alpha[KMAX], beta[KMAX], sig[KMAX + 1], v[KMAX], w[KMAX],
        g[KMAX + 1], gi[KMAX];
    int iv[KMAX];
    int k, kold, kprev, ns, ivc, kgi, phase1;
    int n_acc, n_fail, nfev;
} swag_t;

static double rms(const double* x, int n) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s = fma(x[c], x[c], s);
    return sqrt(s / (double)n);
}

/* shampine.py:99-178 */
static void swag_init(swag_t* S, double t0, const double* y0, double first_step) {
    const int n = S->n;
    S->t = t0;
    memcpy(S->y, y0, sizeof(double) * n);
    S->f(t0, S->y, S->prm, S->yp);
    S->nfev = 1;
    if (first_step > 0.0) {
        S->h = copysign(first_step, S->direction);
    } else {
        const double b = t0 + copysign(fmin(fabs(S->t_bound - t0), S->max_step), S->direction);
        S->h = copysign(xsq_oracle_h_start(S->f, S->prm, n, t0, b, S->y, S->yp, 1, S->rtol,
                                           S->atol, &S->nfev),
                        b - t0);
    }
    for (int i = 0; i <= S->k_max; ++i) S->iqq[i] = 1.0 / ((double)(i + 1) * ((double)(i + 1) + 1.0));
    for (int c = 0; c < n; ++c) {
        const double yb = S->y[c] - S->h * S->yp[c];
        S->wt[c] = fma(S->rtol, fmax(fabs(S->y[c]), fabs(yb)), S->atol[c]);
        S->phi[0][c] = S->yp[c];
        S->phi[1][c] = 0.0;
    }
    S->sig[0] = 1.0;
    S->g[0] = 1.0;
    S->g[1] = 0.5;
    S->hold = 0.0;
    S->k = 1;
    S->kold = S->kprev = 0;
    S->phase1 = 1;
    S->ivc = S->kgi = S->ns = 0;
    S->n_acc = S->n_fail = 0;
}

/* One successful step (with its failed attempts).  shampine.py:180-480. */
static int swag_step(swag_t* S, int max_steps) {
    const int n = S->n;
    const double fouru = 4.0 * SMALL, twou = 2.0 * SMALL;
    double x = S->t, h = S->h;
    double p[MAXN], tmp[MAXN];
    int k = S->k, ns = S->ns;
    memcpy(S->y_old, S->y, sizeof(double) * n);
    const double min_step = fouru * fabs(x);
    const double d = S->t_bound - x;
    if (fabs(d) <= min_step) { /* :210-217 extrapolate onto t_bound */
        S->kold = 0;
        for (int c = 0; c < n; ++c) S->y[c] = fma(d, S->yp[c], S->y[c]);
        S->t = S->t_bound;
        S->n_acc++;
        return ST_RUNNING;
    }
    if (S->direction * (h - d) > 0.0) h = d;
    if (S->max_step != INFINITY) h = copysign(fmin(S->max_step, fabs(h)), S->direction);
    if (fabs(h) < min_step) return ST_TOO_SMALL;
    for (int c = 0; c < n; ++c) tmp[c] = S->y[c] / S->wt[c];
    if (0.5 < twou * rms(tmp, n)) return ST_TOL; /* :234-238 */

    int ifail = 0, knew;
    double erk = 0.0, erkm1 = 0.0, erkm2 = 0.0, absh;
    for (;;) {
        if (S->n_acc + S->n_fail >= max_steps) return ST_BUDGET;
        const int kp1 = k + 1, km1 = k - 1, km2 = k - 2;
        /* ---- block 1: coefficients (:247-316) ---- */
        if (h != S->hold) ns = 0;
        if (ns <= S->kold) ns += 1;
        if (k >= ns) {
            const int nsm1 = ns - 1;
            double psi_old[KMAX];
            for (int i = nsm1; i < km1; ++i) psi_old[i - nsm1] = S->psi[i];
            S->psi[nsm1] = h * ns;
            S->alpha[nsm1] = 1.0 / ns;
            S->beta[nsm1] = 1.0;
            double bprod = 1.0;
            for (int i = ns; i < k; ++i) {
                S->psi[i] = h + psi_old[i - ns];
                S->alpha[i] = h / S->psi[i];
                const double ratio = S->psi[i - 1] / psi_old[i - ns];
                bprod = (i == ns) ? ratio : bprod * ratio;
                S->beta[i] = bprod;
            }
            double sprod = 1.0;
            for (int i = ns; i <= k; ++i) {
                const double term = (double)i * S->alpha[i - 1]; /* iq[i-1] = i */
                sprod = (i == ns) ? term : sprod * term;
                S->sig[i] = sprod;
            }
            if (ns == 1) {
                for (int i = 0; i < k; ++i) S->w[i] = S->v[i] = S->iqq[i];
                S->ivc = S->kgi = 0;
                if (k != 1) { S->kgi = 1; S->gi[0] = S->w[1]; }
            } else {
                if (k > S->kprev) {
                    int jv;
                    if (S->ivc != 0) {
                        S->ivc -= 1;
                        jv = kp1 - S->iv[S->ivc];
                    } else {
                        jv = 1;
                        S->w[km1] = S->v[km1] = S->iqq[km1];
                        if (k == 2) { S->kgi = 1; S->gi[0] = S->w[1]; }
                    }
                    for (int j = jv; j < nsm1; ++j) {
                        const int i = km1 - j;
                        S->v[i] = fma(-S->alpha[j], S->v[i + 1], S->v[i]);
                        S->w[i] = S->v[i];
                    }
                    if (k == ns && jv < nsm1) { S->kgi = nsm1; S->gi[S->kgi - 1] = S->w[1]; }
                }
                const int limit1 = kp1 - ns;
                for (int i = 0; i < limit1; ++i) S->v[i] = fma(-S->alpha[nsm1], S->v[i + 1], S->v[i]);
                for (int i = 0; i <= limit1; ++i) S->w[i] = S->v[i];
                S->g[ns] = S->w[0];
                if (limit1 != 1) { S->kgi = ns; S->gi[nsm1] = S->w[1]; }
                if (k < S->kold) { S->iv[S->ivc] = limit1 + 2; S->ivc += 1; }
            }
            S->kprev = k;
            for (int i = ns; i < k; ++i) {
                const int limit2 = k - i;
                for (int j = 0; j < limit2; ++j) S->w[j] = fma(-S->alpha[i], S->w[j + 1], S->w[j]);
                S->g[i + 1] = S->w[0];
            }
        }
        /* ---- block 2: predict, evaluate, estimate errors (:326-364) ---- */
        for (int i = ns; i < k; ++i)
            for (int c = 0; c < n; ++c) S->phi[i][c] *= S->beta[i];
        for (int c = 0; c < n; ++c) {
            S->phi[kp1][c] = S->phi[k][c];
            S->phi[k][c] = 0.0;
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc = fma(S->phi[i][c], S->g[i], acc);
            p[c] = fma(h, acc, S->y[c]);
        }
        for (int i = km2; i >= 0; --i)
            for (int c = 0; c < n; ++c) S->phi[i][c] += S->phi[i + 1][c];
        const double xold = x;
        x += h;
        absh = fabs(h);
        S->f(x, p, S->prm, S->yp);
        S->nfev++;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int c = 0; c < n; ++c) {
            S->wt[c] = fma(S->rtol, 0.5 * (fabs(p[c]) + fabs(S->y[c])), S->atol[c]);
            const double t3 = 1.0 / S->wt[c], t4 = S->yp[c] - S->phi[0][c];
            if (k > 2) { const double q = (S->phi[km2][c] + t4) * t3; s2 = fma(q, q, s2); }
            if (k > 1) { const double q = (S->phi[km1][c] + t4) * t3; s1 = fma(q, q, s1); }
            const double q = t4 * t3;
            s0 = fma(q, q, s0);
        }
        if (k > 2) { erkm2 = absh * sqrt(s2 / (double)n); erkm2 *= S->sig[km2] * GSTR[km2 - 1]; }
        if (k > 1) { erkm1 = absh * sqrt(s1 / (double)n); erkm1 *= S->sig[km1] * GSTR[km2]; }
        erk = absh * sqrt(s0 / (double)n);
        const double err = erk * (S->g[km1] - S->g[k]);
        erk *= S->sig[k] * GSTR[km1];
        knew = k;
        if (k > 2 && fmax(erkm1, erkm2) < erk) knew = km1;
        else if (k == 2 && erkm1 < 0.5 * erk) knew = km1;
        if (err <= 1.0) break;
        /* ---- block 3: failed step, restore (:376-396) ---- */
        S->phase1 = 0;
        x = xold;
        for (int i = 0; i < k; ++i)
            for (int c = 0; c < n; ++c) S->phi[i][c] = (S->phi[i][c] - S->phi[i + 1][c]) / S->beta[i];
        for (int i = 0; i < km1; ++i) S->psi[i] = S->psi[i + 1] - h;
        S->n_fail++;
        ifail++;
        double temp2 = 0.5;
        if (ifail >= 4 && 0.5 < 0.25 * erk) temp2 = sqrt(0.5 / erk);
        if (ifail >= 3) knew = 1;
        h *= temp2;
        k = knew;
        ns = 0;
        if (fabs(h) < min_step) { S->h = h; S->k = k; S->ns = ns; S->t = x; return ST_TOO_SMALL; }
    }
    /* ---- block 4: correct, evaluate, choose order and step (:407-468) ---- */
    const int kp1 = k + 1, km1 = k - 1;
    S->kold = k;
    S->hold = h;
    const double hg = h * S->g[k];
    for (int c = 0; c < n; ++c) S->y[c] = fma(hg, S->yp[c] - S->phi[0][c], p[c]);
    S->f(x, S->y, S->prm, S->yp);
    S->nfev++;
    for (int c = 0; c < n; ++c) {
        S->phi[k][c] = S->yp[c] - S->phi[0][c];
        S->phi[kp1][c] = S->phi[k][c] - S->phi[kp1][c];
    }
    for (int i = 0; i < k; ++i)
        for (int c = 0; c < n; ++c) S->phi[i][c] += S->phi[k][c];
    if (knew == km1 || k == S->k_max) S->phase1 = 0;
    double erkp1 = 0.0;
    if (S->phase1) {
        k = kp1;
        erk = erkp1;
    } else if (knew == km1) {
        k = km1;
        erk = erkm1;
    } else if (k < ns) {
        for (int c = 0; c < n; ++c) tmp[c] = S->phi[kp1][c] / S->wt[c];
        erkp1 = GSTR[k] * absh * rms(tmp, n);
        if (k == 1) {
            if (erkp1 < 0.5 * erk && k < S->k_max) { k = kp1; erk = erkp1; }
        } else if (erkm1 <= fmin(erk, erkp1)) {
            k = km1;
            erk = erkm1;
        } else if (!(erkp1 > erk || k == S->k_max)) {
            k = kp1;
            erk = erkp1;
        }
    }
    double hnew;
    if (S->phase1 || 0.5 >= erk * TWO[k]) hnew = h + h;
    else if (0.5 >= erk) hnew = h;
    else {
        double r;
        if (xsq_oracle_device_math()) {      /* the kernel's own log2 / exp2 */
            const double q = 0.5 / erk;
            double l = dev_log2(q);
            if (q == 0.0) l = -INFINITY;
            if (!(q < INFINITY)) l = q;
            const double z = l / (double)(k + 1);
            r = (fabs(z) < 1000.0) ? dev_exp2(z) : NAN;
        } else {
            r = pow(0.5 / erk, 1.0 / (k + 1));
        }
        hnew = absh * fmax(0.5, fmin(0.9, r));
        hnew = copysign(fmax(hnew, min_step), h);
    }
    S->h = hnew;
    S->k = k;
    S->ns = ns;
    S->t = x;
    S->n_acc++;
    return ST_RUNNING;
}

/* SwagDenseOutput (:498-587) / LinearDenseOutput (:590-612) at one point. */
static void swag_interp(const swag_t* S, double ox, double xout, double* yout) {
    const int n = S->n, kold = S->kold;
    const double x = S->t;
    if (kold == 0) {
        const double xi = (xout - ox) / (x - ox);
        for (int c = 0; c < n; ++c) yout[c] = xi * (S->y[c] - S->y_old[c]) + S->y_old[c];
        return;
    }
    double gdi;
    if (kold <= S->kgi) gdi = S->gi[kold - 1];
    else {
        int m;
        if (S->ivc == 0) { gdi = S->iqq[kold]; m = 1; }
        else { const int iw = S->iv[S->ivc - 1]; gdi = S->w[iw - 1]; m = kold - iw + 2; }
        for (int i = m; i < kold; ++i) gdi = fma(gdi, -S->alpha[i], S->w[kold - i]);
    }
    double gdif[KMAX + 1], g[KMAX + 2], w[KMAX + 2];
    gdif[0] = S->g[0];
    for (int i = 1; i <= kold; ++i) gdif[i] = S->g[i] - S->g[i - 1];
    const double hi = xout - ox, h = x - ox, xi = hi / h, xim1 = xi - 1.0;
    double pw = 1.0;
    for (int i = 0; i <= kold; ++i) { pw *= xi; w[i] = xi * (pw * S->iqq[i]); }
    g[0] = xi;
    g[1] = 0.5 * xi * xi;
    for (int i = 0; i < kold - 1; ++i) {
        const double alp = S->alpha[i + 1];
        const int lim = kold - i;
        const double gamma = 1.0 + xim1 * alp;
        for (int j = 0; j < lim; ++j) w[j] = gamma * w[j] - alp * w[j + 1];
        g[i + 2] = w[0];
    }
    const double sigma = (w[1] - xim1 * w[0]) / gdi;
    for (int i = kold; i >= 1; --i) g[i] -= g[i - 1];
    for (int c = 0; c < n; ++c) {
        double acc = 0.0;
        for (int i = 0; i <= kold; ++i) acc = fma(S->phi[i][c], g[i] - sigma * gdif[i], acc);
        yout[c] = h * acc + (sigma * S->y[c] + (1.0 - sigma) * S->y_old[c]);
    }
}

typedef struct { const void* S; const swag_ev_t* E; double t_old; int k; } swag_root_ctx;
static double swag_ev_of_t(const void* vc, double tt);
static _Thread_local swag_ev_t* g_swag_ev = 0;   /* set per lane by xsq_oracle_swag_events_batch */

static void swag_solve_one(rhs_fn f, int n, const double* y0, const double* prm, double t0,
                           double tf, double rtol, const double* atol, double first_step,
                           double max_step, int k_max, const double* t_eval, int n_eval,
                           double* y_eval, int max_steps, double* t_final, double* y_final,
                           int32_t* n_acc, int32_t* n_fail, int32_t* nfev, int32_t* status,
                           int32_t* n_eval_done, int32_t* k_final) {
    swag_t* S = (swag_t*)malloc(sizeof(swag_t));
    S->f = f; S->prm = prm; S->n = n; S->k_max = k_max;
    S->rtol = rtol; S->atol = atol; S->t_bound = tf; S->max_step = max_step;
    S->direction = (tf != t0) ? (tf > t0 ? 1.0 : -1.0) : 1.0;
    swag_init(S, t0, y0, first_step);
    swag_ev_t* E = g_swag_ev;
    if (E)
        for (int k = 0; k < E->n_events; ++k) E->g_old[k] = E->g(k, t0, S->y, prm);
    int st = ST_RUNNING, ieval = 0;
    double yout[MAXN];
    if (t0 == tf) { /* scipy base.py:195-200 */
        for (int i = 0; i < n_eval; ++i)
            for (int c = 0; c < n; ++c) y_eval[(size_t)c * n_eval + i] = S->y[c];
        ieval = n_eval;
        st = ST_FINISHED;
    }
    while (st == ST_RUNNING) {
        const double t_old = S->t;
        st = swag_step(S, max_steps);
        if (st != ST_RUNNING) break;
        double t_stop = S->t;
        int terminate = 0;
        if (E) { /* find_active_events / handle_events, ivp.py */
            double g_new[MAXEV], root[MAXEV];
            unsigned active = 0;
            for (int k = 0; k < E->n_events; ++k) {
                g_new[k] = E->g(k, S->t, S->y, prm);
                const int up = E->g_old[k] <= 0.0 && g_new[k] >= 0.0;
                const int down = E->g_old[k] >= 0.0 && g_new[k] <= 0.0;
                const int d = E->direction[k];
                if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) active |= 1u << k;
            }
            if (active) {
                for (int k = 0; k < E->n_events; ++k) {
                    if (!(active >> k & 1u)) continue;
                    swag_root_ctx c = {S, E, t_old, k};
                    root[k] = xsq_oracle_brentq(swag_ev_of_t, &c, t_old, S->t);
                }
                double r_star = 0.0;
                for (int k = 0; k < E->n_events; ++k) {
                    if (!(active >> k & 1u)) continue;
                    ++E->counts[k];
                    if (E->terminal[k] > 0 && E->counts[k] >= E->terminal[k]) {
                        if (!terminate || S->direction * (root[k] - r_star) < 0.0) r_star = root[k];
                        terminate = 1;
                    }
                }
                if (terminate) t_stop = r_star;
                for (int k = 0; k < E->n_events; ++k) {
                    if (!(active >> k & 1u)) continue;
                    if (terminate && S->direction * (root[k] - r_star) > 0.0) continue;
                    const int slot = E->counts[k] - 1;
                    if (slot < E->capacity) {
                        const size_t base = (size_t)k * E->capacity + slot;
                        E->t_events[base] = root[k];
                        swag_interp(S, t_old, root[k], E->y_events + base * n);
                    }
                }
            }
            for (int k = 0; k < E->n_events; ++k) E->g_old[k] = g_new[k];
        }
        while (ieval < n_eval && S->direction * (t_eval[ieval] - t_stop) <= 0.0) {
            swag_interp(S, t_old, t_eval[ieval], yout);
            for (int c = 0; c < n; ++c) y_eval[(size_t)c * n_eval + ieval] = yout[c];
            ++ieval;
        }
        if (terminate) { /* t, y = the event point */
            swag_interp(S, t_old, t_stop, yout);
            memcpy(S->y, yout, sizeof(double) * n);
            S->t = t_stop;
            st = ST_EVENT;
        } else if (S->direction * (S->t - tf) >= 0.0) st = ST_FINISHED;
    }
    for (int i = ieval; i < n_eval; ++i)
        for (int c = 0; c < n; ++c) y_eval[(size_t)c * n_eval + i] = NAN;
    *t_final = S->t;
    memcpy(y_final, S->y, sizeof(double) * n);
    *n_acc = S->n_acc; *n_fail = S->n_fail; *nfev = S->nfev; *status = st == ST_EVENT ? 1 : st;
    if (n_eval_done) *n_eval_done = ieval;
    if (k_final) *k_final = S->k;
    free(S);
}

int xsq_oracle_swag_batch(int rhs, rhs_fn user_f, int n, int p, int64_t n_lanes,
                          const double* y0, const double* params, double t0, double tf,
                          double rtol, const double* atol, double first_step, double max_step,
                          int k_max, const double* t_eval, int n_eval, double* y_eval,
                          int max_steps, double* t_final, double* y_final, int32_t* n_acc,
                          int32_t* n_fail, int32_t* nfev, int32_t* status,
                          int32_t* n_eval_done, int32_t* k_final, int n_threads) {
    rhs_fn f = rhs >= 0 ? xsq_oracle_builtin_rhs(rhs) : user_f;
    if (!f || n > MAXN || k_max < 1 || k_max > KMAX) return -1;
    if (max_steps <= 0) max_steps = 2147483647;
    if (rhs < 0) n_threads = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_lanes; ++i)
        swag_solve_one(f, n, y0 + i * n, params ? params + i * p : 0, t0, tf, rtol, atol,
                       first_step, max_step, k_max, t_eval, n_eval,
                       y_eval ? y_eval + (size_t)i * n * n_eval : 0, max_steps, t_final + i,
                       y_final + i * n, n_acc + i, n_fail + i, nfev + i, status + i,
                       n_eval_done ? n_eval_done + i : 0, k_final ? k_final + i : 0);
    return 0;
}

static double swag_ev_of_t(const void* vc, double tt) {
    const swag_root_ctx* c = (const swag_root_ctx*)vc;
    const swag_t* S = (const swag_t*)c->S;
    double ytmp[MAXN];
    swag_interp(S, c->t_old, tt, ytmp);
    return c->E->g(c->k, tt, ytmp, S->prm);
}

/* xsq_oracle_swag_batch with scipy's `events=` (device arithmetic).  ev_set 0: the
 * Lorenz section functions; < 0: `user_g` (single thread).  Output layout as
 * xsq_oracle_rk_events_batch. */
int xsq_oracle_swag_events_batch(int rhs, rhs_fn user_f, int n, int p, int64_t n_lanes,
                                 const double* y0, const double* params, double t0, double tf,
                                 double rtol, const double* atol, double first_step,
                                 double max_step, int k_max, const double* t_eval, int n_eval,
                                 double* y_eval, int max_steps, double* t_final, double* y_final,
                                 int32_t* n_acc, int32_t* n_fail, int32_t* nfev, int32_t* status,
                                 int32_t* n_eval_done, int32_t* k_final, int n_threads,
                                 int ev_set, event_fn user_g, int n_events,
                                 const int32_t* terminal, const int32_t* direction, int capacity,
                                 double* t_events, double* y_events, int32_t* ev_count) {
    rhs_fn f = rhs >= 0 ? xsq_oracle_builtin_rhs(rhs) : user_f;
    event_fn g = ev_set >= 0 ? xsq_oracle_builtin_events(ev_set) : user_g;
    if (!f || !g || n > MAXN || k_max < 1 || k_max > KMAX || n_events < 1 || n_events > MAXEV ||
        capacity < 1 || !xsq_oracle_device_math())
        return -1;
    if (max_steps <= 0) max_steps = 2147483647;
    if (rhs < 0 || ev_set < 0) n_threads = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_lanes; ++i) {
        swag_ev_t E;
        E.g = g; E.n_events = n_events; E.capacity = capacity;
        for (int k = 0; k < n_events; ++k) { E.terminal[k] = terminal[k]; E.direction[k] = direction[k]; }
        E.t_events = t_events + (size_t)i * n_events * capacity;
        E.y_events = y_events + (size_t)i * n_events * capacity * n;
        E.counts = ev_count + (size_t)i * n_events;
        for (int k = 0; k < n_events; ++k) E.counts[k] = 0;
        g_swag_ev = &E;
        swag_solve_one(f, n, y0 + i * n, params ? params + i * p : 0, t0, tf, rtol, atol,
                       first_step, max_step, k_max, t_eval, n_eval,
                       y_eval ? y_eval + (size_t)i * n * n_eval : 0, max_steps, t_final + i,
                       y_final + i * n, n_acc + i, n_fail + i, nfev + i, status + i,
                       n_eval_done ? n_eval_done + i : 0, k_final ? k_final + i : 0);
        g_swag_ev = 0;
    }
    return 0;
}
