/*
 * TEST INFRASTRUCTURE -- plain-C restatement of extensisq's SSV2stab solver
 * (Sommeijer-Shampine-Verwer Runge-Kutta-Chebyshev, the reference's port of
 * rkc.f).  NOT part of the product.
 *
 * Follows /root/reference/extensisq/sommeijer.py:
 *   SSV2stab.__init__       :93-145   -> rkc_init()
 *   _init_step_size         :147-160
 *   _step_impl (RKCLOW)     :162-271  -> rkc_step()
 *   _stages (STEP)          :273-329  -> rkc_stages()
 *   _rho (RKCRHO)           :331-398  -> rkc_rho()
 *   _dense_output_impl      :400-406  (CubicDenseOutput, common.py:793-821)
 * and scipy's solve_ivp loop / t_eval slicing (ivp.py:659-731).
 *
 * The reference's test-suite does not cover SSV2stab at all ("parity unpinned"
 * by tests); the pins are (a) the notebook tables docs/Demo_SSV2stab.ipynb
 * :350-356 (steps / rejected / f-evals / s-max of the 3-D heat problem), which
 * tests/test_rkc_oracle_golden.py reproduces through this file, and (b) golden
 * vectors of the unmodified reference on 2-D reaction-diffusion grids
 * (tools/gen_golden_rkc.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define UROUND 0x1.0000000000001p-53
#define SQRT_TINY 0x1.0p-511

typedef void (*vec_rhs_fn)(double t, const double* y, double* dy, int64_t n, void* ctx);

enum { ST_RUNNING = 1, ST_FINISHED = 0, ST_TOO_SMALL = -1, ST_OVERFLOW = -2, ST_SPRAD = -4,
       ST_BUDGET = -5 };

/* built-in PDE: u_t = Lap(u) + u - u^3 on (0,1)^2, Dirichlet 0, nx x nx interior
 * points, 5-point stencil, h = 1/(nx+1)   */
typedef struct { int nx; double inv_h2; } heat2d_ctx;
static void f_heat2d(double t, const double* u, double* du, int64_t n, void* vctx) {
    (void)t; (void)n;
    const heat2d_ctx* c = (const heat2d_ctx*)vctx;
    const int nx = c->nx;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < nx; ++j) {
            const int64_t k = (int64_t)i * nx + j;
            const double uc = u[k];
            const double un = i > 0 ? u[k - nx] : 0.0, us = i < nx - 1 ? u[k + nx] : 0.0;
            const double uw = j > 0 ? u[k - 1] : 0.0, ue = j < nx - 1 ? u[k + 1] : 0.0;
            const double lap = (((un + us) + (uw + ue)) - 4.0 * uc) * c->inv_h2;
            du[k] = lap + (uc - uc * uc * uc);
        }
}

typedef struct {
    vec_rhs_fn f;
    void* ctx;
    int64_t n;
    double rtol, atol, t_bound, direction, max_step, hmin0;
    int const_jac, mmax;
    double rho_const;           /* > 0: rho_jac(t, y) = this constant */
    /* state */
    double t, absh, hold, errold, sprad;
    int have_absh, have_hold, newspc, jacatt, nstsig, have_V;
    double *yn, *fn, *v1, *v2, *y, *V;
    int n_acc, n_rej, nfev, nfesig, maxm;
} rkc_t;

static double rms_scaled(const double* est, const double* wt, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) { const double q = est[i] / wt[i]; s = fma(q, q, s); }
    return sqrt(s / (double)n);
}
static double nrm2(const double* x, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s = fma(x[i], x[i], s);
    return sqrt(s);
}

/* sommeijer.py:331-398; returns < 0 on convergence failure */
static double rkc_rho(rkc_t* S) {
    const int64_t n = S->n;
    const double sqrtu = sqrt(UROUND), small = 1.0 / S->max_step;
    double *v = S->v1, *fv = S->v2;
    if (!S->have_V) { memcpy(S->V, S->fn, sizeof(double) * n); S->have_V = 1; }
    memcpy(v, S->V, sizeof(double) * n);
    const double ynrm = nrm2(S->yn, n), vnrm = nrm2(v, n);
    double dynrm;
    if (ynrm != 0.0 && vnrm != 0.0) {
        dynrm = ynrm * sqrtu;
        const double sc = dynrm / vnrm;
        for (int64_t i = 0; i < n; ++i) v[i] = S->yn[i] + v[i] * sc;
    } else if (ynrm != 0.0) {
        dynrm = ynrm * sqrtu;
        for (int64_t i = 0; i < n; ++i) v[i] *= 1.0 + sqrtu;
    } else if (vnrm != 0.0) {
        dynrm = UROUND;
        const double sc = dynrm / vnrm;
        for (int64_t i = 0; i < n; ++i) v[i] *= sc;
    } else {
        dynrm = UROUND;
        for (int64_t i = 0; i < n; ++i) v[i] = dynrm;
    }
    double sigma = 0.0;
    for (int iter = 0; iter < 50; ++iter) {
        S->f(S->t, v, fv, n, S->ctx);
        S->nfesig++;
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) { const double d = fv[i] - S->fn[i]; s = fma(d, d, s); }
        const double dfnrm = sqrt(s);
        const double sigmal = sigma;
        sigma = dfnrm / dynrm;
        const double sprad = 1.2 * sigma;
        if (iter && fabs(sigma - sigmal) <= fmax(sigma, small) * 0.01) {
            for (int64_t i = 0; i < n; ++i) S->V[i] = v[i] - S->yn[i];
            return sprad;
        }
        if (dfnrm != 0.0) {
            const double sc = dynrm / dfnrm;
            for (int64_t i = 0; i < n; ++i) v[i] = S->yn[i] + (fv[i] - S->fn[i]) * sc;
        } else {
            const int64_t idx = iter % n;
            v[idx] = -v[idx];
        }
    }
    return -1.0;
}

/* sommeijer.py:273-329: y <- RKC step of size h with m stages */
static void rkc_stages(rkc_t* S, double h, int m) {
    const int64_t n = S->n;
    double *yn = S->yn, *fn = S->fn, *y = S->y, *yjm1 = S->v1, *yjm2 = S->v2;
    const double w0 = 1.0 + 2.0 / (13.0 * ((double)m * m));
    double temp1 = w0 * w0 - 1.0;
    double temp2 = sqrt(temp1);
    const double arg = m * log(w0 + temp2);
    const double w1 = sinh(arg) * temp1 / (cosh(arg) * m * temp2 - w0 * sinh(arg));
    double bjm1 = 1.0 / ((2.0 * w0) * (2.0 * w0)), bjm2 = bjm1;
    double mus = w1 * bjm1;
    memcpy(yjm2, yn, sizeof(double) * n);
    { const double hm = h * mus; for (int64_t i = 0; i < n; ++i) yjm1[i] = yn[i] + hm * fn[i]; }
    double thjm2 = 0.0, thjm1 = mus, zjm1 = w0, zjm2 = 1.0, dzjm1 = 1.0, dzjm2 = 0.0,
           d2zjm1 = 0.0, d2zjm2 = 0.0;
    for (int j = 2; j <= m; ++j) {
        const double zj = 2.0 * w0 * zjm1 - zjm2;
        const double dzj = 2.0 * w0 * dzjm1 - dzjm2 + 2.0 * zjm1;
        const double d2zj = 2.0 * w0 * d2zjm1 - d2zjm2 + 4.0 * dzjm1;
        const double bj = d2zj / (dzj * dzj);
        const double ajm1 = 1.0 - zjm1 * bjm1;
        const double mu = 2.0 * w0 * bj / bjm1;
        const double nu = -bj / bjm2;
        mus = mu * w1 / w0;
        S->f(S->t + h * thjm1, yjm1, y, n, S->ctx);
        S->nfev++;
        const double c3 = 1.0 - mu - nu, hm = h * mus;
        for (int64_t i = 0; i < n; ++i)
            y[i] = (mu * yjm1[i] + nu * yjm2[i] + c3 * yn[i]) + hm * (y[i] - ajm1 * fn[i]);
        const double thj = mu * thjm1 + nu * thjm2 + mus * (1.0 - ajm1);
        if (j < m) {
            memcpy(yjm2, yjm1, sizeof(double) * n);
            memcpy(yjm1, y, sizeof(double) * n);
            thjm2 = thjm1; thjm1 = thj; bjm2 = bjm1; bjm1 = bj;
            zjm2 = zjm1; zjm1 = zj; dzjm2 = dzjm1; dzjm1 = dzj;
            d2zjm2 = d2zjm1; d2zjm1 = d2zj;
        }
    }
    if (m < 2) memcpy(y, yjm1, sizeof(double) * n);   /* not reachable: m >= 2 */
}

/* one accepted step with its rejected attempts; sommeijer.py:162-271 */
static int rkc_step(rkc_t* S, int max_steps) {
    const int64_t n = S->n;
    const double one3rd = 1.0 / 3.0, two3rd = 2.0 / 3.0;
    double absh = S->absh, h = 0.0, hmin = 0.0, err = 0.0;
    double* wt = (double*)malloc(sizeof(double) * n);
    double* est = (double*)malloc(sizeof(double) * n);
    int st = ST_RUNNING;
    memcpy(S->y, S->yn, sizeof(double) * n);
    for (;;) {
        if (S->n_acc + S->n_rej >= max_steps) { st = ST_BUDGET; break; }
        if (S->newspc) {
            if (S->rho_const > 0.0) S->sprad = S->rho_const;
            else {
                S->sprad = rkc_rho(S);
                if (S->sprad < 0.0) { st = ST_SPRAD; break; }
            }
            S->jacatt = 1;
        }
        if (!S->have_absh) { /* :147-160 */
            absh = S->max_step;
            if (S->sprad * absh > 1.0) absh = 1.0 / S->sprad;
            absh = fmax(absh, S->hmin0);
            for (int64_t i = 0; i < n; ++i) S->v1[i] = S->yn[i] + absh * S->fn[i];
            S->f(S->t + absh, S->v1, S->v2, n, S->ctx);
            S->nfev++;
            for (int64_t i = 0; i < n; ++i) {
                wt[i] = S->atol + S->rtol * fabs(S->yn[i]);
                est[i] = S->v2[i] - S->fn[i];
            }
            const double e = absh * rms_scaled(est, wt, n);
            if (0.1 * absh < S->max_step * sqrt(e)) absh = fmax(0.1 * absh / sqrt(e), S->hmin0);
            else absh = S->max_step;
            S->have_absh = 1;
        }
        if (1.1 * absh >= fabs(S->t_bound - S->t)) absh = fabs(S->t_bound - S->t);
        int m = 1 + (int)sqrt(1.54 * absh * S->sprad + 1.0);
        if (m > S->mmax) {
            m = S->mmax;
            absh = ((double)m * m - 1) / (1.54 * S->sprad);
        }
        if (m > S->maxm) S->maxm = m;
        h = S->direction * absh;
        hmin = fmax(SQRT_TINY, 13.3 * UROUND * (fabs(S->t) + absh) * ((double)m * m - 1));
        rkc_stages(S, h, m);
        S->f(S->t + h, S->y, S->v1, n, S->ctx);
        S->nfev++;
        for (int64_t i = 0; i < n; ++i) {
            wt[i] = S->atol + S->rtol * fmax(fabs(S->y[i]), fabs(S->yn[i]));
            est[i] = 0.8 * (S->yn[i] - S->y[i]) + 0.4 * h * (S->fn[i] + S->v1[i]);
        }
        err = rms_scaled(est, wt, n);
        if (err < 1.0) break;
        if (isnan(err) || isinf(err)) { st = ST_OVERFLOW; break; }
        S->n_rej++;
        absh = 0.8 * absh / pow(err, one3rd);
        if (absh < hmin) { st = ST_TOO_SMALL; break; }
        S->newspc = !S->jacatt;
        S->absh = absh;
    }
    if (st == ST_RUNNING) {
        S->t += h;
        S->jacatt = S->const_jac;
        S->nstsig = (S->nstsig + 1) % 25;
        S->newspc = 0;
        if (S->rho_const > 0.0 || S->nstsig == 0) S->newspc = !S->jacatt;
        /* rotate W: (yn, fn, v1, v2) <- (y, f(y), yn_old, fn_old) */
        for (int64_t i = 0; i < n; ++i) {
            const double ylast = S->yn[i], yplast = S->fn[i];
            S->yn[i] = S->y[i];
            S->fn[i] = S->v1[i];
            S->v1[i] = ylast;
            S->v2[i] = yplast;
        }
        double fac = 10.0;
        if (!S->have_hold) {
            const double t2 = pow(err, one3rd);
            if (0.8 < fac * t2) fac = 0.8 / t2;
        } else {
            const double t1 = 0.8 * absh * pow(S->errold, one3rd);
            const double t2 = fabs(S->hold) * pow(err, two3rd);
            if (t1 < fac * t2) fac = t1 / t2;
        }
        absh = fmax(0.1, fac) * absh;
        S->absh = fmax(hmin, fmin(S->max_step, absh));
        S->errold = err;
        S->hold = h;
        S->have_hold = 1;
        S->n_acc++;
    }
    free(wt);
    free(est);
    return st;
}

/* Solve one IVP.  rhs_kind 0: user callback; 1: built-in heat2d (nx = sqrt(n)).
 * rho_const > 0 plays the role of rho_jac; <= 0 uses the power iteration.
 * y_eval [n_eval][n] (point-major, unlike the ensemble ABI). */
int xsq_oracle_rkc_solve(int rhs_kind, vec_rhs_fn user_f, int64_t n, const double* y0,
                         double t0, double tf, double rtol, double atol, double first_step,
                         double max_step_in, int const_jac, double rho_const,
                         const double* t_eval, int n_eval, double* y_eval, int max_steps,
                         double* t_final, double* y_final, int32_t* counters /* [6]:
                         n_acc, n_rej, nfev, nfesig, maxm, status */) {
    rkc_t S;
    memset(&S, 0, sizeof S);
    heat2d_ctx hc;
    if (rhs_kind == 1) {
        hc.nx = (int)llround(sqrt((double)n));
        hc.inv_h2 = ((double)hc.nx + 1.0) * ((double)hc.nx + 1.0);
        S.f = f_heat2d; S.ctx = &hc;
    } else { S.f = user_f; S.ctx = 0; }
    if (!S.f) return -1;
    if (max_steps <= 0) max_steps = 2147483647;
    S.n = n; S.rtol = fmin(fmax(rtol, 0x1.4p-50), 0.1); S.atol = fmax(atol, SQRT_TINY);
    S.t = t0; S.t_bound = tf; S.direction = (tf != t0) ? (tf > t0 ? 1.0 : -1.0) : 1.0;
    S.const_jac = const_jac; S.rho_const = rho_const;
    double* buf = (double*)malloc(sizeof(double) * n * 8);
    S.yn = buf; S.fn = buf + n; S.v1 = buf + 2 * n; S.v2 = buf + 3 * n; S.y = buf + 4 * n;
    S.V = buf + 5 * n;
    double* yo = buf + 6 * n; double* fo = buf + 7 * n;
    /* :134-145 */
    int mmax = (int)llround(sqrt(S.rtol / (10.0 * UROUND)));   /* Python round(): half-even */
    { const double r = sqrt(S.rtol / (10.0 * UROUND)); mmax = (int)nearbyint(r); }
    S.mmax = mmax > 2 ? mmax : 2;
    S.newspc = 1; S.jacatt = 0;
    memcpy(S.yn, y0, sizeof(double) * n);
    S.f(t0, S.yn, S.fn, n, S.ctx);
    S.nfev = 1;
    double ms = fmin(max_step_in, fabs(tf - t0));
    S.max_step = fmin(ms, 0x1.fffffffffffffp+511);
    double hm = fabs(t0);
    if (tf != INFINITY) hm = fmax(hm, fabs(S.max_step));
    S.hmin0 = fmax(SQRT_TINY, 10.0 * UROUND * hm);
    if (first_step > 0.0) { S.absh = first_step; S.have_absh = 1; }
    int st = ST_RUNNING, ieval = 0;
    if (t0 == tf) {
        for (int i = 0; i < n_eval; ++i) memcpy(y_eval + (size_t)i * n, S.yn, sizeof(double) * n);
        ieval = n_eval; st = ST_FINISHED;
    }
    while (st == ST_RUNNING) {
        const double t_old = S.t;
        if (n_eval > 0) { memcpy(yo, S.yn, sizeof(double) * n); memcpy(fo, S.fn, sizeof(double) * n); }
        st = rkc_step(&S, max_steps);
        if (st != ST_RUNNING) break;
        while (ieval < n_eval && S.direction * (t_eval[ieval] - S.t) <= 0.0) {
            /* CubicDenseOutput, common.py:793-821 */
            const double hh = S.t - t_old, x = (t_eval[ieval] - t_old) / hh, omx = 1.0 - x;
            const double h00 = (1.0 + 2.0 * x) * (omx * omx), h10 = x * (omx * omx) * hh;
            const double h01 = (x * x) * (3.0 - 2.0 * x), h11 = (x * x) * (x - 1.0) * hh;
            double* out = y_eval + (size_t)ieval * n;
            for (int64_t i = 0; i < n; ++i)
                out[i] = ((h00 * yo[i] + h10 * fo[i]) + h01 * S.yn[i]) + h11 * S.fn[i];
            ++ieval;
        }
        if (S.direction * (S.t - tf) >= 0.0) st = ST_FINISHED;
    }
    for (int i = ieval; i < n_eval; ++i)
        for (int64_t k = 0; k < n; ++k) y_eval[(size_t)i * n + k] = NAN;
    *t_final = S.t;
    memcpy(y_final, S.yn, sizeof(double) * n);
    counters[0] = S.n_acc; counters[1] = S.n_rej; counters[2] = S.nfev;
    counters[3] = S.nfesig; counters[4] = S.maxm; counters[5] = st;
    free(buf);
    return 0;
}
