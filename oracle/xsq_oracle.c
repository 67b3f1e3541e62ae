/*
 * TEST INFRASTRUCTURE -- plain-C restatement of extensisq's explicit adaptive
 * Runge-Kutta path (scalar loops, one trajectory at a time, OpenMP over
 * trajectories).  NOT part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * It follows, function by function (file:line in /root/reference/extensisq):
 *   validate_tol                 common.py:30-54     (done by the caller, oracle/c_oracle.py)
 *   calculate_scale, norm        common.py:57-66     -> scaled_norm(), rms()
 *   _init_min_step_parameters    common.py:123-148   -> h_min_a()
 *   _init_sc_control             common.py:166-185   -> in rk_solve_one()
 *   RungeKutta.__init__          common.py:187-220
 *   _step_impl                   common.py:222-308
 *   _reassess_stepsize           common.py:310-331
 *   _comp_sol_err, _rk_stage     common.py:333-356
 *   _dense_output_impl, Horner   common.py:358-368, 766-790 (+ cubic 793-821)
 *   h_start                      common.py:519-763
 *   BS5._step_impl, pre-error    bogacki.py:238-346, interpolants 348-393
 *   CFMR7osc._step_impl          calvo.py:152-261
 *   solve_ivp loop / t_eval      scipy/integrate/_ivp/ivp.py:659-731
 *
 * Pinning: tests/test_oracle_golden.py checks this library against the golden
 * vectors produced by the unmodified reference (tools/gen_golden.py): equal
 * accepted/rejected/nfev counts and states to 1e-11 relative.
 *
 * Arithmetic: IEEE double, sums accumulated in index order with fma(), i.e.
 * the summation order differs from the OpenBLAS dgemv the reference uses
 *  -- hence a tolerance, not bit equality, against the
 * reference; but it is the SAME order the CUDA kernels use, so forced-step
 * runs of the CUDA path are compared bit-for-bit against this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "xsq_devmath.h"

/* Device arithmetic (xsq_oracle_set_device_math(1)): err/scale by the kernel's
 * seeded reciprocal, the step-size controller in the log2 domain with the
 * kernel's table-driven log2 / exp2, h_start's tolerance power likewise --
 * every place where extensisq_b200/csrc/xsq_rk_core.cuh deliberately departs
 * from libm / IEEE division.  With it a device run and this file agree BIT FOR
 * BIT in adaptive mode (tests/test_gpu_exact.py); without it this file follows
 * the reference's own expressions (pow, true division). */
const uint8_t* xsq_rcp64h_delta = 0;
static int g_device_math = 0;
void xsq_oracle_set_rcp_table(const uint8_t* bits) { xsq_rcp64h_delta = bits; }
int xsq_oracle_device_math(void) { return g_device_math; }
int xsq_oracle_set_device_math(int on) {
    if (on && !xsq_rcp64h_delta) return -1;
    g_device_math = on;
    return 0;
}

#define MAXS 18
#define MAXN 1024
#define MAXPOL 8
#define KROWS (MAXS + 4)

#define SQRT_TINY 0x1.0p-511
#define BIG 0x1.fffffffffffffp+511
#define SMALL 0x1.0000000000001p-53
#define RELPER 0x1.172b83c7d517bp-20

enum { V_GENERIC = 0, V_BS5 = 1, V_CFMR = 2, V_CKDISC = 3, V_NYSTROM = 4 };
enum { ST_FINISHED = 0, ST_EVENT = 2, ST_TOO_SMALL = -1, ST_OVERFLOW = -2, ST_BUDGET = -5 };
/* ST_EVENT is internal (1 is "running" in the loops below); reported as status 1 */
enum { IP_FREE = 1, IP_LOW = 2, IP_BEST = 3 };
enum { RHS_LORENZ = 0, RHS_VDP = 1, RHS_ARENSTORF = 2, RHS_NBODY32 = 3 };

typedef struct {
    int32_t s, order, order2, npol, variant, pad;
    double A[MAXS][MAXS], B[MAXS], C[MAXS], E[MAXS + 1];
    double P[MAXS + 1][MAXPOL];
    /* BS5 only */
    double E_pre[MAXS], B_scale_pre[MAXS], C_extra[3], A_extra[3][MAXS];
    double Plow[MAXS + 2][MAXPOL], Pbest[MAXS + 4][MAXPOL];
    int32_t npol_low, npol_best;
    double sc[4];
    double stbrad, tanang;      /* <= 0: stiffness diagnosis not implemented */
    /* CKdisc only (cash.py:184-236) */
    double B_assess[2][MAXS], E_assess[2][MAXS], B_fallback[2][MAXS], E_fallback[2][MAXS];
    double C_fallback[2], ck_max_factor, ck_min_factor, ck_safety;
    /* Runge-Kutta-Nystrom methods only (common.py:1207-1309): A', B', E' */
    double Ap[MAXS][MAXS], Bp[MAXS], Ep[MAXS + 1];
} otab_t;

typedef void (*rhs_fn)(double t, const double* y, const double* p, double* dy);

/* ---- built-in right-hand sides (same expression trees as xsq_rhs.cuh) ---- */
static void f_lorenz(double t, const double* y, const double* p, double* dy) {
    (void)t;
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = fma(y[0], p[1] - y[2], -y[1]);
    dy[2] = fma(y[0], y[1], -(p[2] * y[2]));
}
static void f_vdp(double t, const double* y, const double* p, double* dy) {
    (void)t;
    dy[0] = y[1];
    dy[1] = fma(p[0] * fma(-y[0], y[0], 1.0), y[1], -y[0]);
}
static void f_arenstorf(double t, const double* y, const double* p, double* dy) {
    (void)t;
    const double mu = p[0], mup = 1.0 - mu;
    const double xa = y[0] + mu, xb = y[0] - mup;
    double d1 = fma(xa, xa, y[1] * y[1]);
    d1 = d1 * sqrt(d1);
    double d2 = fma(xb, xb, y[1] * y[1]);
    d2 = d2 * sqrt(d2);
    dy[0] = y[2];
    dy[1] = y[3];
    dy[2] = fma(2.0, y[3], y[0]) - mup * xa / d1 - mu * xb / d2;
    dy[3] = fma(-2.0, y[2], y[1]) - mup * y[1] / d1 - mu * y[1] / d2;
}
/* y = [pos(3*32), vel(3*32)], p = (eps2, m[32]) */
static void f_nbody32(double t, const double* y, const double* p, double* dy) {
    (void)t;
    const int nb = 32;
    for (int i = 0; i < nb; ++i) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < nb; ++j) {
            const double dx = y[3 * j] - y[3 * i];
            const double dyy = y[3 * j + 1] - y[3 * i + 1];
            const double dz = y[3 * j + 2] - y[3 * i + 2];
            const double r2 = fma(dz, dz, fma(dyy, dyy, fma(dx, dx, p[0])));
            const double inv = 1.0 / (r2 * sqrt(r2));
            const double w = (j == i) ? 0.0 : p[1 + j] * inv;
            ax = fma(w, dx, ax);
            ay = fma(w, dyy, ay);
            az = fma(w, dz, az);
        }
        dy[3 * i] = y[3 * nb + 3 * i];
        dy[3 * i + 1] = y[3 * nb + 3 * i + 1];
        dy[3 * i + 2] = y[3 * nb + 3 * i + 2];
        dy[3 * nb + 3 * i] = ax;
        dy[3 * nb + 3 * i + 1] = ay;
        dy[3 * nb + 3 * i + 2] = az;
    }
}
static rhs_fn builtin_rhs(int id) {
    switch (id) {
        case RHS_LORENZ: return f_lorenz;
        case RHS_VDP: return f_vdp;
        case RHS_ARENSTORF: return f_arenstorf;
        case RHS_NBODY32: return f_nbody32;
        default: return 0;
    }
}

struct ev_ctx_s;
typedef struct {
    const otab_t* T;
    rhs_fn f;
    const double* prm;
    int n;
    double rtol;
    const double* atol; /* [n] */
    double t_bound, direction, max_step;
    double err_exp, minbeta1, minbeta2, minalpha, safety, safety_sc, h_min_a;
    int interpolant;
    /* state */
    double t, h_abs, h_prev, err_old, max_factor, min_step;
    double l2_old;                 /* device arithmetic: log2 of the last accepted ss */
    double log2n;
    double ctl_a1s, ctl_a0s, ctl_a1c, ctl_a2c, ctl_a0c;   /* xsq_api.cu build_params */
    double y[MAXN], fcur[MAXN];
    double K[KROWS][MAXN];
    int n_acc, n_rej, nfev, standard_sc;
    int force_cubic;               /* CKdisc: a fallback solution was accepted (cash.py:406-416) */
    struct ev_ctx_s* ev;           /* scipy's `events=` (device arithmetic only), or NULL */
    /* stiffness diagnosis, common.py:150-164 */
    int nfev_stiff_detect, jflstp, okstp, stiff_flags, n_stiff_tests;
    double havg;
} lane_t;

/* Order of the sums over the components.  Default: sequential, as NumPy's
 * and the lane-per-system kernels'.  "warp strided" (xsq_oracle_set_warp_strided)
 * repeats a warp-per-system kernel with component c in lane c % 32
 * (xsq_rhs.cuh WideSystem): every lane accumulates its own components in
 * turn, then the 32 partial sums are combined by the xor butterfly of
 * sys_sum (xsq_rk_core.cuh). */
static int g_warp_strided = 0;
int xsq_oracle_set_warp_strided(int on) {
    const int old = g_warp_strided;
    g_warp_strided = on != 0;
    return old;
}
typedef struct { double part[32]; } wacc_t;
static void wacc_init(wacc_t* w) {
    const int m = g_warp_strided ? 32 : 1;
    for (int l = 0; l < m; ++l) w->part[l] = 0.0;
}
static double* wacc_at(wacc_t* w, int c) { return &w->part[g_warp_strided ? (c & 31) : 0]; }
static double wacc_total(wacc_t* w) {
    if (!g_warp_strided) return w->part[0];
    for (int o = 16; o > 0; o >>= 1) {
        double t[32];
        for (int l = 0; l < 32; ++l) t[l] = w->part[l] + w->part[l ^ o];
        for (int l = 0; l < 32; ++l) w->part[l] = t[l];
    }
    return w->part[0];
}

/* common.py:64-66 */
static double rms(const double* x, int n) {
    wacc_t w;
    wacc_init(&w);
    for (int c = 0; c < n; ++c) { double* s = wacc_at(&w, c); *s = fma(x[c], x[c], *s); }
    return sqrt(wacc_total(&w) / (double)n);
}

/* common.py:519-763 */
static double h_start(lane_t* L, double a, double b, int morder) {
    const int n = L->n;
    double spy[MAXN], pv[MAXN], yp[MAXN], sf[MAXN];
    const double* y = L->y;
    const double* yprime = L->fcur;
    const double dx = b - a, absdx = fabs(dx);
    double da = copysign(fmax(fmin(RELPER * fabs(a), absdx), 100.0 * SMALL * fabs(a)), dx);
    if (da == 0.0) da = RELPER * dx;
    L->f(a + da, y, L->prm, sf);
    L->nfev++;
    for (int c = 0; c < n; ++c) yp[c] = sf[c] - yprime[c];
    double delf = rms(yp, n);
    double dfdxb = BIG;
    if (delf < BIG * fabs(da)) dfdxb = delf / fabs(da);
    double fbnd = rms(sf, n);
    double dely = RELPER * rms(y, n);
    if (dely == 0.0) dely = RELPER;
    dely = copysign(dely, dx);
    delf = rms(yprime, n);
    fbnd = fmax(fbnd, delf);
    if (delf != 0.0) {
        for (int c = 0; c < n; ++c) { spy[c] = yprime[c]; yp[c] = yprime[c]; }
    } else {
        for (int c = 0; c < n; ++c) { spy[c] = 0.0; yp[c] = 1.0; }
        delf = rms(yp, n);
    }
    double dfdub = 0.0;
    const int lk = n + 1 < 3 ? n + 1 : 3;
    for (int k = 1; k <= lk; ++k) {
        const double q = dely / delf;
        for (int c = 0; c < n; ++c) pv[c] = fma(q, yp[c], y[c]);
        if (k == 2) {
            L->f(a + da, pv, L->prm, yp);
            for (int c = 0; c < n; ++c) pv[c] = yp[c] - sf[c];
        } else {
            L->f(a, pv, L->prm, yp);
            for (int c = 0; c < n; ++c) pv[c] = yp[c] - yprime[c];
        }
        L->nfev++;
        fbnd = fmax(fbnd, rms(yp, n));
        delf = rms(pv, n);
        if (delf >= BIG * fabs(dely)) { dfdub = BIG; break; }
        dfdub = fmax(dfdub, delf / fabs(dely));
        if (k == lk) break;
        if (delf == 0.0) delf = 1.0;
        for (int c = 0; c < n; ++c) {
            double dy;
            if (k == 2) dy = (y[c] != 0.0) ? y[c] : dely / RELPER;
            else dy = (pv[c] != 0.0) ? pv[c] : delf;
            if (spy[c] == 0.0) spy[c] = yp[c];
            yp[c] = (spy[c] != 0.0) ? copysign(dy, spy[c]) : dy;
        }
        delf = rms(yp, n);
    }
    const double ydpb = fma(dfdub, fbnd, dfdxb);
    double tolmin = INFINITY;
    wacc_t tolacc;
    wacc_init(&tolacc);
    for (int c = 0; c < n; ++c) {
        const double etol = fma(L->rtol, fabs(y[c]), L->atol[c]);
        /* device: 10^(c log10 x) as 2^(c log2 x) with its own log2 / exp2 */
        const double te = g_device_math ? dev_log2(etol) : log10(etol);
        *wacc_at(&tolacc, c) += te;
        tolmin = fmin(tolmin, te);
    }
    const double tolsum = wacc_total(&tolacc);
    tolmin = fmin(tolmin, BIG);
    const double texp = 0.5 * (tolsum / (double)n + tolmin) / (double)(morder + 1);
    const double tolp = g_device_math ? dev_exp2(texp) : pow(10.0, texp);
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
    h = fmax(h, 100.0 * SMALL * fabs(a));
    if (h == 0.0) h = SMALL * fabs(b);
    return fabs(h);
}

/* sum_j w[j] K[j][c] over nonzero weights, first term a plain product */
static double wsum_first(lane_t* L, const double* w, int m, int c) {
    double acc = 0.0;
    int first = 1;
    for (int j = 0; j < m; ++j) {
        if (w[j] != 0.0) {
            acc = first ? w[j] * L->K[j][c] : fma(w[j], L->K[j][c], acc);
            first = 0;
        }
    }
    return acc;
}
/* same but accumulated with fma from 0.0 */
static double wsum(lane_t* L, const double* w, int m, int c) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j)
        if (w[j] != 0.0) acc = fma(w[j], L->K[j][c], acc);
    return acc;
}

/* RungeKuttaNystrom._rk_stage, common.py:1279-1285 (xsq_rk_core.cuh Lane::stage):
 * slots [0, n/2) are positions, [n/2, n) their velocities; rows of K are whole
 * derivative vectors of which only the acceleration half is read */
static void rkn_stage(lane_t* L, double h, int i) {
    const int nh = L->n / 2;
    const otab_t* T = L->T;
    double ys[MAXN];
    const double dt = T->C[i] * h, hh = h * h;
    for (int c = 0; c < nh; ++c) {
        double au = 0.0, av = 0.0;
        int fu = 1, fv = 1;
        for (int j = 0; j < i; ++j) {
            if (T->A[i][j] != 0.0) {
                au = fu ? T->A[i][j] * L->K[j][nh + c] : fma(T->A[i][j], L->K[j][nh + c], au);
                fu = 0;
            }
            if (T->Ap[i][j] != 0.0) {
                av = fv ? T->Ap[i][j] * L->K[j][nh + c] : fma(T->Ap[i][j], L->K[j][nh + c], av);
                fv = 0;
            }
        }
        ys[c] = L->y[c] + (au * hh + dt * L->y[nh + c]);
        ys[nh + c] = L->y[nh + c] + av * h;
    }
    L->f(L->t + dt, ys, L->prm, L->K[i]);
    L->nfev++;
}

/* common.py:353-356 */
static void rk_stage(lane_t* L, double h, int i) {
    if (L->T->variant == V_NYSTROM) { rkn_stage(L, h, i); return; }
    double ys[MAXN];
    for (int c = 0; c < L->n; ++c) ys[c] = fma(h, wsum_first(L, L->T->A[i], i, c), L->y[c]);
    L->f(L->t + L->T->C[i] * h, ys, L->prm, L->K[i]);
    L->nfev++;
}

/* norm(err / scale(y, yref)), common.py:57-66, 338-339 */
static double scaled_norm(lane_t* L, const double* errv, const double* yref) {
    double q[MAXN];
    for (int c = 0; c < L->n; ++c) {
        const double scale = fma(L->rtol, fmax(fabs(L->y[c]), fabs(yref[c])), L->atol[c]);
        q[c] = errv[c] / scale;
    }
    return rms(q, L->n);
}

/* device form: sum((err * rcp(scale))^2); error_norm < 1  <=>  ss < n exactly */
static double scaled_ss_dev(lane_t* L, const double* errv, const double* yref) {
    wacc_t w;
    wacc_init(&w);
    for (int c = 0; c < L->n; ++c) {
        const double big = fabs(yref[c]) > fabs(L->y[c]) ? yref[c] : L->y[c];
        const double scale = fma(L->rtol, fabs(big), L->atol[c]);
        const double q = errv[c] * dev_rcp_scale(scale);
        double* ss = wacc_at(&w, c);
        *ss = fma(q, q, *ss);
    }
    return wacc_total(&w);
}

/* xsq_rk_core.cuh ctl_factor */
static double ctl_factor_dev(const lane_t* L, double l2, double z_extra, int use_extra, int accept,
                             int second, int rej, int tiny) {
    const double z_std = fma(L->ctl_a1s, l2, L->ctl_a0s);
    double z_sc = fma(L->ctl_a1c, l2, fma(L->ctl_a2c, L->l2_old, L->ctl_a0c));
    if (use_extra) z_sc += z_extra;
    const double raw = dev_exp2(second ? z_sc : z_std);
    double factor = raw;
    if ((!accept || second) && !(raw > 0.2)) factor = 0.2;
    const double hi = (accept && rej) ? 1.0 : (second ? L->max_factor : INFINITY);
    if (!(factor < hi)) factor = hi;
    if (accept && tiny) factor = rej ? 1.0 : L->max_factor;
    return factor;
}
static double log2_fast_dev(double x) {
    double r = dev_log2(x);
    if (x == 0.0) r = -INFINITY;
    if (!(x < INFINITY)) r = x;
    return r;
}

static void reassess(lane_t* L) { /* common.py:310-331 */
    L->min_step = fmax(L->h_min_a * (fabs(L->t) + L->h_abs), SQRT_TINY);
    if (L->h_abs < L->min_step || L->h_abs > L->max_step) {
        L->h_abs = fmin(L->max_step, fmax(L->min_step, L->h_abs));
        L->standard_sc = 1;
    }
    const double d = fabs(L->t_bound - L->t);
    if (d < 2.0 * L->h_abs) {
        if (d > L->h_abs) {
            L->h_abs = fmax(0.5 * d, L->min_step);
            L->standard_sc = 1;
        } else {
            L->h_abs = d;
        }
    }
}

/* Dense output over [t, t_new] (K complete); writes points into out[c*n_eval+i] */
static void bs5_extra(lane_t* L, double h, int r) {
    const int row = L->T->s + 1 + r;
    double ys[MAXN];
    for (int c = 0; c < L->n; ++c) ys[c] = fma(wsum(L, L->T->A_extra[r], row, c), h, L->y[c]);
    L->f(L->t + L->T->C_extra[r] * h, ys, L->prm, L->K[row]);
    L->nfev++;
}

/* The dense output of the step just accepted (common.py:766-821), formed once per
 * step and evaluated at the t_eval points and wherever the root finder asks. */
typedef struct {
    double Q[MAXPOL][MAXN];
    double t_anchor, h_anchor;
    int npol, anchor_end, cubic, built;
} dense_t;

static void dense_build(lane_t* L, dense_t* D, double h, double t_new, const double* y_new) {
    const otab_t* T = L->T;
    const int n = L->n, s = T->s;
    D->built = 1;
    D->cubic = T->npol == 0 || L->force_cubic;
    D->npol = T->npol;
    D->t_anchor = L->t;
    D->h_anchor = t_new - L->t;
    D->anchor_end = 0;
    if (D->cubic) return;
    if (T->variant == V_BS5 && L->interpolant == IP_LOW) {
        bs5_extra(L, h, 0);
        D->npol = T->npol_low;
        for (int k = 0; k < D->npol; ++k)
            for (int c = 0; c < n; ++c) {
                double acc = 0.0;
                for (int i = 0; i <= s + 1; ++i)
                    if (T->Plow[i][k] != 0.0) acc = fma(T->Plow[i][k], L->K[i][c], acc);
                D->Q[k][c] = acc;
            }
    } else if (T->variant == V_BS5 && L->interpolant == IP_BEST) {
        bs5_extra(L, h, 0);
        bs5_extra(L, h, 1);
        bs5_extra(L, h, 2);
        D->npol = T->npol_best;
        for (int c = 0; c < n; ++c) { /* bogacki.py:372-388 */
            double kp[11];
            D->Q[0][c] = L->K[7][c];
#define KP(col) for (int i = 0; i < 11; ++i) kp[i] = L->K[i][c] * T->Pbest[i][col];
            KP(1) D->Q[1][c] = (kp[4] + ((kp[5] + kp[7]) + kp[0]) + ((kp[2] + kp[8]) + kp[9]) +
                                ((kp[3] + kp[10]) + kp[6]));
            KP(2) D->Q[2][c] = (kp[4] + kp[5] + ((kp[2] + kp[8]) + (kp[9] + kp[7]) + kp[0]) +
                                ((kp[3] + kp[10]) + kp[6]));
            KP(3) D->Q[3][c] = (((kp[3] + kp[7]) + (kp[6] + kp[5]) + kp[4]) +
                                ((kp[9] + kp[8]) + (kp[2] + kp[10]) + kp[0]));
            KP(4) D->Q[4][c] = ((kp[9] + kp[8]) + ((kp[6] + kp[5]) + kp[4]) +
                                ((kp[3] + kp[7]) + (kp[2] + kp[10]) + kp[0]));
            KP(5) D->Q[5][c] = (kp[4] + ((kp[9] + kp[7]) + (kp[6] + kp[5])) +
                                ((kp[3] + kp[8]) + (kp[2] + kp[10]) + kp[0]));
#undef KP
        }
        D->anchor_end = 1;
        D->t_anchor = t_new;
        D->h_anchor = (t_new + h) - t_new;
    } else { /* Q = K.T @ P, common.py:363 */
        for (int k = 0; k < D->npol; ++k)
            for (int c = 0; c < n; ++c) {
                double acc = 0.0;
                for (int i = 0; i <= s; ++i)
                    if (T->P[i][k] != 0.0) acc = fma(T->P[i][k], L->K[i][c], acc);
                D->Q[k][c] = acc;
            }
    }
    for (int k = 0; k < D->npol; ++k)
        for (int c = 0; c < n; ++c) D->Q[k][c] *= D->h_anchor;
}

/* sol(te) of the step; out[c * stride] */
static void dense_eval(const lane_t* L, const dense_t* D, double t_new, const double* y_new,
                       double te, double* out, size_t stride) {
    const int n = L->n, s = L->T->s;
    if (D->cubic) { /* CubicDenseOutput, common.py:793-821 */
        const double hh = t_new - L->t;
        const double x = (te - L->t) / hh, omx = 1.0 - x;
        const double h00 = (1.0 + 2.0 * x) * (omx * omx);
        const double h10 = x * (omx * omx) * hh;
        const double h01 = (x * x) * (3.0 - 2.0 * x);
        const double h11 = (x * x) * (x - 1.0) * hh;
        for (int c = 0; c < n; ++c)
            out[(size_t)c * stride] =
                ((h00 * L->y[c] + h10 * L->K[0][c]) + h01 * y_new[c]) + h11 * L->K[s][c];
        return;
    }
    const double x = (te - D->t_anchor) / D->h_anchor;
    for (int c = 0; c < n; ++c) { /* Horner, common.py:781-785 */
        double v = D->Q[D->npol - 1][c] * x;
        for (int k = D->npol - 2; k >= 0; --k) v = (v + D->Q[k][c]) * x;
        out[(size_t)c * stride] = v + (D->anchor_end ? y_new[c] : L->y[c]);
    }
}

/* the t_eval points of (t_old, t_stop]; writes out[c*n_eval+i] */
static int emit_upto(lane_t* L, dense_t* D, double h, double t_new, const double* y_new,
                     double t_stop, const double* t_eval, int n_eval, int ieval, double* out) {
    if (ieval >= n_eval) return ieval;
    if (L->direction * (t_eval[ieval] - t_stop) > 0.0) return ieval;
    if (!D->built) dense_build(L, D, h, t_new, y_new);
    while (ieval < n_eval && L->direction * (t_eval[ieval] - t_stop) <= 0.0) {
        dense_eval(L, D, t_new, y_new, t_eval[ieval], out + ieval, (size_t)n_eval);
        ++ieval;
    }
    return ieval;
}

static int emit(lane_t* L, double h, double t_new, const double* y_new,
                const double* t_eval, int n_eval, int ieval, double* out) {
    dense_t D;
    D.built = 0;
    return emit_upto(L, &D, h, t_new, y_new, t_new, t_eval, n_eval, ieval, out);
}

/* ---- scipy's `events=` in the kernels' arithmetic (xsq_rk_core.cuh after_step,
 * events_now, brentq_dev; scipy/integrate/_ivp/ivp.py find_active_events,
 * handle_events, solve_event_equation; scipy.optimize brentq with
 * xtol = rtol = 4 eps).  Device arithmetic only: the reference's arithmetic with
 * events is oracle/rk_oracle.py. ---- */
typedef double (*event_fn)(int k, double t, const double* y, const double* p);
#define MAXEV 8
typedef struct ev_ctx_s {
    event_fn g;
    int n_events, capacity;
    int terminal[MAXEV], direction[MAXEV];
    double* t_events;   /* [n_events][capacity] */
    double* y_events;   /* [n_events][capacity][n] */
    int32_t* counts;    /* [n_events] */
    double g_old[MAXEV];
} ev_ctx_t;

event_fn xsq_oracle_builtin_events(int id);
static double ev_lorenz_sections(int k, double t, const double* y, const double* p) {
    (void)t; (void)p;
    if (k == 0) return y[2] - 27.0;
    if (k == 1) return y[0];
    return y[0] * y[1] - 30.0;
}

event_fn xsq_oracle_builtin_events(int id) { return id == 0 ? ev_lorenz_sections : 0; }

typedef struct {
    lane_t* L; const dense_t* D; const ev_ctx_t* E; const double* y_new; double t_new; int k;
} ev_root_ctx;
static double ev_of_t(const void* vc, double tt) {
    const ev_root_ctx* c = (const ev_root_ctx*)vc;
    double ytmp[MAXN];
    dense_eval(c->L, c->D, c->t_new, c->y_new, tt, ytmp, 1);
    return c->E->g(c->k, tt, ytmp, c->L->prm);
}
/* scipy.optimize.brentq(xtol = rtol = 4 eps) as the kernels evaluate it (brentq_dev);
 * shared with the SWAG restatement */
double xsq_oracle_brentq(double (*fn)(const void*, double), const void* c, double xa, double xb) {
    const double tol = 4.0 * 0x1.0p-52;
    double xpre = xa, xcur = xb;
    double xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = fn(c, xpre);
    double fcur = fn(c, xcur);
    if (fpre == 0.0) return xpre;
    if (fcur == 0.0) return xcur;
    if ((signbit(fpre) != 0) == (signbit(fcur) != 0)) return xcur;
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0.0 && fcur != 0.0 && (signbit(fpre) != 0) != (signbit(fcur) != 0)) {
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
            const double lim_a = fabs(spre), lim_b = 3 * fabs(sbis) - delta;
            if (2 * fabs(stry) < (lim_b < lim_a ? lim_b : lim_a)) {
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
        fcur = fn(c, xcur);
    }
    return xcur;
}

/* Everything solve_ivp does after solver.step() returned: events, then the t_eval
 * points of the step.  Returns 1 when a terminal event ends the trajectory; *t_new
 * / y_new are then the event point. */
static int events_after_step(lane_t* L, double h, double* t_new, double* y_new,
                             const double* t_eval, int n_eval, int* ieval, double* y_eval) {
    ev_ctx_t* E = L->ev;
    dense_t D;
    D.built = 0;
    double g_new[MAXEV], root[MAXEV];
    unsigned active = 0;
    for (int k = 0; k < E->n_events; ++k) {
        g_new[k] = E->g(k, *t_new, y_new, L->prm);
        const int up = E->g_old[k] <= 0.0 && g_new[k] >= 0.0;
        const int down = E->g_old[k] >= 0.0 && g_new[k] <= 0.0;
        const int d = E->direction[k];
        if ((up && d > 0) || (down && d < 0) || ((up || down) && d == 0)) active |= 1u << k;
    }
    int terminate = 0;
    double t_stop = *t_new;
    if (active) {
        dense_build(L, &D, h, *t_new, y_new);
        int any_term = 0;
        double r_star = 0.0;
        for (int k = 0; k < E->n_events; ++k) {
            if (!(active >> k & 1u)) continue;
            ev_root_ctx c = {L, &D, E, y_new, *t_new, k};
            root[k] = xsq_oracle_brentq(ev_of_t, &c, L->t, *t_new);
        }
        for (int k = 0; k < E->n_events; ++k) {
            if (!(active >> k & 1u)) continue;
            ++E->counts[k];
            if (E->terminal[k] > 0 && E->counts[k] >= E->terminal[k]) {
                if (!any_term || L->direction * (root[k] - r_star) < 0.0) r_star = root[k];
                any_term = 1;
            }
        }
        terminate = any_term;
        if (terminate) t_stop = r_star;
        for (int k = 0; k < E->n_events; ++k) {
            if (!(active >> k & 1u)) continue;
            if (terminate && L->direction * (root[k] - r_star) > 0.0) continue;
            const int slot = E->counts[k] - 1;
            if (slot < E->capacity) {
                const size_t base = (size_t)k * E->capacity + slot;
                E->t_events[base] = root[k];
                dense_eval(L, &D, *t_new, y_new, root[k], E->y_events + base * L->n, 1);
            }
        }
    }
    for (int k = 0; k < E->n_events; ++k) E->g_old[k] = g_new[k];
    if (n_eval > 0)
        *ieval = emit_upto(L, &D, h, *t_new, y_new, t_stop, t_eval, n_eval, *ieval, y_eval);
    if (terminate) {
        if (!D.built) dense_build(L, &D, h, *t_new, y_new);
        double ys[MAXN];
        dense_eval(L, &D, *t_new, y_new, t_stop, ys, 1);
        memcpy(y_new, ys, sizeof(double) * L->n);
        *t_new = t_stop;
    }
    return terminate;
}

static double wdot(const double* a, const double* b, const double* wt, int n) {
    wacc_t w;
    wacc_init(&w);
    for (int c = 0; c < n; ++c) { double* s = wacc_at(&w, c); *s = fma(a[c] / wt[c], b[c] / wt[c], *s); }
    return wacc_total(&w);
}
/* stiff_d: z ~ havg * J * v by a difference of f; returns <z, z> */
static double jac_times(lane_t* L, const double* v, double havg, double x, const double* y,
                        const double* fxy, const double* wt, double scale, double vdotv,
                        double* z) {
    const int n = L->n;
    const double temp1 = scale / sqrt(vdotv);
    double yp[MAXN];
    for (int c = 0; c < n; ++c) yp[c] = fma(temp1, v[c], y[c]);
    L->f(x, yp, L->prm, z);
    L->nfev++;
    const double q = havg / temp1;
    for (int c = 0; c < n; ++c) z[c] = q * (z[c] - fxy[c]);
    return wdot(z, z, wt, n);
}
/* stiff_b */
static int dominant_real(double v1v1, double v0v1, double v0v0, double* rold, double* rho,
                         double* root1, double* root2) {
    const double r = v0v1 / v0v0;
    *rho = fabs(r);
    const double det = v0v0 * v1v1 - v0v1 * v0v1;
    const double res = fabs(det / v0v0);
    const int rootre = det == 0.0 || (res <= 1e-6 * v1v1 && fabs(r - *rold) <= 0.001 * *rho);
    root1[0] = rootre ? r : 0.0; root1[1] = 0.0; root2[0] = root2[1] = 0.0;
    *rold = r;
    return rootre;
}
/* stiff_c: roots of x^2 + alpha x + beta */
static void quadratic_roots(double alpha, double beta, double* r1, double* r2) {
    r1[0] = r1[1] = r2[0] = r2[1] = 0.0;
    const double temp = alpha / 2;
    const double disc = temp * temp - beta;
    if (disc == 0.0) { r1[0] = r2[0] = -temp; return; }
    const double sqdisc = sqrt(fabs(disc));
    if (disc < 0.0) { r1[0] = r2[0] = -temp; r1[1] = sqdisc; r2[1] = -sqdisc; }
    else { r1[0] = temp > 0.0 ? -temp - sqdisc : -temp + sqdisc; r2[0] = beta / r1[0]; }
}
/* stiff_a.  returns stif: 1 true, 0 false, -1 unsure; *rootre: 1/0/-1;
 * *have_root: roots/rho valid */
static int stiff_probe(lane_t* L, double x, const double* y, double hnow, double havg,
                       double xend, int maxfcn, const double* wt, const double* fxy,
                       double* v0, int cost, int* rootre, int* have_root, double* root1,
                       double* root2, double* rho) {
    const int n = L->n;
    const double epsneg = 0x1.0p-53;
    *rootre = -1; *have_root = 0;
    if (fabs(hnow / havg) > 5 || fabs(hnow / havg) < 0.2) return 0;
    const double xtrfcn = cost * fabs((xend - x) / havg);
    if (xtrfcn <= maxfcn) return 0;
    double ynrm = sqrt(wdot(y, y, wt, n));
    const double sqrrmc = sqrt(epsneg);
    double scale = ynrm * sqrrmc;
    if (scale == 0.0) {
        ynrm = sqrt(wdot(v0, v0, wt, n));
        scale = ynrm * sqrrmc;
        if (scale == 0.0) return -1;
    }
    double v0v0 = wdot(v0, v0, wt, n);
    if (v0v0 == 0.0) {
        for (int c = 0; c < n; ++c) v0[c] = 1.0;
        v0v0 = wdot(v0, v0, wt, n);
    }
    const double v0nrm = sqrt(v0v0);
    for (int c = 0; c < n; ++c) v0[c] /= v0nrm;
    v0v0 = 1.0;
    double v1[MAXN], v2[MAXN], v3[MAXN], rold = 0.0;
    int converged = 0;
    for (int ntry = 0; ntry < 8; ++ntry) {
        const double v1v1 = jac_times(L, v0, havg, x, y, fxy, wt, scale, v0v0, v1);
        if (sqrt(v1v1) > 1.0e10 * sqrt(v0v0)) { *rootre = -1; return -1; }
        const double v0v1 = wdot(v0, v1, wt, n);
        if (ntry == 0) {
            rold = v0v1 / v0v0;
            if (fabs(rold) < cbrt(epsneg)) { *rootre = -1; return 0; }
        } else {
            if (dominant_real(v1v1, v0v1, v0v0, &rold, rho, root1, root2)) { *rootre = 1; converged = 1; break; }
            *rootre = 0;
        }
        const double v2v2 = jac_times(L, v1, havg, x, y, fxy, wt, scale, v1v1, v2);
        const double v0v2 = wdot(v0, v2, wt, n), v1v2 = wdot(v1, v2, wt, n);
        if (dominant_real(v2v2, v1v2, v1v1, &rold, rho, root1, root2)) { *rootre = 1; converged = 1; break; }
        *rootre = 0;
        const double det1 = v0v0 * v1v1 - v0v1 * v0v1;
        const double alpha1 = (-v0v0 * v1v2 + v0v1 * v0v2) / det1;
        const double beta1 = (v0v1 * v1v2 - v1v1 * v0v2) / det1;
        const double v3v3 = jac_times(L, v2, havg, x, y, fxy, wt, scale, v2v2, v3);
        const double v1v3 = wdot(v1, v3, wt, n), v2v3 = wdot(v2, v3, wt, n);
        if (dominant_real(v3v3, v2v3, v2v2, &rold, rho, root1, root2)) { *rootre = 1; converged = 1; break; }
        const double det2 = v1v1 * v2v2 - v1v2 * v1v2;
        const double alpha2 = (-v1v1 * v2v3 + v1v2 * v1v3) / det2;
        const double beta2 = (v1v2 * v2v3 - v2v2 * v1v3) / det2;
        const double res2 = fabs(v3v3 + v2v2 * (alpha2 * alpha2) + v1v1 * (beta2 * beta2) +
                                 2 * v2v3 * alpha2 + 2 * v1v3 * beta2 + 2 * v1v2 * alpha2 * beta2);
        if (res2 <= 1e-6 * v3v3) {
            double r1[2], r2[2];
            quadratic_roots(alpha1, beta1, r1, r2);
            quadratic_roots(alpha2, beta2, root1, root2);
            *rho = sqrt(root1[0] * root1[0] + root1[1] * root1[1]);
            const double D1 = (root1[0] - r1[0]) * (root1[0] - r1[0]) + (root1[1] - r1[1]) * (root1[1] - r1[1]);
            const double D2 = (root1[0] - r2[0]) * (root1[0] - r2[0]) + (root1[1] - r2[1]) * (root1[1] - r2[1]);
            if (sqrt(fmin(D1, D2)) <= 0.001 * *rho) { converged = 1; break; }
        }
        const double v3nrm = sqrt(v3v3);
        for (int c = 0; c < n; ++c) v0[c] = v3[c] / v3nrm;
        v0v0 = 1.0;
    }
    if (!converged) { *rootre = -1; return -1; }
    *have_root = 1;
    return -1;
}

/* called after every accepted step (state already advanced); errv = h * K^T E */
static void diagnose_stiffness(lane_t* L, const double* y_old, const double* errv, double h) {
    if (L->nfev_stiff_detect == 0) return;
    const otab_t* T = L->T;
    const int n = L->n;
    L->okstp += 1;
    L->havg = 0.9 * L->havg + 0.1 * h;
    if (L->okstp == 20) { L->havg = h; L->jflstp = 0; }
    int lotsfl = 0;
    if (L->okstp % 40 == 39) { lotsfl = L->jflstp >= 10; L->jflstp = 0; }
    const int many_steps = L->nfev_stiff_detect / T->s;
    const int toomch = L->okstp % many_steps == many_steps - 1;
    if (!(toomch || lotsfl)) return;
    double wt[MAXN], v0[MAXN];
    for (int c = 0; c < n; ++c) {
        wt[c] = fmax(0.5 * (fabs(L->y[c]) + fabs(y_old[c])), SQRT_TINY);
        v0[c] = errv[c];
    }
    int rootre, have_root;
    double root1[2], root2[2], rho = 0.0;
    int stif = stiff_probe(L, L->t, L->y, h, L->havg, L->t_bound, L->nfev_stiff_detect, wt,
                           L->fcur, v0, T->s, &rootre, &have_root, root1, root2, &rho);
    L->n_stiff_tests++;
    if (have_root) {
        rootre = root1[1] == 0.0;
        if (root1[0] > 0.0) stif = 0;
        else {
            const double rho2 = sqrt(root2[0] * root2[0] + root2[1] * root2[1]);
            if (rho2 >= 0.9 * rho && root2[0] > 0.0) stif = 0;
            else if (fabs(root1[1]) > fabs(root1[0]) * T->tanang) stif = -1;
            else stif = rho >= 0.9 * T->stbrad;
        }
    }
    if (stif < 0) { if (rootre == 0 && lotsfl) L->stiff_flags |= 4; }
    else if (stif == 1 && rootre >= 0) L->stiff_flags |= rootre ? 1 : 2;
}

/* ---- CKdisc._step_impl, cash.py:245-404 (xsq_rk_core.cuh attempt_ckdisc) ----
 * Reference arithmetic: norm(err/tol) ** (1/p) with sqrt and pow.  Device
 * arithmetic: (ss/n) ** (1/(2p)) through the kernels' own log2 / exp2, and
 * norm < 1 decided as ss < n. */
static double pymax(double a, double b) { return b > a ? b : a; }
static double pymin(double a, double b) { return b < a ? b : a; }
/* sel: 0/1 assessment pair, 2/3 fallback pair, 4 the fifth order pair.  Returns
 * sum((err/tol)^2) in device arithmetic, norm(err/tol) otherwise. */
static double ck_solution(lane_t* L, int sel, double h, double* sol) {
    const otab_t* T = L->T;
    const int ns = (sel == 0 || sel == 2) ? 2 : (sel == 4 ? 6 : 4);
    const double* wb = sel < 2 ? T->B_assess[sel] : (sel < 4 ? T->B_fallback[sel - 2] : T->B);
    const double* we = sel < 2 ? T->E_assess[sel] : (sel < 4 ? T->E_fallback[sel - 2] : T->E);
    double errv[MAXN];
    for (int c = 0; c < L->n; ++c) {
        sol[c] = fma(h, wsum(L, wb, ns, c), L->y[c]);
        errv[c] = h * wsum(L, we, ns, c);
    }
    return g_device_math ? scaled_ss_dev(L, errv, sol) : scaled_norm(L, errv, sol);
}
static double ck_root(const lane_t* L, double v, int p) {
    if (!g_device_math) return pow(v, 1.0 / (double)p);
    if (v == 0.0) return 0.0;
    if (!(v < INFINITY)) return v;
    const double inv2p = p == 2 ? 0.25 : (p == 3 ? 1.0 / 6.0 : 0.1);
    const double z = inv2p * (dev_log2(v) - L->log2n);
    if (!(fabs(z) < 1000.0)) return NAN;
    return dev_exp2(z);
}
static int ck_below_one(const lane_t* L, double v) {
    return g_device_math ? v < (double)L->n : v < 1.0;
}

static void ck_solve_one(lane_t* L, double tf, const double* t_eval, int n_eval, double* y_eval,
                         int max_steps, int* ieval_out, int* st_out) {
    const otab_t* T = L->T;
    const int n = L->n, s = T->s;
    double tw[2] = {1.5, 1.1}, q[2] = {100.0, 100.0};
    int ieval = 0, st = 1;
    while (st == 1) {
        int step_rejected = 0, accepted = 0;
        double y_new[MAXN], h = 0.0;
        reassess(L);
        while (!accepted) {
            if (L->h_abs < L->min_step) { st = ST_TOO_SMALL; break; }
            h = L->h_abs * L->direction;
            memcpy(L->K[0], L->fcur, sizeof(double) * n);
            rk_stage(L, h, 1);
            const double E1 = ck_root(L, ck_solution(L, 0, h, y_new), 2);
            double esttol = E1 / q[0];
            int retried = 0;
            if (E1 < tw[0] * q[0]) {
                rk_stage(L, h, 2);
                rk_stage(L, h, 3);
                const double E2 = ck_root(L, ck_solution(L, 1, h, y_new), 3);
                esttol = E2 / q[1];
                if (E2 < tw[1] * q[1]) {
                    rk_stage(L, h, 4);
                    rk_stage(L, h, 5);
                    double E4 = ck_root(L, ck_solution(L, 4, h, y_new), 5);
                    if (E4 == 0.0) E4 = 1e-160;
                    esttol = E4;
                    if (E4 < 1.0) {
                        accepted = 4;
                        double factor = pymin(T->ck_max_factor, T->ck_safety / E4);
                        if (step_rejected) factor = pymin(1.0, factor);
                        L->h_abs *= factor;
                        const double e12[2] = {E1, E2};
                        for (int j = 0; j < 2; ++j) {
                            double qq = e12[j] / E4;
                            if (qq > q[j]) qq = pymin(qq, 10 * q[j]);
                            else qq = pymax(qq, 2.0 / 3.0 * q[j]);
                            q[j] = pymax(1.0, pymin(10000.0, qq));
                        }
                        break;
                    }
                    if (!(E4 < INFINITY)) { st = ST_OVERFLOW; break; }
                    const double e12[2] = {E1, E2};
                    for (int i = 0; i < 2; ++i) {
                        const double EQ = e12[i] / q[i];
                        if (EQ < tw[i]) tw[i] = pymax(1.1, EQ);
                    }
                    if (E2 < 1.0 && ck_below_one(L, ck_solution(L, 3, h, y_new))) {
                        accepted = 2;
                        L->h_abs *= T->C_fallback[1];
                        h = L->h_abs * L->direction;
                        break;
                    }
                }
                if (E1 < 1.0) {
                    if (ck_below_one(L, ck_solution(L, 2, h, y_new))) {
                        accepted = 1;
                        L->h_abs *= T->C_fallback[0];
                        h = L->h_abs * L->direction;
                        break;
                    }
                    step_rejected = 1;
                    L->h_abs *= T->C_fallback[0];
                    L->n_rej++;
                    retried = 1;
                }
            }
            if (!retried) {
                step_rejected = 1;
                L->h_abs *= pymax(T->ck_min_factor, T->ck_safety / esttol);
                L->n_rej++;
            }
            if (L->n_acc + L->n_rej >= max_steps) { st = ST_BUDGET; break; }
        }
        if (st != 1) break;
        const double t_new = L->t + h;
        L->f(t_new, y_new, L->prm, L->K[s]);
        L->nfev++;
        L->force_cubic = accepted != 4;
        double t_end = t_new;
        int ev_stop = 0;
        if (L->ev) ev_stop = events_after_step(L, h, &t_end, y_new, t_eval, n_eval, &ieval, y_eval);
        else if (n_eval > 0) ieval = emit(L, h, t_new, y_new, t_eval, n_eval, ieval, y_eval);
        L->t = t_end;
        memcpy(L->y, y_new, sizeof(double) * n);
        memcpy(L->fcur, L->K[s], sizeof(double) * n);
        L->n_acc++;
        if (ev_stop) st = ST_EVENT;
        else if (L->direction * (L->t - tf) >= 0.0) st = ST_FINISHED;
        else if (L->n_acc + L->n_rej >= max_steps) st = ST_BUDGET;
    }
    *ieval_out = ieval;
    *st_out = st;
}

static _Thread_local ev_ctx_t* g_ev_lane = 0;

/* One trajectory: solve_ivp(fun, (t0, tf), y0, method=T, ...). */
static void rk_solve_one(const otab_t* T, rhs_fn f, int n, const double* y0,
                         const double* prm, double t0, double tf, double rtol,
                         const double* atol, double first_step, double max_step,
                         const double* sc, int interpolant, const double* t_eval,
                         int n_eval, double* y_eval, const double* h_forced,
                         int n_forced, int max_steps, double* t_final,
                         double* y_final, double* h_next, int32_t* n_acc,
                         int32_t* n_rej, int32_t* nfev, int32_t* status,
                         int32_t* n_eval_done, int nfev_stiff_detect,
                         int32_t* stiff_flags) {
    lane_t* L = (lane_t*)malloc(sizeof(lane_t));
    L->force_cubic = 0;
    L->ev = g_ev_lane;                 /* set per lane by xsq_oracle_rk_events_batch */
    L->nfev_stiff_detect = (T->stbrad > 0.0 && T->tanang > 0.0) ? nfev_stiff_detect : 0;
    L->jflstp = L->okstp = L->stiff_flags = L->n_stiff_tests = 0;
    L->havg = 0.0;
    const int s = T->s;
    L->T = T; L->f = f; L->prm = prm; L->n = n;
    L->rtol = rtol; L->atol = atol;
    L->t_bound = tf;
    L->direction = (tf != t0) ? (tf > t0 ? 1.0 : -1.0) : 1.0;
    L->max_step = max_step;
    const int oe = T->order2 < T->order ? T->order2 : T->order;
    L->err_exp = -1.0 / (oe + 1);
    L->minbeta1 = sc[0] * L->err_exp;
    L->minbeta2 = sc[1] * L->err_exp;
    L->minalpha = -sc[2];
    L->safety = sc[3];
    L->safety_sc = pow(sc[3], sc[0] + sc[1]);
    {   /* the same expressions as xsq_api.cu build_params */
        const double log2n = log2((double)n);
        L->log2n = log2n;
        L->ctl_a1s = 0.5 * L->err_exp;
        L->ctl_a0s = log2(L->safety) - L->ctl_a1s * log2n;
        L->ctl_a1c = 0.5 * L->minbeta1;
        L->ctl_a2c = 0.5 * L->minbeta2;
        L->ctl_a0c = log2(L->safety_sc) - (L->ctl_a1c + L->ctl_a2c) * log2n;
        L->l2_old = 0.0;
    }
    double cdiff = 1.0; /* common.py:129-137 */
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
            const double d = fabs(T->C[i] - T->C[j]);
            if (d != 0.0 && d < cdiff) cdiff = d;
        }
    if (cdiff < 1e-3) cdiff = 1e-3;
    L->h_min_a = 10 * 0x1.0p-53 / cdiff;
    L->interpolant = interpolant ? interpolant : IP_LOW;
    L->t = t0;
    memcpy(L->y, y0, sizeof(double) * n);
    L->n_acc = L->n_rej = 0;
    L->nfev = 1;
    f(t0, L->y, prm, L->fcur);
    L->standard_sc = 1;
    L->max_factor = 10.0;
    L->h_prev = 0.0; L->err_old = 0.0; L->min_step = 0.0;
    const int forced = n_forced > 0;
    if (forced) L->h_abs = h_forced[0];
    else if (first_step > 0.0) L->h_abs = first_step;
    else {
        const double b = t0 + L->direction * fmin(fabs(tf - t0), max_step);
        L->h_abs = h_start(L, t0, b, T->order2);
    }
    if (L->ev)
        for (int k = 0; k < L->ev->n_events; ++k) L->ev->g_old[k] = L->ev->g(k, t0, L->y, prm);
    int ieval = 0, st = 1, attempts = 0;
    const int early = T->variant == V_BS5 || T->variant == V_CFMR;
    const int fsal = T->E[s] != 0.0 || (T->variant == V_NYSTROM && T->Ep[s] != 0.0);
    if (!forced && t0 == tf) { /* scipy base.py:195-200 */
        for (int i = 0; i < n_eval; ++i)
            for (int c = 0; c < n; ++c) y_eval[(size_t)c * n_eval + i] = L->y[c];
        ieval = n_eval;
        st = ST_FINISHED;
    }
    if (T->variant == V_CKDISC && st == 1) ck_solve_one(L, tf, t_eval, n_eval, y_eval, max_steps, &ieval, &st);
    while (st == 1) {
        int step_rejected = 0;
        /* the step budget (max_steps: a safety net of the device path, no reference
         * analogue) ends the lane before the next step is prepared, as in the kernels */
        if (!forced && attempts >= max_steps) { st = ST_BUDGET; break; }
        if (!forced) reassess(L);
        for (;;) { /* while not step_accepted */
            if (forced) L->h_abs = h_forced[L->n_acc];
            else {
                if (L->h_abs < L->min_step) { st = ST_TOO_SMALL; break; }
                if (attempts >= max_steps) { st = ST_BUDGET; break; }
            }
            ++attempts;
            const double h = L->h_abs * L->direction;
            const double t_new = L->t + h;
            memcpy(L->K[0], L->fcur, sizeof(double) * n);
            const int nfirst = early ? s - 1 : s;
            for (int i = 1; i < nfirst; ++i) rk_stage(L, h, i);
            double y_new[MAXN], errv[MAXN];
            int err_dev_done = 0;
            (void)err_dev_done;
            if (g_device_math) {
                /* the kernel's attempt (xsq_rk_core.cuh attempt_rk / xsq_rk_fast.cuh) */
                const double NTOT = (double)n;
                double ss = 0.0;
                int pre_reject = 0;
                if (early) {
                    const double* wb = T->variant == V_BS5 ? T->B_scale_pre : T->A[s - 1];
                    const double* we = T->variant == V_BS5 ? T->E_pre : T->E;
                    for (int c = 0; c < n; ++c) {
                        y_new[c] = fma(h, wsum(L, wb, s - 1, c), L->y[c]);
                        errv[c] = h * wsum(L, we, s - 1, c);
                    }
                    ss = scaled_ss_dev(L, errv, y_new);
                    pre_reject = !forced && ss > NTOT &&
                        (ss >= NTOT * (1.0 + 0x1.0p-48) || sqrt(ss / NTOT) > 1.0);
                }
                if (!pre_reject && T->variant == V_NYSTROM) {
                    /* RungeKuttaNystrom._comp_sol_err / _estimate_error, common.py:1287-1309 */
                    const int nh = n / 2;
                    const double hh = h * h;
                    for (int c = 0; c < nh; ++c) {
                        y_new[c] = L->y[c] + (wsum(L, T->B, s, nh + c) * hh + h * L->y[nh + c]);
                        y_new[nh + c] = L->y[nh + c] + wsum(L, T->Bp, s, nh + c) * h;
                    }
                    if (fsal) { f(t_new, y_new, prm, L->K[s]); L->nfev++; }
                    for (int c = 0; c < nh; ++c) {
                        errv[c] = wsum(L, T->E, s + fsal, nh + c) * hh;
                        errv[nh + c] = wsum(L, T->Ep, s + fsal, nh + c) * h;
                    }
                    ss = scaled_ss_dev(L, errv, y_new);
                } else if (!pre_reject) {
                    if (early) rk_stage(L, h, s - 1);
                    for (int c = 0; c < n; ++c) y_new[c] = fma(h, wsum(L, T->B, s, c), L->y[c]);
                    if (fsal) { f(t_new, y_new, prm, L->K[s]); L->nfev++; }
                    for (int c = 0; c < n; ++c) errv[c] = h * wsum(L, T->E, s + fsal, c);
                    ss = scaled_ss_dev(L, errv, y_new);
                }
                const int accept = forced || (!pre_reject && ss < NTOT);
                const int bad = !forced && !pre_reject && !(ss < INFINITY);
                if (T->variant == V_BS5 && bad) { st = ST_OVERFLOW; break; }
                const int tiny = ss < NTOT * 0x1.0p-1022;
                const int second = accept && !L->standard_sc;
                const double l2 = dev_log2(ss);
                double factor;
                if (L->minalpha != 0.0) {
                    const double zx = second ? L->minalpha * log2_fast_dev(h / L->h_prev) : 0.0;
                    factor = ctl_factor_dev(L, l2, zx, 1, accept, second, step_rejected, tiny);
                } else {
                    factor = ctl_factor_dev(L, l2, 0.0, 0, accept, second, step_rejected, tiny);
                }
                if (bad || (pre_reject && !(ss < INFINITY))) factor = 0.2;
                if (!forced) L->h_abs *= factor;
                if (!accept) {
                    step_rejected = 1;
                    L->n_rej++;
                    L->jflstp++;
                    if (bad) { st = ST_OVERFLOW; break; }
                    continue;
                }
                if (!forced) {
                    L->standard_sc = tiny;
                    if (factor < 4.0) L->max_factor = 4.0;
                }
                L->l2_old = l2;
                err_dev_done = 1;
            }
            double err = 0.0;
            if (!g_device_math) {
            if (early) { /* bogacki.py:340-346, calvo.py:255-261 */
                const double* wb = T->variant == V_BS5 ? T->B_scale_pre : T->A[s - 1];
                const double* we = T->variant == V_BS5 ? T->E_pre : T->E;
                for (int c = 0; c < n; ++c) {
                    y_new[c] = fma(h, wsum(L, wb, s - 1, c), L->y[c]);
                    errv[c] = h * wsum(L, we, s - 1, c);
                }
                const double err_pre = scaled_norm(L, errv, y_new);
                if (!forced && err_pre > 1.0) {
                    step_rejected = 1;
                    L->h_abs *= fmax(0.2, L->safety * pow(err_pre, L->err_exp));
                    L->n_rej++;
                    if (L->nfev_stiff_detect) L->jflstp++;
                    continue;
                }
                rk_stage(L, h, s - 1);
            }
            for (int c = 0; c < n; ++c) y_new[c] = fma(h, wsum(L, T->B, s, c), L->y[c]);
            if (fsal) { f(t_new, y_new, prm, L->K[s]); L->nfev++; }
            for (int c = 0; c < n; ++c) errv[c] = h * wsum(L, T->E, s + fsal, c);
            err = scaled_norm(L, errv, y_new);
            if (!forced) {
                if (err < 1.0) { /* common.py:249-276 */
                    double factor;
                    if (err < SQRT_TINY) {
                        factor = L->max_factor;
                        L->standard_sc = 1;
                    } else if (L->standard_sc) {
                        factor = L->safety * pow(err, L->err_exp);
                        L->standard_sc = 0;
                    } else {
                        const double h_ratio = h / L->h_prev;
                        double fac = pow(err, L->minbeta1);
                        if (L->minbeta2 != 0.0) fac *= pow(L->err_old, L->minbeta2);
                        if (L->minalpha != 0.0) fac *= pow(h_ratio, L->minalpha);
                        factor = L->safety_sc * fac;
                        factor = fmin(L->max_factor, fmax(0.2, factor));
                    }
                    if (step_rejected) factor = fmin(1.0, factor);
                    L->h_abs *= factor;
                    if (factor < 4.0) L->max_factor = 4.0;
                } else {
                    const int bad = isnan(err) || isinf(err);
                    if (T->variant == V_BS5 && bad) { st = ST_OVERFLOW; break; }
                    step_rejected = 1;
                    L->h_abs *= fmax(0.2, L->safety * pow(err, L->err_exp));
                    L->n_rej++;
                    L->jflstp++;
                    if (bad) { st = ST_OVERFLOW; break; }
                    continue;
                }
            }
            }
            if (!fsal) { f(t_new, y_new, prm, L->K[s]); L->nfev++; }
            /* the probe of the stiffness diagnosis sees the step as taken, whatever the
             * events do with it afterwards (it runs inside solver.step()) */
            double t_end = t_new, y_end[MAXN];
            memcpy(y_end, y_new, sizeof(double) * n);
            int ev_stop = 0;
            if (L->ev) ev_stop = events_after_step(L, h, &t_end, y_end, t_eval, n_eval, &ieval, y_eval);
            else if (n_eval > 0) ieval = emit(L, h, t_new, y_new, t_eval, n_eval, ieval, y_eval);
            L->h_prev = h;
            L->err_old = err;
            L->t = t_new;
            double y_old[MAXN];
            memcpy(y_old, L->y, sizeof(double) * n);
            memcpy(L->y, y_new, sizeof(double) * n);
            memcpy(L->fcur, L->K[s], sizeof(double) * n);
            L->n_acc++;
            if (!forced) diagnose_stiffness(L, y_old, errv, h);
            if (ev_stop) {                        /* t, y = the event point (ivp.py) */
                L->t = t_end;
                memcpy(L->y, y_end, sizeof(double) * n);
                st = ST_EVENT;
            } else if (forced) { if (L->n_acc >= n_forced) st = ST_FINISHED; }
            else if (L->direction * (L->t - tf) >= 0.0) st = ST_FINISHED;
            break;
        }
    }
    for (int i = ieval; i < n_eval; ++i)
        for (int c = 0; c < n; ++c) y_eval[(size_t)c * n_eval + i] = NAN;
    *t_final = L->t;
    memcpy(y_final, L->y, sizeof(double) * n);
    if (h_next) *h_next = L->h_abs;
    *n_acc = L->n_acc; *n_rej = L->n_rej; *nfev = L->nfev; *status = st == ST_EVENT ? 1 : st;
    if (n_eval_done) *n_eval_done = ieval;
    if (stiff_flags) *stiff_flags = L->stiff_flags;
    free(L);
}

/* Batch entry: AoS inputs y0[N][n], params[N][p]; outputs y_final[N][n],
 * y_eval[N][n][n_eval].  rhs >= 0 selects a built-in; rhs < 0 uses `user_f`
 * (a C function pointer, e.g. a ctypes callback, single thread only). */
int xsq_oracle_rk_batch(const otab_t* T, int rhs, rhs_fn user_f, int n, int p,
                        int64_t n_lanes, const double* y0, const double* params,
                        double t0, double tf, double rtol, const double* atol,
                        double first_step, double max_step, const double* sc,
                        int interpolant, const double* t_eval, int n_eval,
                        double* y_eval, const double* h_forced, int n_forced,
                        int max_steps, double* t_final, double* y_final,
                        double* h_next, int32_t* n_acc, int32_t* n_rej,
                        int32_t* nfev, int32_t* status, int32_t* n_eval_done,
                        int n_threads, int nfev_stiff_detect, int32_t* stiff_flags) {
    rhs_fn f = rhs >= 0 ? builtin_rhs(rhs) : user_f;
    if (!f || n > MAXN || T->s >= MAXS) return -1;
    /* the Nystrom methods are restated in device arithmetic only (the reference's
     * arithmetic: oracle/rk_oracle.py, bit-identical to the reference) */
    if (T->variant == V_NYSTROM && (!g_device_math || (n & 1))) return -1;
    if (max_steps <= 0) max_steps = 2147483647;
    if (rhs < 0) n_threads = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_lanes; ++i) {
        rk_solve_one(T, f, n, y0 + i * n, params ? params + i * p : 0, t0, tf, rtol,
                     atol, first_step, max_step, sc, interpolant, t_eval, n_eval,
                     y_eval ? y_eval + (size_t)i * n * n_eval : 0, h_forced, n_forced,
                     max_steps, t_final + i, y_final + i * n, h_next ? h_next + i : 0,
                     n_acc + i, n_rej + i, nfev + i, status + i,
                     n_eval_done ? n_eval_done + i : 0, nfev_stiff_detect,
                     stiff_flags ? stiff_flags + i : 0);
    }
    return 0;
}

/* The same with scipy's `events=` (device arithmetic).  ev_set 0: the three Lorenz
 * section functions of oracle/problems.py EVENT_SETS["lorenz_sections"]; < 0:
 * `user_g` (a C function pointer, single thread).  t_events [N][n_events][capacity],
 * y_events [N][n_events][capacity][n] (NaN where unused), ev_count [N][n_events]. */
int xsq_oracle_rk_events_batch(const otab_t* T, int rhs, rhs_fn user_f, int n, int p,
                               int64_t n_lanes, const double* y0, const double* params,
                               double t0, double tf, double rtol, const double* atol,
                               double first_step, double max_step, const double* sc,
                               int interpolant, const double* t_eval, int n_eval,
                               double* y_eval, int max_steps, double* t_final, double* y_final,
                               double* h_next, int32_t* n_acc, int32_t* n_rej,
                               int32_t* nfev, int32_t* status, int32_t* n_eval_done,
                               int n_threads, int nfev_stiff_detect, int32_t* stiff_flags,
                               int ev_set, event_fn user_g, int n_events, const int32_t* terminal,
                               const int32_t* direction, int capacity, double* t_events,
                               double* y_events, int32_t* ev_count) {
    rhs_fn f = rhs >= 0 ? builtin_rhs(rhs) : user_f;
    event_fn g = ev_set == 0 ? ev_lorenz_sections : user_g;
    if (!f || !g || n > MAXN || T->s >= MAXS || n_events < 1 || n_events > MAXEV || capacity < 1 ||
        !g_device_math)
        return -1;
    if (max_steps <= 0) max_steps = 2147483647;
    if (rhs < 0 || ev_set < 0) n_threads = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_lanes; ++i) {
        ev_ctx_t E;
        E.g = g; E.n_events = n_events; E.capacity = capacity;
        for (int k = 0; k < n_events; ++k) { E.terminal[k] = terminal[k]; E.direction[k] = direction[k]; }
        E.t_events = t_events + (size_t)i * n_events * capacity;
        E.y_events = y_events + (size_t)i * n_events * capacity * n;
        E.counts = ev_count + (size_t)i * n_events;
        for (int k = 0; k < n_events; ++k) E.counts[k] = 0;
        g_ev_lane = &E;
        rk_solve_one(T, f, n, y0 + i * n, params ? params + i * p : 0, t0, tf, rtol,
                     atol, first_step, max_step, sc, interpolant, t_eval, n_eval,
                     y_eval ? y_eval + (size_t)i * n * n_eval : 0, 0, 0,
                     max_steps, t_final + i, y_final + i * n, h_next ? h_next + i : 0,
                     n_acc + i, n_rej + i, nfev + i, status + i,
                     n_eval_done ? n_eval_done + i : 0, nfev_stiff_detect,
                     stiff_flags ? stiff_flags + i : 0);
        g_ev_lane = 0;
    }
    return 0;
}

int xsq_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t xsq_oracle_tab_size(void) { return sizeof(otab_t); }

/* ---- shared with the SWAG / RKC restatements (other translation units) ---- */
rhs_fn xsq_oracle_builtin_rhs(int id) { return builtin_rhs(id); }

double xsq_oracle_h_start(rhs_fn f, const double* prm, int n, double a, double b,
                          const double* y, const double* yprime, int morder,
                          double rtol, const double* atol, int* nfev) {
    lane_t* L = (lane_t*)malloc(sizeof(lane_t));
    L->f = f; L->prm = prm; L->n = n; L->rtol = rtol; L->atol = atol; L->nfev = 0;
    memcpy(L->y, y, sizeof(double) * n);
    memcpy(L->fcur, yprime, sizeof(double) * n);
    const double h = h_start(L, a, b, morder);
    *nfev += L->nfev;
    free(L);
    return h;
}

/* element-wise access to the restated device functions (tests):
 * fn 0 rcp_scale, 1 log2, 2 exp2, 3 rcp64h */
int xsq_oracle_devmath(int fn, const double* x, double* out, int64_t n) {
    if (!xsq_rcp64h_delta && (fn == 0 || fn == 3)) return -1;
    for (int64_t i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = dev_rcp_scale(x[i]); break;
            case 1: out[i] = dev_log2(x[i]); break;
            case 2: out[i] = dev_exp2(x[i]); break;
            case 3: out[i] = dev_rcp64h(x[i]); break;
            default: return -1;
        }
    }
    return 0;
}

/* the controller restated (ctl_factor_dev) with explicit constants (tests) */
int xsq_oracle_ctl(const double* c, const double* l2, const double* l2_old, const double* zx,
                   const int* flags, const double* mf, double* out, int64_t n) {
    lane_t L;
    L.ctl_a1s = c[0]; L.ctl_a0s = c[1]; L.ctl_a1c = c[2]; L.ctl_a2c = c[3]; L.ctl_a0c = c[4];
    for (int64_t i = 0; i < n; ++i) {
        const int f = flags[i];
        L.l2_old = l2_old[i];
        L.max_factor = mf[i];
        out[i] = ctl_factor_dev(&L, l2[i], zx[i], (f & 16) != 0, f & 1, (f & 2) != 0, (f & 4) != 0,
                                (f & 8) != 0);
    }
    return 0;
}
