"""TEST INFRASTRUCTURE — Python right-hand sides of the golden-vector problems
(same expressions as ``tools/gen_golden.py`` fed to the reference) plus the
CUDA source of the ones that are not built into the library, so the GPU tests
can register them through ``xsq_rhs_register_source``.
"""
from math import sin

import numpy as np

from .rk_oracle import lorenz63, vanderpol, arenstorf, nbody  # noqa: F401


def rational(t, y):           # reference tests/test_ivp.py:19-21
    return np.array([y[1] / t,
                     y[1] * (y[0] + 2 * y[1] - 1) / (t * (y[0] - 1))])


def duffing(t, y):            # reference docs/Demo_BS5.ipynb
    return np.array([y[1], y[0] ** 3 / 6 - y[0] + 2 * sin(2.78535 * t)])


def forced_osc(t, y):         # reference docs/Demo_CFMR7osc.ipynb
    return np.array([y[1], -100. * y[0] + 99. * sin(t)])


def detest_b3(t, y):          # reference docs/Demo_CFMR7osc.ipynb
    return np.array([-y[0], y[0] - 2 * y[1] ** 2, y[1] ** 2])


def mass_spring_damper(t, y):  # reference docs/Demo_own_RK.ipynb
    return np.array([y[1], 1. - (y[0] + y[1] / 2)])


def detest_f2(t, y):          # reference docs/Cash_Karp.ipynb (DETEST F2)
    return np.array([(55 - 1.5 * y[0]) if (t % 2 >= 1.) else
                     (55 - 0.5 * y[0])])


def square_forced(t, y):      # oscillator driven by a square wave (non-smooth)
    return np.array([y[1], -y[0] + (1.0 if sin(3.0 * t) >= 0.0 else -1.0)])


def ballistic(t, y):          # falling body with linear drag (events: ground hit)
    return np.array([y[1], -9.81 - 0.1 * y[1]])


def linear(lam):
    return lambda t, y: np.array([lam * y[0], lam * y[1]])


def make_fun(problem, params):
    if problem == "lorenz63":
        return lorenz63(*params)
    if problem == "vanderpol":
        return vanderpol(params[0])
    if problem == "arenstorf":
        return arenstorf(params[0])
    if problem == "linear":
        return linear(params[0])
    return {"rational": rational, "duffing": duffing,
            "forced_osc": forced_osc, "detest_b3": detest_b3,
            "mass_spring_damper": mass_spring_damper,
            "detest_f2": detest_f2, "square_forced": square_forced,
            "ballistic": ballistic}[problem]


# CUDA device-function sources for xsq_rhs_register_source (entry name "rhs";
# signature fixed by include/xsq.h).
CUDA_SOURCES = {
    "rational": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1] / t;
    dy[1] = y[1] * (y[0] + 2 * y[1] - 1) / (t * (y[0] - 1));
}"""),
    "duffing": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = y[0] * y[0] * y[0] / 6 - y[0] + 2 * sin(2.78535 * t);
}"""),
    "forced_osc": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = -100. * y[0] + 99. * sin(t);
}"""),
    "detest_b3": (3, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = -y[0];
    dy[1] = y[0] - 2 * (y[1] * y[1]);
    dy[2] = y[1] * y[1];
}"""),
    "mass_spring_damper": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = 1. - (y[0] + y[1] / 2);
}"""),
    "detest_f2": (1, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    double r = fmod(t, 2.0);               // Python's t % 2: sign of the divisor
    if (r < 0) r += 2.0;
    dy[0] = (r >= 1.) ? (55 - 1.5 * y[0]) : (55 - 0.5 * y[0]);
}"""),
    "square_forced": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = -y[0] + (sin(3.0 * t) >= 0.0 ? 1.0 : -1.0);
}"""),
    "zero3": (3, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = 0.0; dy[1] = 0.0; dy[2] = 0.0;
}"""),
    "minus_y": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = -y[0]; dy[1] = -y[1];
}"""),
    "ballistic": (2, 0, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = -9.81 - 0.1 * y[1];
}"""),
    "linear": (2, 1, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = p[0] * y[0];
    dy[1] = p[0] * y[1];
}"""),
}


# ---- PDE problems for SSV2stab ------------------------------------------------
def heat2d_reaction(nx):
    """u_t = Lap(u) + u - u^3 on (0,1)^2, Dirichlet 0, nx x nx interior points,
    5-point stencil.  Returns (fun, y0, rho)."""
    h = 1.0 / (nx + 1)
    inv_h2 = (nx + 1.0) * (nx + 1.0)
    x = np.arange(1, nx + 1) * h
    y0 = np.outer(np.sin(np.pi * x), np.sin(np.pi * x)).reshape(-1)
    work = np.zeros((nx + 2, nx + 2))

    def fun(t, y):
        work[1:-1, 1:-1] = y.reshape(nx, nx)
        u = work[1:-1, 1:-1]
        lap = (((work[:-2, 1:-1] + work[2:, 1:-1]) +
                (work[1:-1, :-2] + work[1:-1, 2:])) - 4.0 * u) * inv_h2
        return (lap + (u - u * u * u)).reshape(-1)
    rho = 8.0 * inv_h2 + 2.0
    return fun, y0, rho


def heat3d_notebook(N=39):
    """The linear 3-D heat problem of the reference's docs/Demo_SSV2stab.ipynb
    (cells 7-9), restated; returns (fun, y0, rho)."""
    def solution(x, y, z, t):
        return np.tanh(5 * x + 10 * y + 7.5 * z - (2.5 + 5 * t))

    def src(x, y, z, t):
        s = solution(x, y, z, t)
        return 362.5 * (s - s ** 3) + 5 * s ** 2 - 5
    g = np.linspace(0., 1., N + 2)
    X, Y, Z = np.meshgrid(g, g, g)
    W = solution(X, Y, Z, 0)
    y0 = W[1:-1, 1:-1, 1:-1].copy().reshape(-1)
    h = 1. / (N + 1.)

    def fun(t, y):
        W[0, :, :] = solution(X[0, :, :], Y[0, :, :], Z[0, :, :], t)
        W[-1, :, :] = solution(X[-1, :, :], Y[-1, :, :], Z[-1, :, :], t)
        W[:, 0, :] = solution(X[:, 0, :], Y[:, 0, :], Z[:, 0, :], t)
        W[:, -1, :] = solution(X[:, -1, :], Y[:, -1, :], Z[:, -1, :], t)
        W[:, :, 0] = solution(X[:, :, 0], Y[:, :, 0], Z[:, :, 0], t)
        W[:, :, -1] = solution(X[:, :, -1], Y[:, :, -1], Z[:, :, -1], t)
        W[1:-1, 1:-1, 1:-1] = y.reshape(N, N, N)
        lap = (1. / h ** 2) * (-6 * W[1:-1, 1:-1, 1:-1] +
                               W[:-2, 1:-1, 1:-1] + W[2:, 1:-1, 1:-1] +
                               W[1:-1, :-2, 1:-1] + W[1:-1, 2:, 1:-1] +
                               W[1:-1, 1:-1, :-2] + W[1:-1, 1:-1, 2:])
        return (lap + src(X, Y, Z, t)[1:-1, 1:-1, 1:-1]).reshape(-1)
    return fun, y0, 12 / h ** 2


# ---- event functions (scipy solve_ivp `events=`), Python and CUDA twins -------
# name -> (python callables [g_k(t, y)], CUDA source of
#          `double event(int k, double t, const double* y, const double* p)`)
EVENT_SETS = {
    "ground": ([lambda t, y: y[0], lambda t, y: y[1]], r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    return k == 0 ? y[0] : y[1];
}"""),
    "lorenz_sections": ([lambda t, y: y[2] - 27.0, lambda t, y: y[0],
                         lambda t, y: y[0] * y[1] - 30.0], r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[2] - 27.0;
    if (k == 1) return y[0];
    return y[0] * y[1] - 30.0;
}"""),
    # reference tests/test_ivp.py:371-379 (event_rational_1..3) and :759-760
    "rational": ([lambda t, y: y[0] - y[1] ** 0.7, lambda t, y: y[1] ** 0.6 - y[0],
                  lambda t, y: t - 7.4], r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[0] - pow(y[1], 0.7);
    if (k == 1) return pow(y[1], 0.6) - y[0];
    return t - 7.4;
}"""),
    "early": ([lambda t, y: t - 7.0], r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    return t - 7.0;
}"""),
    "vdp_cross": ([lambda t, y: y[0], lambda t, y: y[1] - 1.0,
                   lambda t, y: t - 7.25], r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[0];
    if (k == 1) return y[1] - 1.0;
    return t - 7.25;
}"""),
}


# The same 3-D heat problem as ONE device function per component (the general
# right-hand side of extensisq_b200.PdeRHS.from_vector_source): y index
# i = (a N + b) N + c, interior point (x, y, z) = (g[b+1], g[a+1], g[c+1]) of
# numpy's meshgrid(g, g, g); boundary values from the exact solution.
HEAT3D_VECTOR_SRC = r"""
#define NG %d
__device__ __forceinline__ double h3_sol(double x, double y, double z, double t) {
    return tanh(5 * x + 10 * y + 7.5 * z - (2.5 + 5 * t));
}
__device__ double heat3d(int i, double t, const double* u, const double* p) {
    const int c = i %% NG, b = (i / NG) %% NG, a = i / (NG * NG);
    const double h = 1.0 / (NG + 1.0);
    const double x = (b + 1) * h, y = (a + 1) * h, z = (c + 1) * h;
    const double w = u[i];
    const double am = a > 0 ? u[i - NG * NG] : h3_sol(x, 0.0, z, t);
    const double ap = a < NG - 1 ? u[i + NG * NG] : h3_sol(x, 1.0, z, t);
    const double bm = b > 0 ? u[i - NG] : h3_sol(0.0, y, z, t);
    const double bp = b < NG - 1 ? u[i + NG] : h3_sol(1.0, y, z, t);
    const double cm = c > 0 ? u[i - 1] : h3_sol(x, y, 0.0, t);
    const double cp = c < NG - 1 ? u[i + 1] : h3_sol(x, y, 1.0, t);
    const double lap = (1.0 / (h * h)) * (-6 * w + am + ap + bm + bp + cm + cp);
    const double s = h3_sol(x, y, z, t);
    return lap + (362.5 * (s - s * s * s) + 5 * (s * s) - 5);
}
"""


# The two-species combustion problem of docs/Demo_SSV2stab.ipynb (cell 1) as one
# device function per component: y = [c (N^3), T (N^3)], Neumann boundaries at
# the low faces (ghost = first interior value), Dirichlet 1 at the high faces,
# h = 1 / (N + 0.5); p = (L, alpha, delta, D).
COMBUSTION_VECTOR_SRC = r"""
#define NG %d
__device__ __forceinline__ double comb_lap(const double* A, int j, int a, int b, int c, double inv_h2) {
    const double w = A[j];
    const double am = a > 0 ? A[j - NG * NG] : w, ap = a < NG - 1 ? A[j + NG * NG] : 1.0;
    const double bm = b > 0 ? A[j - NG] : w, bp = b < NG - 1 ? A[j + NG] : 1.0;
    const double cm = c > 0 ? A[j - 1] : w, cp = c < NG - 1 ? A[j + 1] : 1.0;
    return inv_h2 * (-6 * w + am + ap + bm + bp + cm + cp);
}
__device__ double combustion(int i, double t, const double* y, const double* p) {
    const int n3 = NG * NG * NG;
    const int j = i < n3 ? i : i - n3;
    const int c = j %% NG, b = (j / NG) %% NG, a = j / (NG * NG);
    const double h = 1.0 / (NG + 0.5);
    const double inv_h2 = 1.0 / (h * h);
    const double conc = y[j], T = y[n3 + j];
    const double Dce = p[3] * conc * exp(-p[2] / T);
    if (i < n3) return comb_lap(y, j, a, b, c, inv_h2) - Dce;
    return (comb_lap(y + n3, j, a, b, c, inv_h2) + p[1] * Dce) / p[0];
}
"""
