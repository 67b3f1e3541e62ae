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
            "mass_spring_damper": mass_spring_damper}[problem]


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
    "linear": (2, 1, r"""
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = p[0] * y[0];
    dy[1] = p[0] * y[1];
}"""),
}
