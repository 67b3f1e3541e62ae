"""TEST INFRASTRUCTURE -- restatement of the reference's ``sens_forward``
(extensisq/sensitivity.py:60-217) on top of the restated solvers, plus the
sensitivity test problems (Python callables and their CUDA twins).
Only tests/ may import this."""
import numpy as np

from . import rk_oracle as RO


def sens_forward(tab, fun, t_span, y0, jac, dfdp, dy0dp, p, atol=1e-6, rtol=1e-3,
                 t_eval=None, **options):
    y0 = np.asarray(y0, dtype=float)
    p = np.asarray(p, dtype=float)
    Ny, Np = y0.size, p.size
    dy0dp = np.asarray(dy0dp, dtype=float)
    total_atol = np.empty((Np + 1) * Ny)                 # sensitivity.py:157-162
    total_atol[:Ny] = atol
    for i, _p in enumerate(p, start=1):
        factor = abs(_p) or 1.
        total_atol[i * Ny:(i + 1) * Ny] = atol / factor

    def total_fun(t, total_y):                           # sensitivity.py:165-170
        y = total_y[:Ny]
        s = total_y[Ny:].reshape(Ny, Np, order='F')
        dy = np.asarray(fun(t, y, *p))
        ds = np.asarray(jac(t, y, *p)) @ s + np.asarray(dfdp(t, y, *p))
        return np.concatenate([dy, ds.reshape(-1, order='F')])
    total_y0 = np.concatenate([y0, dy0dp.reshape(-1, order='F')])
    sol = RO.rk_solve(tab, total_fun, t_span, total_y0, atol=total_atol, rtol=rtol,
                      t_eval=t_eval, **options)
    yf = sol["y"][:Ny, -1]
    sensf = sol["y"][Ny:, -1].reshape(Ny, Np, order='F')
    return sensf, yf, sol


# ---- problems ------------------------------------------------------------------
def rob_fun(t, y, *p):                   # reference tests/test_sens.py:11-16
    y1, y2, y3 = y
    p1, p2, p3 = p
    return np.array([-p1 * y1 + p2 * y2 * y3,
                     p1 * y1 - p2 * y2 * y3 - p3 * y2 ** 2,
                     p3 * y2 ** 2])


def rob_jac(t, y, *p):                   # tests/test_sens.py:19-24
    y1, y2, y3 = y
    p1, p2, p3 = p
    return np.array([[-p1, p2 * y3, p2 * y2],
                     [p1, -p2 * y3 - 2 * p3 * y2, -p2 * y2],
                     [0., 2 * p3 * y2, 0.]])


def rob_dfdp(t, y, *p):                  # tests/test_sens.py:27-32
    y1, y2, y3 = y
    return np.array([[-y1, y2 * y3, 0.],
                     [y1, -y2 * y3, -y2 ** 2],
                     [0., 0., y2 ** 2]])


def lor_fun(t, y, *p):
    s, r, b = p
    return np.array([s * (y[1] - y[0]), y[0] * (r - y[2]) - y[1], y[0] * y[1] - b * y[2]])


def lor_jac(t, y, *p):
    s, r, b = p
    return np.array([[-s, s, 0.], [r - y[2], -1., -y[0]], [y[1], y[0], -b]])


def lor_dfdp(t, y, *p):
    return np.array([[y[1] - y[0], 0., 0.], [0., y[0], 0.], [0., 0., -y[2]]])


def vdp_fun(t, y, *p):
    return np.array([y[1], p[0] * (1 - y[0] ** 2) * y[1] - y[0]])


def vdp_jac(t, y, *p):
    return np.array([[0., 1.], [-2 * p[0] * y[0] * y[1] - 1, p[0] * (1 - y[0] ** 2)]])


def vdp_dfdp(t, y, *p):
    return np.array([[0.], [(1 - y[0] ** 2) * y[1]]])


PROBLEMS = {
    "robertson": (rob_fun, rob_jac, rob_dfdp, r"""
__device__ void fun(double t, const double* y, const double* p, double* dy) {
    dy[0] = -p[0] * y[0] + p[1] * y[1] * y[2];
    dy[1] = p[0] * y[0] - p[1] * y[1] * y[2] - p[2] * (y[1] * y[1]);
    dy[2] = p[2] * (y[1] * y[1]);
}
__device__ void jac(double t, const double* y, const double* p, double* J) {
    J[0] = -p[0];  J[1] = p[1] * y[2];                       J[2] = p[1] * y[1];
    J[3] = p[0];   J[4] = -p[1] * y[2] - 2 * p[2] * y[1];    J[5] = -p[1] * y[1];
    J[6] = 0.;     J[7] = 2 * p[2] * y[1];                   J[8] = 0.;
}
__device__ void dfdp(double t, const double* y, const double* p, double* D) {
    D[0] = -y[0];  D[1] = y[1] * y[2];   D[2] = 0.;
    D[3] = y[0];   D[4] = -y[1] * y[2];  D[5] = -(y[1] * y[1]);
    D[6] = 0.;     D[7] = 0.;            D[8] = y[1] * y[1];
}"""),
    "lorenz": (lor_fun, lor_jac, lor_dfdp, r"""
__device__ void fun(double t, const double* y, const double* p, double* dy) {
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = y[0] * (p[1] - y[2]) - y[1];
    dy[2] = y[0] * y[1] - p[2] * y[2];
}
__device__ void jac(double t, const double* y, const double* p, double* J) {
    J[0] = -p[0];        J[1] = p[0];  J[2] = 0.;
    J[3] = p[1] - y[2];  J[4] = -1.;   J[5] = -y[0];
    J[6] = y[1];         J[7] = y[0];  J[8] = -p[2];
}
__device__ void dfdp(double t, const double* y, const double* p, double* D) {
    D[0] = y[1] - y[0];  D[1] = 0.;    D[2] = 0.;
    D[3] = 0.;           D[4] = y[0];  D[5] = 0.;
    D[6] = 0.;           D[7] = 0.;    D[8] = -y[2];
}"""),
    "vanderpol": (vdp_fun, vdp_jac, vdp_dfdp, r"""
__device__ void fun(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = p[0] * (1 - y[0] * y[0]) * y[1] - y[0];
}
__device__ void jac(double t, const double* y, const double* p, double* J) {
    J[0] = 0.;                          J[1] = 1.;
    J[2] = -2 * p[0] * y[0] * y[1] - 1; J[3] = p[0] * (1 - y[0] * y[0]);
}
__device__ void dfdp(double t, const double* y, const double* p, double* D) {
    D[0] = 0.;
    D[1] = (1 - y[0] * y[0]) * y[1];
}"""),
}
