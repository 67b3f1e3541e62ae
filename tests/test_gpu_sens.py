"""GPU: sens_forward as an ensemble client
against the reference's golden vectors and, per lane, the restated reference."""
import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from extensisq_b200.sensitivity import unscale
from oracle import sens_oracle as SO
from test_sens_golden import CASES, TABS, case_args, unhex

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_sens_forward_vs_reference_golden(c):
    atol, te = case_args(c)
    src = SO.PROBLEMS[c["problem"]][3]
    sens, yf, sol = xb.sens_forward(src, c["t_span"], [c["y0"]], np.array(c["dy0dp"]), c["p"],
                                    atol=atol, rtol=c["rtol"], method=getattr(xb, c["method"]),
                                    t_eval=te, max_steps=200000)
    torch.cuda.synchronize()
    assert int(sol.status[0]) == 0
    ny, npar = len(c["y0"]), len(c["p"])
    sg = unhex(c["sens"]).reshape(ny, npar)
    tol = 20 * c["rtol"]
    # the device integrates |p_j| S_j (exactly equivalent error test, different
    # rounding), so counts may move by a step on stability-limited problems
    assert abs(int(sol.nfev[0]) - c["nfev"]) <= 0.02 * c["nfev"] + 14
    np.testing.assert_allclose(yf.cpu().numpy()[0], unhex(c["yf"]), rtol=tol, atol=1e-12)
    scale = np.abs(sg).max(axis=0, keepdims=True)
    assert (np.abs(sens.cpu().numpy()[0] - sg) <= tol * np.abs(sg) + 1e-3 * tol * scale).all()
    if te is not None:
        y, S = unscale(sol.y, c["p"], ny)
        yg = unhex(c["y"]).reshape(c["y_shape"])
        np.testing.assert_allclose(y.cpu().numpy()[0], yg[:ny], rtol=tol, atol=10 * c["rtol"])
        Sg = yg[ny:].reshape(npar, ny, -1).transpose(1, 0, 2)
        np.testing.assert_allclose(S.cpu().numpy()[0], Sg, rtol=tol, atol=10 * c["rtol"])


def test_sens_forward_parameter_sweep_vs_oracle_and_finite_differences():
    """32 Van der Pol lanes with different mu: dy/dmu per lane against the
    restated reference and against a central difference of two plain solves."""
    N = 32
    mu = np.linspace(0.5, 4.0, N)[:, None]
    y0 = np.tile([2.0, 0.0], (N, 1))
    fun, jac, dfdp, src = SO.PROBLEMS["vanderpol"]
    kw = dict(rtol=1e-8, atol=1e-10)
    sens, yf, sol = xb.sens_forward(src, (0.0, 4.0), y0, np.zeros((2, 1)), mu, method=xb.Pr8, **kw)
    torch.cuda.synchronize()
    assert (sol.status == 0).all()
    s = sens.cpu().numpy()
    for i in range(0, N, 5):
        so, yo, _ = SO.sens_forward(TABS["Pr8"], fun, (0.0, 4.0), y0[i], jac, dfdp,
                                    np.zeros((2, 1)), mu[i], **kw)
        np.testing.assert_allclose(s[i], so, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(yf[i].cpu().numpy(), yo, rtol=1e-7, atol=1e-10)
    d = 1e-5
    a = xb.solve_ivp_batched("vanderpol", (0.0, 4.0), y0, xb.Pr8, params=mu + d, rtol=1e-11, atol=1e-12)
    b = xb.solve_ivp_batched("vanderpol", (0.0, 4.0), y0, xb.Pr8, params=mu - d, rtol=1e-11, atol=1e-12)
    fd = ((a.y_final - b.y_final) / (2 * d)).cpu().numpy()
    np.testing.assert_allclose(s[:, :, 0], fd, rtol=2e-5, atol=2e-6)


def test_sens_forward_argument_checks():
    src = SO.PROBLEMS["lorenz"][3]
    with pytest.raises(AssertionError):
        xb.sens_forward(src, (0.0, 1.0), [[1.0, 1.0, 1.0]], np.zeros((2, 3)), [10.0, 28.0, 2.0])
    with pytest.raises(ValueError):      # 300 * (1 + 3) states: beyond a warp-per-system kernel
        xb.sens_forward(src, (0.0, 1.0), [[1.0] * 300], np.zeros((300, 3)), [10.0, 28.0, 2.0])


CHAIN_SRC = """
// a chain of 6 compartments with 5 rate constants: y0 -> y1 -> ... -> y5
__device__ void fun(double t, const double* y, const double* p, double* dy) {
    dy[0] = -p[0] * y[0];
    for (int i = 1; i < 5; ++i) dy[i] = p[i - 1] * y[i - 1] - p[i] * y[i];
    dy[5] = p[4] * y[4];
}
__device__ void jac(double t, const double* y, const double* p, double* J) {
    for (int k = 0; k < 36; ++k) J[k] = 0.0;
    for (int i = 0; i < 5; ++i) { J[i * 6 + i] = -p[i]; J[(i + 1) * 6 + i] = p[i]; }
}
__device__ void dfdp(double t, const double* y, const double* p, double* D) {
    for (int k = 0; k < 30; ++k) D[k] = 0.0;
    for (int q = 0; q < 5; ++q) { D[q * 5 + q] = -y[q]; D[(q + 1) * 5 + q] = y[q]; }
}
"""


def test_sens_forward_beyond_16_states_runs_a_warp_per_system():
    """ny (1 + np) = 36 combined states: the generated right-hand side is
    compiled for the warp-per-system kernel (xsq_rhs.cuh WideSystem); against
    the restated reference per lane."""
    def fun(t, y, *p):
        p = np.asarray(p)
        d = np.empty(6)
        d[0] = -p[0] * y[0]
        d[1:5] = p[:4] * y[:4] - p[1:5] * y[1:5]
        d[5] = p[4] * y[4]
        return d

    def jac(t, y, *p):
        J = np.zeros((6, 6))
        for i in range(5):
            J[i, i] = -p[i]
            J[i + 1, i] = p[i]
        return J

    def dfdp(t, y, *p):
        D = np.zeros((6, 5))
        for q in range(5):
            D[q, q] = -y[q]
            D[q + 1, q] = y[q]
        return D

    N = 6
    P = np.array([1.0, 0.7, 1.3, 0.4, 0.9])[None, :] * (1.0 + 0.1 * np.arange(N))[:, None]
    y0 = np.tile([1.0, 0.0, 0.0, 0.0, 0.0, 0.0], (N, 1))
    kw = dict(rtol=1e-8, atol=1e-10)
    sens, yf, sol = xb.sens_forward(CHAIN_SRC, (0.0, 3.0), y0, np.zeros((6, 5)), P,
                                    method=xb.Ts5, **kw)
    torch.cuda.synchronize()
    assert (sol.status == 0).all()
    assert sol.y_final.shape == (N, 36)
    for i in range(N):
        so, yo, _ = SO.sens_forward(TABS["Ts5"], fun, (0.0, 3.0), y0[i], jac, dfdp,
                                    np.zeros((6, 5)), P[i], **kw)
        np.testing.assert_allclose(yf[i].cpu().numpy(), yo, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(sens[i].cpu().numpy(), so, rtol=1e-5, atol=1e-8)
    # first compartment: y0(t) = exp(-p0 t), dy0/dp0 = -t exp(-p0 t)
    np.testing.assert_allclose(sens[:, 0, 0].cpu().numpy(), -3.0 * np.exp(-3.0 * P[:, 0]), rtol=1e-6)


def test_sens_forward_combined_system_bit_identical_to_oracle():
    """The combined system [y, |p| S] that sens_forward generates and compiles
    (sensitivity.py:60-217: S' = J S + df/dp) against the C oracle in device
    arithmetic integrating the same system through a Python mirror of the
    generated right-hand side (same operations in the same order, libm's fma):
    counts, final combined state, next step -- bit for bit on every lane."""
    import ctypes
    import ctypes.util
    from oracle import c_oracle as CO
    from oracle import rk_oracle as O
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    fma = libm.fma
    N, ny, npar = 48, 2, 1
    mu = np.linspace(0.5, 4.0, N)[:, None]
    y0 = np.tile([2.0, 0.0], (N, 1))
    src = SO.PROBLEMS["vanderpol"][3]

    def combined(t, Y, p):                  # _combined_source(...) for Van der Pol
        y0_, y1_, p0 = float(Y[0]), float(Y[1]), float(p[0])
        J = [0.0, 1.0, -2 * p0 * y0_ * y1_ - 1, p0 * (1 - y0_ * y0_)]
        D = [0.0, (1 - y0_ * y0_) * y1_]
        fac = abs(p0) if p0 != 0.0 else 1.0
        out = [y1_, p0 * (1 - y0_ * y0_) * y1_ - y0_]
        for r in range(ny):
            acc = 0.0
            for k in range(ny):
                acc = fma(J[r * ny + k], float(Y[ny + k]), acc)
            out.append(fma(fac, D[r * npar], acc))
        return np.array(out)

    tabs = O.load_tableaux()
    for m, kw in ((xb.Ts5, dict(rtol=1e-7, atol=1e-9)), (xb.Pr8, dict(rtol=1e-9, atol=1e-11))):
        sens, yf, sol = xb.sens_forward(src, (0.0, 4.0), y0, np.zeros((ny, npar)), mu, method=m, **kw)
        torch.cuda.synchronize()
        total_y0 = np.concatenate([y0, np.zeros((N, ny * npar))], axis=1)
        with CO.device_math():
            o = CO.rk_batch(tabs[m.__name__], None, (0.0, 4.0), total_y0, params=mu, user_fn=combined,
                            user_fn_params=True, **kw)
        for k in ("n_accepted", "n_rejected", "nfev", "status"):
            g = getattr(sol, k).cpu().numpy()
            assert np.array_equal(g, o[k]), (m.__name__, k, np.flatnonzero(g != o[k])[:5])
        for k in ("t_final", "y_final", "h_next"):
            g = np.ascontiguousarray(getattr(sol, k).cpu().numpy())
            assert np.array_equal(g.view(np.uint64), np.ascontiguousarray(o[k]).view(np.uint64)), \
                (m.__name__, k)
