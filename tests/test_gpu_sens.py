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
    with pytest.raises(ValueError):      # 6 * (1 + 3) states do not fit one lane
        xb.sens_forward(src, (0.0, 1.0), [[1.0] * 6], np.zeros((6, 3)), [10.0, 28.0, 2.0])
