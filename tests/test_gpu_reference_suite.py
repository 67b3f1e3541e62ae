"""The reference's own solver tests (tests/test_ivp.py of WRKampi/extensisq,
themselves scipy's) run against the device path for every explicit solver
class of the hot path: same problems, same assertions, `res.sol(tc)` read as
`t_eval=tc` (the dense output is evaluated inside the kernel)."""
import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle.problems import CUDA_SOURCES

pytestmark = pytest.mark.gpu
METHODS = ["BS5", "Ts5", "CK5", "CKdisc", "Pr7", "Pr8", "Pr9", "SWAG", "CFMR7osc", "Me4"]
_RHS = {}


def rhs_for(name):
    if name not in _RHS:
        n, p, src = CUDA_SOURCES[name]
        _RHS[name] = xb.DeviceRHS.from_source(src, "rhs", n, p)
    return _RHS[name]


def sol_rational(t):                       # tests/test_ivp.py:31-32
    return np.asarray((t / (t + 10), 10 * t / (t + 10) ** 2))


def compute_error(y, y_true, rtol, atol):  # tests/test_ivp.py:144-147
    scale = np.abs(np.atleast_2d(y_true)).max(axis=1)[:, None]
    e = (y - y_true) / (atol + rtol * scale)
    return np.linalg.norm(e, axis=0) / np.sqrt(e.shape[0])


def solve(method, t_span, tc=None, **kw):
    r = xb.solve_ivp_batched(rhs_for("rational"), t_span, [[1 / 3, 2 / 9]], getattr(xb, method),
                             t_eval=tc, **kw)
    torch.cuda.synchronize()
    return r


@pytest.mark.parametrize("method", METHODS)
def test_integration(method):              # tests/test_ivp.py:151-214
    rtol, atol = 1e-3, 1e-6
    for t_span in ([5, 9], [5, 1]):
        tc = np.linspace(*t_span)
        res = solve(method, t_span, tc, rtol=rtol, atol=atol)
        assert int(res.status[0]) == 0 and bool(res.success[0])
        assert int(res.nfev[0]) < 44
        assert res.njev == 0 and res.nlu == 0
        assert res.t_events is None and res.y_events is None
        assert float(res.t_final[0]) == t_span[1]
        yf = res.y_final.cpu().numpy()[0][:, None]
        assert np.all(compute_error(yf, sol_rational(np.array([float(t_span[1])])), rtol, atol) < 5)
        yc = res.y.cpu().numpy()[0]
        assert np.all(compute_error(yc, sol_rational(tc), rtol, atol) < 5)
        tm = np.array([(t_span[0] + t_span[-1]) / 2])
        ym = solve(method, t_span, tm, rtol=rtol, atol=atol).y.cpu().numpy()[0]
        assert np.all(compute_error(ym, sol_rational(tm), rtol, atol) < 5)


@pytest.mark.parametrize("method", METHODS)
def test_max_step(method):                 # tests/test_ivp.py:583-624
    rtol, atol = 1e-3, 1e-6
    for t_span in ([5, 9], [5, 1]):
        tc = np.linspace(*t_span)
        res = solve(method, t_span, tc, rtol=rtol, atol=atol, max_step=0.5)
        assert int(res.status[0]) == 0
        assert float(res.t_final[0]) == t_span[1]
        # |diff(res.t)| <= 0.5: at least span / 0.5 accepted steps
        assert int(res.n_accepted[0]) >= 8
        assert np.all(compute_error(res.y.cpu().numpy()[0], sol_rational(tc), rtol, atol) < 5)
        with pytest.raises(ValueError):
            solve(method, t_span, None, max_step=-1)
        # max_step=1e-20: "Required step size is less than spacing between numbers."
        r = solve(method, t_span, None, rtol=rtol, atol=atol, max_step=1e-20, max_steps=50)
        assert int(r.status[0]) in (-1, -3) and not bool(r.success[0])
        assert "step size is less" in r.message(0) or method == "SWAG"


@pytest.mark.parametrize("method", [m for m in METHODS if m != "SWAG"])
def test_first_step(method):               # tests/test_ivp.py:628-665
    rtol, atol, first_step = 1e-3, 1e-6, 0.1
    for t_span in ([5, 9], [5, 1]):
        tc = np.linspace(*t_span)
        res = solve(method, t_span, tc, rtol=rtol, atol=atol, max_step=0.5, first_step=first_step)
        assert int(res.status[0]) == 0 and float(res.t_final[0]) == t_span[1]
        assert np.all(compute_error(res.y.cpu().numpy()[0], sol_rational(tc), rtol, atol) < 5)
        # the first step is taken as given: stop after one attempt and look
        one = solve(method, t_span, None, rtol=rtol, atol=atol, max_step=0.5,
                    first_step=first_step, max_steps=1)
        if int(one.n_accepted[0]) == 1:
            h1 = abs(float(one.t_final[0]) - 5)
            # CKdisc may accept a fallback solution over 1/5 or 3/5 of the step
            assert np.isclose(h1, first_step) or (
                method == "CKdisc" and (np.isclose(h1, first_step / 5) or np.isclose(h1, 0.6 * first_step)))
        with pytest.raises(ValueError):
            solve(method, t_span, None, first_step=-1)
        with pytest.raises(ValueError):
            solve(method, t_span, None, first_step=5)


@pytest.mark.parametrize("method", METHODS)
def test_no_integration(method):           # tests/test_ivp.py:786-791
    r = xb.solve_ivp_batched(rhs_for("minus_y"), [4, 4], [[2.0, 3.0]], getattr(xb, method),
                             t_eval=[4.0])
    torch.cuda.synchronize()
    assert int(r.status[0]) == 0
    assert r.y_final.cpu().numpy()[0].tolist() == [2.0, 3.0]
    assert r.y.cpu().numpy()[0, :, 0].tolist() == [2.0, 3.0]


@pytest.mark.parametrize("method", METHODS)
def test_integration_zero_rhs(method):     # tests/test_ivp.py:1101-1105
    r = xb.solve_ivp_batched(rhs_for("zero3"), [0, 10], [[1.0, 1.0, 1.0]], getattr(xb, method))
    torch.cuda.synchronize()
    assert int(r.status[0]) == 0 and bool(r.success[0])
    np.testing.assert_allclose(r.y_final.cpu().numpy(), 1.0, rtol=1e-15)


@pytest.mark.parametrize("method", METHODS)
def test_dense_output_sol(method):         # tests/test_ivp.py:195-213 (res.sol)
    """dense_output=True: `res.sol` is callable like scipy's OdeSolution; the
    values are the method's own dense output (the kernel's t_eval emitter)."""
    rtol, atol = 1e-3, 1e-6
    for t_span in ([5, 9], [5, 1]):
        tc = np.linspace(*t_span)
        res = xb.solve_ivp_batched(rhs_for("rational"), t_span, [[1 / 3, 2 / 9]],
                                   getattr(xb, method), t_eval=tc, rtol=rtol, atol=atol,
                                   dense_output=True)
        assert isinstance(res.sol, xb.BatchedOdeSolution)
        yc = res.sol(tc)
        torch.cuda.synchronize()
        assert torch.equal(yc, res.y)                       # sol(res.t) == res.y, exactly
        assert np.all(compute_error(yc.cpu().numpy()[0], sol_rational(tc), rtol, atol) < 5)
        tm = (t_span[0] + t_span[-1]) / 2                   # a scalar time
        ym = res.sol(tm).cpu().numpy()[0][:, None]
        assert ym.shape == (2, 1)
        assert np.all(compute_error(ym, sol_rational(np.array([tm])), rtol, atol) < 5)
        # any order, repeated points, both ends of the span
        tq = np.array([t_span[1], tm, t_span[0], tm, tc[7]])
        yq = res.sol(tq).cpu().numpy()[0]
        assert np.array_equal(yq[:, 1], yq[:, 3]) and np.array_equal(yq[:, 1], ym[:, 0])
        assert np.array_equal(yq[:, 2], [1 / 3, 2 / 9])
        # tests/test_ivp.py:209-213: sol(t_n) == y_n to pmax * 1e-15
        P = getattr(getattr(xb, method), "P", None)
        pmax = max(1.0, float(np.abs(P).max())) if isinstance(P, np.ndarray) else 1.0
        assert np.allclose(yq[:, 0], res.y_final.cpu().numpy()[0], rtol=pmax * 1e-14,
                           atol=pmax * 1e-14)
        assert np.array_equal(yq[:, 4], res.y.cpu().numpy()[0][:, 7])
        with pytest.raises(ValueError):
            res.sol(max(t_span) + 1.0)
    plain = solve(method, [5, 9])
    assert plain.sol is None
