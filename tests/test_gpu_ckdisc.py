"""GPU parity of CKdisc (cash.py:115-416),
through the C ABI, against the reference's golden vectors and the NumPy
restatement (bit-identical to the reference) on seeded ensembles."""
import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from extensisq_b200 import _lib
from oracle import rk_oracle as RO
from oracle.problems import CUDA_SOURCES, make_fun
from test_ckdisc_golden import CASES, TAB, ck_options, ck_t_eval, unhex

pytestmark = pytest.mark.gpu
BUILTIN_PROBLEMS = {"lorenz63", "vanderpol", "arenstorf"}
_RHS = {}


def rhs_for(problem):
    if problem in BUILTIN_PROBLEMS:
        return problem
    if problem not in _RHS:
        n, p, src = CUDA_SOURCES[problem]
        _RHS[problem] = xb.DeviceRHS.from_source(src, "rhs", n, p)
    return _RHS[problem]


def close(a, b, rtol, atol):
    return (np.abs(a - b) <= rtol * np.abs(b) + atol).all()


def oracle_sensitivity(problem, params, t_span, y0, t_eval, opts, n_noise=6):
    """How much the REFERENCE algorithm itself moves under last-digit changes:
    (a) its error norms scaled by 1 +- 1e-15 (the device evaluates
    norm**(1/p) as exp2(log2(.)/2p)), (b) every RHS value changed by at most
    one ulp (FMA contraction and summation order differ on the device).  The
    fifth order error estimate is a difference of nearly equal terms, so (b)
    moves E4 -- and through quit = E/E4 the later assessments -- by up to
    ~1e-7 relative.  Non-smooth right-hand sides put steps on top of the
    discontinuities, where one flipped assessment changes the step sequence.
    Returns (max |d nfev|, max |d NFS|, max |d y|) over the perturbed runs."""
    fun = make_fun(problem, params)
    orig = RO.norm
    base = RO.rk_solve(TAB, fun, t_span, y0, t_eval=t_eval, **opts)
    runs = []
    try:
        for eps in (1e-15, -1e-15, 4e-15):
            RO.norm = (lambda x, e=eps: orig(x) * (1 + e))
            runs.append(RO.rk_solve(TAB, fun, t_span, y0, t_eval=t_eval, **opts))
    finally:
        RO.norm = orig
    for seed in range(n_noise):
        rng = np.random.default_rng(seed)

        def noisy(t, y, rng=rng):
            f = np.asarray(fun(t, y), dtype=float)
            return f * (1.0 + 1.1102230246251565e-16 * rng.integers(-1, 2, f.shape))
        runs.append(RO.rk_solve(TAB, noisy, t_span, y0, t_eval=t_eval, **opts))
    return (max(abs(r["nfev"] - base["nfev"]) for r in runs),
            max(abs(r["n_rejected"] - base["n_rejected"]) for r in runs),
            max(np.abs(r["y"] - base["y"]).max() if t_eval is not None
                else np.abs(r["y_final"] - base["y_final"]).max() for r in runs))


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_ckdisc_vs_reference_golden(c):
    o = ck_options(c)
    te = ck_t_eval(c)
    r = xb.solve_ivp_batched(rhs_for(c["problem"]), c["t_span"], [c["y0"]], xb.CKdisc,
                             params=[c["params"]] if c["params"] else None,
                             t_eval=te, max_steps=100000, **o)
    torch.cuda.synchronize()
    assert int(r.status[0]) == 0
    t_g, y_g = unhex(c["t"]), unhex(c["y"])
    rtol = o.get("rtol", 1e-3)
    atol = np.max(np.atleast_1d(o.get("atol", 1e-6)))
    nfev, nfs = int(r.nfev[0]), int(r.n_rejected[0])
    same = nfev == c["nfev"] and nfs == c["nfs"]
    d_nfev, d_nfs, d_y = oracle_sensitivity(c["problem"], c["params"], c["t_span"], c["y0"],
                                            te, o)
    # identical decisions wherever the reference itself is insensitive to the
    # last digit of its error norms; elsewhere within a few times its own spread
    assert abs(nfev - c["nfev"]) <= 3 * d_nfev + (0 if d_nfev == 0 and d_nfs == 0 else 6), \
        (nfev, c["nfev"], d_nfev)
    assert abs(nfs - c["nfs"]) <= 3 * d_nfs + (0 if d_nfev == 0 and d_nfs == 0 else 3), \
        (nfs, c["nfs"], d_nfs)
    tol = max(1e-9, 100 * 2.3e-16 * len(t_g))
    slack = 10 * atol * (0 if same else 1) + 10 * d_y
    yf = r.y_final.cpu().numpy()[0]
    if te is None:
        if same:
            assert int(r.n_accepted[0]) == len(t_g) - 1
        assert close(yf, y_g[:, -1], tol, slack + 1e-300), (yf, y_g[:, -1], d_y)
    else:
        y = r.y.cpu().numpy()[0]
        assert int(r.n_eval_done[0]) == te.size
        assert close(y, y_g, tol, slack + 1e-300), (np.abs(y - y_g).max(), d_y)
    assert float(r.t_final[0]) == c["t_span"][1]


def test_ckdisc_ensemble_vs_numpy_oracle():
    """48 lanes of the non-smooth DETEST F2 problem with different initial
    values: per-lane counts and end states against the restated reference."""
    N = 48
    rng = np.random.default_rng(7)
    y0 = rng.uniform(60.0, 140.0, (N, 1))
    kw = dict(rtol=1e-6, atol=1e-8)
    r = xb.solve_ivp_batched(rhs_for("detest_f2"), (0.0, 6.0), y0, xb.CKdisc,
                             max_steps=100000, **kw)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all()
    fun = make_fun("detest_f2", [])
    same = stable = 0
    for i in range(N):
        o = RO.rk_solve(TAB, fun, (0.0, 6.0), y0[i], **kw)
        d_nfev, d_nfs, d_y = oracle_sensitivity("detest_f2", [], (0.0, 6.0), y0[i], None, kw)
        eq = (int(r.nfev[i]) == o["nfev"] and int(r.n_rejected[i]) == o["n_rejected"]
              and int(r.n_accepted[i]) == o["n_accepted"])
        same += eq
        stable += d_nfev == 0 and d_nfs == 0
        err = abs(float(r.y_final[i, 0]) - o["y_final"][0])
        tol = 1e-9 * abs(o["y_final"][0]) + 10 * d_y
        assert err <= (tol if eq else max(tol, 1e-4 * abs(o["y_final"][0]))), (i, err, d_y)
        assert abs(int(r.nfev[i]) - o["nfev"]) <= max(3 * d_nfev + 6, 0.06 * o["nfev"])
    # The exact statement -- every lane bit-identical to the C oracle in the
    # kernel's arithmetic -- is tests/test_gpu_exact.py::
    # test_ckdisc_nonsmooth_ensemble_bit_identical_and_cost_of_arithmetic.  Here
    # the other arithmetic (NumPy = the reference) is the yardstick, so the
    # fraction of identical step sequences measures how sensitive the problem
    # is (non-smooth on purpose), and is reported, not asserted.
    print("identical step sequences:", same, "of", N, "; reference-stable lanes:", stable)


def test_ckdisc_lowers_its_order_at_discontinuities():
    """The point of the method (docs/Cash_Karp.ipynb): on DETEST F2 it needs
    fewer evaluations than the fixed-order CK5 for the same tolerance."""
    kw = dict(rtol=1e-13, atol=1e-6, max_steps=100000)
    a = xb.solve_ivp_batched(rhs_for("detest_f2"), (0.0, 10.0), [[110.0]], xb.CKdisc, **kw)
    b = xb.solve_ivp_batched(rhs_for("detest_f2"), (0.0, 10.0), [[110.0]], xb.CK5,
                             sc_params="standard", **kw)
    torch.cuda.synchronize()
    assert int(a.status[0]) == 0 and int(b.status[0]) == 0
    assert int(a.nfev[0]) < int(b.nfev[0])
    assert abs(float(a.y_final[0, 0]) - float(b.y_final[0, 0])) < 1e-3


def test_ckdisc_options_and_tableau_image():
    with pytest.raises(ValueError):
        xb.solve_ivp_batched("lorenz63", (0.0, 1.0), [[1.0, 1.0, 1.0]], xb.CKdisc,
                             params=[[10.0, 28.0, 8 / 3]], forced_steps=[0.1, 0.1])
    # the device image of the coefficients equals the class attributes
    import ctypes as C
    lib = _lib.load()
    t = _lib.XsqTableau()
    assert lib.xsq_tableau_get(_lib.METHOD_IDS["CKdisc"], C.byref(t)) == 0
    s = t.n_stages
    assert s == 6 and t.order == 5 and t.order_secondary == 4
    A = np.array([[t.A[i][j] for j in range(s)] for i in range(s)])
    assert np.array_equal(A, xb.CKdisc.A)
    assert np.array_equal(np.array([t.B[i] for i in range(s)]), xb.CKdisc.B)
    assert np.array_equal(np.array([t.E[i] for i in range(s + 1)]), xb.CKdisc.E)
    # stiffness diagnosis is off for CKdisc whatever is asked (cash.py:238-240)
    r = xb.solve_ivp_batched("vanderpol", (0.0, 5.0), [[2.0, 0.0]], xb.CKdisc,
                             params=[[100.0]], nfev_stiff_detect=100, max_steps=200000)
    torch.cuda.synchronize()
    assert int(r.stiff_flags[0]) == 0
