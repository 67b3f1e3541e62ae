"""GPU: the Runge-Kutta-Nystrom methods Fi4N, Fi5N, Mu5Nmb, MR6NN (SURVEY.md
section 8f, rank 4, second half; reference common.py:1207-1320, fine.py,
murua.py, mikkawy.py) against golden runs of the unmodified reference
(tests/golden/rkn_golden.json) and the NumPy restatement that reproduces them
bit for bit (tests/test_rkn_oracle.py).

Tolerances (fp64): forced step sequences 1e-12 relative (the device sums the
stages in index order with FMA, NumPy through dgemv); adaptive runs: the same
accepted steps and nfev, states within 100 x rtol of the reference's."""
import json
import os

import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle import rk_oracle as O
from oracle.problems_rkn import make_fun, nbody32_setup

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "rkn_golden.json")) as fh:
    CASES = json.load(fh)["cases"]
TABS = O.load_tableaux_rkn()
METHODS = [xb.Fi4N, xb.Fi5N, xb.Mu5Nmb, xb.MR6NN]

SOURCES = {
    "oscillator": (2, 1, """
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = -(p[0] * p[0]) * y[0];
}"""),
    "kepler": (4, 1, """
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    const double r2 = y[0] * y[0] + y[1] * y[1];
    const double r3 = r2 * sqrt(r2);
    dy[0] = y[2];
    dy[1] = y[3];
    dy[2] = -y[0] / r3;
    dy[3] = -y[1] / r3;
}"""),
    "damped": (2, 2, """
__device__ void rhs(double t, const double* y, const double* p, double* dy) {
    dy[0] = y[1];
    dy[1] = -p[0] * y[0] - p[1] * y[1];
}"""),
}
_RHS = {}


def rhs_for(problem):
    if problem in ("vanderpol", "arenstorf", "nbody32"):
        return problem
    if problem not in _RHS:
        n, p, src = SOURCES[problem]
        _RHS[problem] = xb.DeviceRHS.from_source(src, "rhs", n, p)
    return _RHS[problem]


def params_for(c):
    if c["problem"] == "nbody32":
        m, eps2, _ = nbody32_setup()
        return [np.concatenate([[eps2], m])]
    return [c["params"]]


def unhex(a):
    if a and isinstance(a[0], list):
        return np.array([[float.fromhex(v) for v in r] for r in a])
    return np.array([float.fromhex(v) for v in a])


def to_np(res):
    torch.cuda.synchronize()
    return {k: getattr(res, k).cpu().numpy()
            for k in ("t_final", "y_final", "n_accepted", "n_rejected", "nfev", "status")}


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_rkn_vs_reference_golden(c):
    opt = dict(c["options"])
    rtol = opt.get("rtol", 1e-3)
    y0 = unhex(c["y0"])
    g = to_np(xb.solve_ivp_batched(rhs_for(c["problem"]), c["t_span"], [y0], getattr(xb, c["method"]),
                                   params=params_for(c), **opt))
    assert g["status"][0] == 0
    t_ref, y_ref = unhex(c["t"]), unhex(c["y"])
    assert g["t_final"][0] == t_ref[-1]
    # the same step sequence: accepted steps and evaluations (the reference does
    # not count its stiffness probe's evaluations for these methods)
    assert abs(int(g["n_accepted"][0]) - (c["n_t"] - 1)) <= (0 if c["n_t"] < 1000 else c["n_t"] // 200)
    if c["n_t"] < 1000:
        assert int(g["nfev"][0]) == c["nfev"]
    scale = np.abs(y_ref[:, -1]) + 1e-3 * np.abs(y_ref[:, -1]).max()
    atol = np.asarray(opt.get("atol", 1e-6), dtype=float)
    assert (np.abs(g["y_final"][0] - y_ref[:, -1]) <= 100 * (atol + rtol * scale)).all()


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_rkn_forced_steps_vs_numpy_restatement(m):
    prob = "kepler"
    hs = 0.02 * (1.0 + 0.6 * np.sin(0.37 * np.arange(60))) + 1e-4
    y0 = np.array([[0.5, 0.0, 0.0, np.sqrt(3.0)], [0.7, 0.0, 0.0, np.sqrt(1.3 / 0.7)]])
    g = to_np(xb.solve_ivp_batched(rhs_for(prob), (0.0, 1.0), y0, m, params=[[0.5], [0.3]],
                                   forced_steps=hs))
    for i in range(2):
        r = O.rk_solve(TABS[m.__name__], make_fun(prob, [0.5]), (0.0, 1.0), y0[i], forced_h=hs)
        assert g["n_accepted"][i] == hs.size and g["nfev"][i] == r["nfev"]
        assert np.abs(g["y_final"][i] - r["y"][:, -1]).max() <= 1e-12 * np.abs(r["y"][:, -1]).max()
        assert abs(g["t_final"][i] - r["t"][-1]) <= 1e-15 * abs(r["t"][-1])


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_rkn_kepler_ensemble_counts(m):
    """2048 orbits of different eccentricity: energy is conserved to the
    tolerance, and a sample of lanes takes the steps the restated reference takes."""
    N = 2048
    e = np.linspace(0.0, 0.8, N)
    y0 = np.stack([1.0 - e, np.zeros(N), np.zeros(N), np.sqrt((1.0 + e) / (1.0 - e))], axis=1)
    kw = dict(rtol=1e-8, atol=1e-10)
    g = to_np(xb.solve_ivp_batched(rhs_for("kepler"), (0.0, 20.0), y0, m, params=e[:, None], **kw))
    assert (g["status"] == 0).all()
    yf = g["y_final"]
    energy = 0.5 * (yf[:, 2] ** 2 + yf[:, 3] ** 2) - 1.0 / np.hypot(yf[:, 0], yf[:, 1])
    assert np.abs(energy + 0.5).max() < 2e-5
    same = 0
    idx = list(range(0, N, 128))
    for i in idx:
        r = O.rk_solve(TABS[m.__name__], make_fun("kepler", []), (0.0, 20.0), y0[i], **kw)
        same += int(r["n_accepted"] == g["n_accepted"][i] and r["nfev"] == g["nfev"][i])
        assert np.abs(r["y"][:, -1] - yf[i]).max() <= 1e-5
    print(f"{m.__name__}: identical accepted/nfev on {same} of {len(idx)} sampled lanes")
    assert same >= len(idx) - 2


def test_rkn_nbody32_warp_per_system():
    """The N-body systems these methods exist for: 64 systems of 32 bodies,
    warp per system, MR6NN vs Pr8 at tight tolerance."""
    rng = np.random.default_rng(3)
    S = 64
    m = rng.uniform(0.5, 1.5, (S, 32)) / 32.0
    y0 = np.concatenate([rng.uniform(-1, 1, (S, 96)), rng.uniform(-0.3, 0.3, (S, 96))], axis=1)
    prm = np.concatenate([np.full((S, 1), 0.01), m], axis=1)
    a = to_np(xb.solve_ivp_batched("nbody32", (0.0, 0.5), y0, xb.MR6NN, params=prm, rtol=1e-8, atol=1e-10))
    b = to_np(xb.solve_ivp_batched("nbody32", (0.0, 0.5), y0, xb.Pr8, params=prm, rtol=1e-11, atol=1e-13))
    assert (a["status"] == 0).all()
    assert np.abs(a["y_final"] - b["y_final"]).max() <= 1e-6
    # velocity independent: 5 evaluations per step attempt + 1 per accepted step
    assert (a["nfev"] < b["nfev"]).all()


def test_rkn_argument_checks():
    with pytest.raises(AssertionError):              # odd number of states
        xb.solve_ivp_batched("lorenz63", (0.0, 1.0), [[1.0, 1.0, 1.0]], xb.Fi4N, params=[[10.0, 28.0, 2.6]])
    with pytest.raises(AssertionError):              # velocity dependent problem, MR6NN
        xb.solve_ivp_batched("vanderpol", (0.0, 1.0), [[2.0, 0.0]], xb.MR6NN, params=[[1.0]])
    with pytest.raises(ValueError):                  # no dense output on the device
        xb.solve_ivp_batched("vanderpol", (0.0, 1.0), [[2.0, 0.0]], xb.Fi5N, params=[[1.0]],
                             t_eval=[0.0, 0.5, 1.0])


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_rkn_bit_identical_to_oracle_in_device_arithmetic(m):
    """The Nystrom variant of rk_persistent against the C oracle's restatement in
    the kernels' arithmetic (oracle/xsq_oracle.c rkn_stage and the Nystrom
    solution / error block; that restatement takes the same steps as the NumPy
    one, which is bit-identical to the reference): accepted / rejected / nfev
    counts, final time and state, next step size -- bit for bit on every lane.
    Fi4N, Fi5N, Mu5Nmb on the built-in Van der Pol sweep and the perturbed
    Arenstorf orbits; MR6NN (velocity independent problems only) on Kepler orbits
    through a user right-hand side, a Python callback playing it in the oracle."""
    from oracle import c_oracle as CO
    from test_gpu_rk import vdp_lanes, arenstorf_lanes
    tab = TABS[m.__name__]
    threads = max(1, len(os.sched_getaffinity(0)))
    cases = []
    if m.velocity_dependent:
        y0, prm = vdp_lanes(512)
        keep = prm[:, 0] <= 30.0                       # explicit method: leave the stiff end out
        cases.append(("vanderpol", (0.0, 20.0), y0[keep], prm[keep], None))
        y0, prm = arenstorf_lanes(512)
        cases.append(("arenstorf", (0.0, 17.0652165601579625588917206249), y0, prm, None))
    else:
        ecc = np.linspace(0.05, 0.7, 48)
        y0 = np.stack([1.0 - ecc, 0 * ecc, 0 * ecc, np.sqrt((1.0 + ecc) / (1.0 - ecc))], 1)
        cases.append(("kepler", (0.0, 12.0), y0, np.zeros((len(ecc), 1)), make_fun("kepler", [0.0])))
    for prob, span, y0, prm, pyfun in cases:
        for kw in (dict(rtol=1e-8, atol=1e-10), dict(rtol=1e-5, atol=1e-7)):
            r = xb.solve_ivp_batched(rhs_for(prob), span, y0, m, params=prm, max_steps=1000000, **kw)
            torch.cuda.synchronize()
            with CO.device_math():
                if pyfun is None:
                    o = CO.rk_batch(tab, prob, span, y0, params=prm, n_threads=threads,
                                    nfev_stiff_detect=0, **kw)
                else:
                    o = CO.rk_batch(tab, None, span, y0, user_fn=pyfun, nfev_stiff_detect=0, **kw)
            assert int(r.n_accepted.min()) > 10 and int(r.n_rejected.sum()) > 0
            for k in ("n_accepted", "n_rejected", "nfev", "status"):
                g = getattr(r, k).cpu().numpy()
                bad = np.flatnonzero(g != o[k])
                assert bad.size == 0, (m.__name__, prob, kw, k, bad[:5], g[bad[:5]], o[k][bad[:5]])
            for k in ("t_final", "y_final", "h_next"):
                g = getattr(r, k).cpu().numpy()
                assert np.array_equal(_bits(g), _bits(o[k])), (m.__name__, prob, kw, k)
