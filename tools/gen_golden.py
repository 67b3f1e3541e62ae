#!/usr/bin/env python
"""Dev-time tool: run the UNMODIFIED reference (``/root/reference``, imported,
never copied) through ``scipy.integrate.solve_ivp`` on a fixed list of cases
and store inputs + outputs in ``tests/golden/rk_golden.npz``.

These vectors pin the oracle (``oracle/rk_oracle.py``, ``oracle/xsq_oracle.c``)
and, through it, the CUDA path.  ``/root/reference`` does not exist on the GPU
box, so nothing at test time reads it; only this script does.

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden.py
"""
import json
import os
import sys
import warnings
from math import sin

import numpy as np
from scipy.integrate import solve_ivp

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import extensisq as ref  # noqa: E402
from extensisq.common import RungeKutta  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "rk_golden.npz")

METHODS = ["Ts5", "BS5", "CK5", "Me4", "Pr7", "Pr8", "Pr9", "CFMR7osc"]


# ---- problems: name -> (python fun factory, n) ---------------------------
def lorenz(p):
    s, r, b = p
    return lambda t, y: [s * (y[1] - y[0]), y[0] * (r - y[2]) - y[1],
                         y[0] * y[1] - b * y[2]]


def vdp(p):
    mu = p[0]
    return lambda t, y: [y[1], mu * (1.0 - y[0] * y[0]) * y[1] - y[0]]


def rational(p):          # reference tests/test_ivp.py:19-21
    return lambda t, y: [y[1] / t, y[1] * (y[0] + 2 * y[1] - 1) /
                         (t * (y[0] - 1))]


def duffing(p):           # docs/Demo_BS5.ipynb
    return lambda t, y: [y[1], y[0] ** 3 / 6 - y[0] + 2 * sin(2.78535 * t)]


def forced_osc(p):        # docs/Demo_CFMR7osc.ipynb
    return lambda t, y: [y[1], -100. * y[0] + 99. * sin(t)]


def detest_b3(p):         # docs/Demo_CFMR7osc.ipynb (problem2)
    return lambda x, y: [-y[0], y[0] - 2 * y[1] ** 2, y[1] ** 2]


def msd(p):               # docs/Demo_own_RK.ipynb
    return lambda t, y: [y[1], 1. - (y[0] + y[1] / 2)]


def expdecay(p):          # overflow / zero-rhs style probes
    return lambda t, y: [p[0] * y[0], p[0] * y[1]]


PROBLEMS = dict(lorenz63=lorenz, vanderpol=vdp, rational=rational,
                duffing=duffing, forced_osc=forced_osc, detest_b3=detest_b3,
                mass_spring_damper=msd, linear=expdecay)


def run_case(case):
    fun = PROBLEMS[case["problem"]](case["params"])
    cls = getattr(ref, case["method"])
    opts = dict(case.get("options", {}))
    if "atol_vec" in case:
        opts["atol"] = np.array(case["atol_vec"])
    t_eval = case.get("t_eval")
    if t_eval is not None:
        t_eval = np.linspace(*t_eval)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = solve_ivp(fun, case["t_span"], case["y0"], method=cls,
                        t_eval=t_eval, **opts)
    nfs = int(ref.NFS)
    out = dict(t=sol.t, y=sol.y, nfev=sol.nfev, status=sol.status, nfs=nfs,
               message=sol.message)
    return out


def run_forced(case):
    """Forced step sequence: open-ended t_bound, the
    tolerances so loose that nothing is rejected, h_abs overwritten before
    every step."""
    fun = PROBLEMS[case["problem"]](case["params"])
    cls = getattr(ref, case["method"])
    hs = np.array(case["h"])
    direction = case.get("direction", 1.0)
    solver = cls(fun, case["t_span"][0], np.array(case["y0"], float),
                 direction * np.inf, rtol=0.1, atol=1e100, first_step=hs[0],
                 nfev_stiff_detect=0)
    ts, ys = [solver.t], [solver.y.copy()]
    for h in hs:
        solver.h_abs = h
        msg = solver.step()
        assert msg is None and int(ref.NFS) == 0
        ts.append(solver.t)
        ys.append(solver.y.copy())
    return dict(t=np.array(ts), y=np.array(ys).T, nfev=solver.nfev, status=0,
                nfs=0, message="")


def main():
    cases = []
    lor = [10.0, 28.0, 8.0 / 3.0]
    tol = dict(rtol=1e-8, atol=1e-10, nfev_stiff_detect=0)
    for m in METHODS:
        cases.append(dict(id=f"lorenz_T10_{m}", method=m, problem="lorenz63",
                          params=lor, y0=[1.0, 1.0, 1.0], t_span=[0.0, 10.0],
                          options=tol))
        # BASELINE.json configs[0] and its siblings;
        # default stiffness detection on, as published there
        cases.append(dict(id=f"lorenz_T100_{m}", method=m,
                          problem="lorenz63", params=lor, y0=[1.0, 1.0, 1.0],
                          t_span=[0.0, 100.0], keep="counts",
                          options=dict(rtol=1e-8, atol=1e-10)))
        cases.append(dict(id=f"lorenz_T100_nostiff_{m}", method=m,
                          problem="lorenz63", params=lor, y0=[1.0, 1.0, 1.0],
                          t_span=[0.0, 100.0], keep="counts", options=tol))
        # reference tests/test_ivp.py:150-213 (forward and backward)
        for span in ([5.0, 9.0], [5.0, 1.0]):
            cases.append(dict(
                id=f"rational_{m}_{span[1]:.0f}", method=m,
                problem="rational", params=[], y0=[1 / 3, 2 / 9],
                t_span=span,
                options=dict(rtol=1e-3, atol=1e-6, nfev_stiff_detect=0)))
        cases.append(dict(
            id=f"rational_teval_{m}", method=m, problem="rational", params=[],
            y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0], t_eval=[5.0, 9.0, 41],
            options=dict(rtol=1e-6, atol=1e-9, nfev_stiff_detect=0)))
        cases.append(dict(
            id=f"rational_teval_back_{m}", method=m, problem="rational",
            params=[], y0=[1 / 3, 2 / 9], t_span=[5.0, 1.0],
            t_eval=[5.0, 1.0, 23],
            options=dict(rtol=1e-6, atol=1e-9, nfev_stiff_detect=0)))
        # first_step / max_step semantics, tests/test_ivp.py:582-665
        cases.append(dict(
            id=f"rational_firststep_{m}", method=m, problem="rational",
            params=[], y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0],
            options=dict(rtol=1e-6, atol=1e-9, first_step=0.1, max_step=0.5,
                         nfev_stiff_detect=0)))
        cases.append(dict(
            id=f"rational_toosmall_{m}", method=m, problem="rational",
            params=[], y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0],
            options=dict(rtol=1e-6, atol=1e-9, max_step=1e-20,
                         nfev_stiff_detect=0)))
        # vector atol
        cases.append(dict(
            id=f"lorenz_atolvec_{m}", method=m, problem="lorenz63",
            params=lor, y0=[-3.0, 2.0, 21.0], t_span=[0.0, 3.0],
            atol_vec=[1e-9, 1e-6, 1e-12],
            options=dict(rtol=1e-7, nfev_stiff_detect=0)))
        # user controller tuple
        cases.append(dict(
            id=f"lorenz_sc_{m}", method=m, problem="lorenz63",
            params=lor, y0=[-3.0, 2.0, 21.0], t_span=[0.0, 3.0],
            options=dict(rtol=1e-6, atol=1e-8, nfev_stiff_detect=0,
                         sc_params=(0.6, -0.25, 0.1, 0.85))))
        # forced step sequence
        k = np.arange(160)
        hs = 0.01 * (1.0 + 0.6 * np.sin(0.37 * k)) + 1e-4
        cases.append(dict(id=f"forced_lorenz_{m}", method=m, forced=True,
                          problem="lorenz63", params=lor, y0=[1.0, 1.0, 1.0],
                          t_span=[0.0, np.inf], h=hs.tolist()))
        cases.append(dict(id=f"forced_vdp_back_{m}", method=m, forced=True,
                          problem="vanderpol", params=[5.0], y0=[2.0, 0.0],
                          t_span=[0.0, -np.inf], direction=-1.0,
                          h=(0.5 * hs[:80]).tolist()))
        # overflow: y' = 1e3 y on a long span with loose tolerance
        cases.append(dict(
            id=f"linear_growth_{m}", method=m, problem="linear",
            params=[40.0], y0=[1.0, -2.0], t_span=[0.0, 1.0],
            options=dict(rtol=1e-5, atol=1e-8, nfev_stiff_detect=0)))

    # C3 samples: Van der Pol mu sweep with t_eval
    for m in ("Pr8", "Pr9", "Pr7", "Ts5"):
        for mu in (0.1, 1.0, 10.0, 100.0):
            cases.append(dict(
                id=f"vdp_mu{mu:g}_{m}", method=m, problem="vanderpol",
                params=[mu], y0=[2.0, 0.0], t_span=[0.0, 20.0],
                t_eval=[0.0, 20.0, 101], options=tol))
    # C2 sample: the first 6 lanes of the seed-12345 Lorenz ensemble, T=10
    rng = np.random.default_rng(12345)
    n_l = 24
    y0s = np.stack([rng.uniform(-15, 15, n_l), rng.uniform(-20, 20, n_l),
                    rng.uniform(5, 40, n_l)], axis=1)
    prm = np.stack([rng.uniform(9, 11, n_l), rng.uniform(24, 32, n_l),
                    rng.uniform(2.4, 2.9, n_l)], axis=1)
    for m in ("Ts5", "CK5"):
        for i in range(6):
            cases.append(dict(
                id=f"c2_lane{i}_{m}", method=m, problem="lorenz63",
                params=prm[i].tolist(), y0=y0s[i].tolist(),
                t_span=[0.0, 10.0], options=tol))
    # BS5 interpolants (bogacki.py:348-393) on Duffing, docs/Demo_BS5.ipynb
    for ip in ("free", "low", "best"):
        cases.append(dict(
            id=f"duffing_BS5_{ip}", method="BS5", problem="duffing",
            params=[], y0=[0.0, 0.0], t_span=[0.0, 20.0],
            t_eval=[0.0, 20.0, 201], options=dict(interpolant=ip)))
        cases.append(dict(
            id=f"duffing_BS5_{ip}_coarse", method="BS5", problem="duffing",
            params=[], y0=[0.0, 0.0], t_span=[0.0, 20.0],
            t_eval=[0.0, 20.0, 201],
            options=dict(interpolant=ip, atol=0.022, first_step=2.05)))
    # notebook known answers (default options => stiffness detection armed)
    cases.append(dict(id="known_duffing_BS5", method="BS5",
                      problem="duffing", params=[], y0=[0.0, 0.0],
                      t_span=[0.0, 20.0], options={}, expect_nfev=212))
    cases.append(dict(id="known_duffing_Ts5", method="Ts5",
                      problem="duffing", params=[], y0=[0.0, 0.0],
                      t_span=[0.0, 20.0], options={}, expect_nfev=341))
    cases.append(dict(id="known_forcedosc_CFMR7osc", method="CFMR7osc",
                      problem="forced_osc", params=[], y0=[1.0, 11.0],
                      t_span=[0.0, 1000.0], keep="counts",
                      options=dict(rtol=1e-5, atol=1e-8),
                      expect_nfev=109091))
    cases.append(dict(id="known_b3_CFMR7osc", method="CFMR7osc",
                      problem="detest_b3", params=[], y0=[1.0, 0.0, 0.0],
                      t_span=[0.0, 20.0],
                      options=dict(rtol=1e-6, atol=1e-9), expect_nfev=275))
    cases.append(dict(id="known_msd_CFMR7osc", method="CFMR7osc",
                      problem="mass_spring_damper", params=[],
                      y0=[0.0, -1.0], t_span=[0.0, 16.0],
                      options=dict(rtol=1e-6), expect_nfev=212,
                      expect_tsize=24))

    arrays = {}
    meta = []
    for c in cases:
        out = run_forced(c) if c.get("forced") else run_case(c)
        if "expect_nfev" in c:
            assert out["nfev"] == c["expect_nfev"], (c["id"], out["nfev"])
        if "expect_tsize" in c:
            assert out["t"].size == c["expect_tsize"], (c["id"], out["t"].size)
        m = dict(c)
        m.update(nfev=int(out["nfev"]), status=int(out["status"]),
                 nfs=int(out["nfs"]), n_t=int(out["t"].size),
                 message=out["message"])
        if c.get("keep") == "counts":
            arrays[c["id"] + "/t_final"] = np.array(out["t"][-1])
            arrays[c["id"] + "/y_final"] = np.array(out["y"][:, -1])
        else:
            arrays[c["id"] + "/t"] = np.asarray(out["t"], float)
            arrays[c["id"] + "/y"] = np.asarray(out["y"], float)
        if "h" in m:
            arrays[c["id"] + "/h"] = np.array(m.pop("h"))
        m["t_span"] = [repr(x) for x in m["t_span"]]
        meta.append(m)
        print(c["id"], m["nfev"], m["nfs"], m["n_t"], m["status"])
    arrays["__meta__"] = np.array(json.dumps(
        dict(cases=meta, numpy=np.__version__, reference=ref.__version__)))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
