#!/usr/bin/env python
"""Dev-time tool: golden vectors for the Runge-Kutta-Nystrom methods
(second half).

Runs the UNMODIFIED reference (``extensisq.Fi4N / Fi5N / Mu5Nmb / MR6NN``
through scipy's ``solve_ivp``, imported from /root/reference, this container
only) on second order problems in first order form [v, a] = fun(t, [x, v]) and
stores every accepted (t, y), nfev, the status and the stiffness warnings,
losslessly (hex floats), in ``tests/golden/rkn_golden.json``.  The right-hand
sides are ``oracle/problems_rkn.py``.

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_rkn.py
"""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import extensisq as ref                      # noqa: E402
from scipy.integrate import solve_ivp        # noqa: E402
from oracle.problems_rkn import PROBLEMS, make_fun   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rkn_golden.json")
REAL, COMPLEX, OSC = 1, 2, 4


def hx(a):
    a = np.asarray(a, dtype=float)
    if a.ndim == 2:
        return [[float(v).hex() for v in row] for row in a]
    return [float(v).hex() for v in a]


def flags_of(ws):
    f = 0
    for w in ws:
        m = str(w.message)
        if "real dominant root" in m and "diagnosed as stiff" in m:
            f |= REAL
        elif "complex pair of dominant roots" in m and "diagnosed as stiff" in m:
            f |= COMPLEX
        elif "near the imaginary axis" in m:
            f |= OSC
    return f


# id suffix, problem, params, t_span, options
BASE = [
    ("osc_default", "oscillator", [1.0], [0., 10.], {}),
    ("osc_tight_back", "oscillator", [2.5], [3., -4.], dict(rtol=1e-9, atol=1e-11)),
    ("kepler", "kepler", [0.5], [0., 12.], dict(rtol=1e-7, atol=1e-9)),
    ("kepler_atolvec", "kepler", [0.3], [0., 7.], dict(rtol=1e-5, atol=[1e-7, 1e-6, 1e-8, 1e-7])),
    ("kepler_steps", "kepler", [0.6], [0., 6.], dict(rtol=1e-6, atol=1e-8, first_step=1e-3, max_step=0.25)),
    ("nbody32", "nbody32", [], [0., 0.5], dict(rtol=1e-6, atol=1e-8)),
]
VELDEP = [
    ("vdp_mu2", "vanderpol", [2.0], [0., 10.], dict(rtol=1e-6, atol=1e-8)),
    ("vdp_mu30", "vanderpol", [30.0], [0., 40.], dict(rtol=1e-4, atol=1e-6)),
    ("arenstorf", "arenstorf", [0.012277471], [0., 4.], dict(rtol=1e-7, atol=1e-9)),
    ("damped_stiff", "damped", [400.0, 800.0], [0., 60.], dict(rtol=1e-4, atol=1e-6)),
]


def main():
    cases = []
    # Mu5Nmb scales its embedded weights IN PLACE on the class attribute
    # (`self.E *= factor`, murua.py:224-227), so in the reference every further
    # instantiation in the same process multiplies them by 0.75 again.  The
    # golden vectors are those of a fresh process: the class attributes are
    # restored before each run.
    E0, Ep0 = ref.Mu5Nmb.E.copy(), ref.Mu5Nmb.Ep.copy()
    for name in ("Fi4N", "Fi5N", "Mu5Nmb", "MR6NN"):
        cls = getattr(ref, name)
        todo = BASE + ([] if name == "MR6NN" else VELDEP)
        for cid, prob, prm, span, opt in todo:
            fun = make_fun(prob, prm)
            y0 = PROBLEMS[prob]["y0"](prm)
            ref.Mu5Nmb.E[:], ref.Mu5Nmb.Ep[:] = E0, Ep0
            with warnings.catch_warnings(record=True) as ws:
                warnings.simplefilter("always")
                sol = solve_ivp(fun, span, y0, method=cls, **opt)
            keep = slice(None) if sol.t.size <= 1000 else slice(-1, None)   # long runs: end only
            c = dict(id=f"{name}_{cid}", method=name, problem=prob, params=prm,
                     y0=hx(y0), t_span=span, options=opt, n_t=int(sol.t.size),
                     t=hx(sol.t[keep]), y=hx(sol.y[:, keep]),
                     nfev=int(sol.nfev), status=int(sol.status), flags=flags_of(ws))
            cases.append(c)
            print(c["id"], "steps", sol.t.size - 1, "nfev", sol.nfev, "status", sol.status,
                  "flags", c["flags"])
    with open(OUT, "w") as fh:
        json.dump({"reference_version": ref.__version__, "cases": cases}, fh)
    print("wrote", OUT, len(cases), "cases", os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
