#!/usr/bin/env python
"""Dev-time tool: what the UNMODIFIED reference's stiffness diagnosis
(common.py:370-516, 824-1204) reports -- which warning, how many extra RHS
evaluations -- on a fixed list of cases -> tests/golden/stiff_golden.json."""
import json
import os
import sys
import warnings

import numpy as np
from scipy.integrate import solve_ivp

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import extensisq as ref  # noqa: E402
from oracle.problems import make_fun  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "stiff_golden.json")
REAL, COMPLEX, OSC = 1, 2, 4


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


def main():
    cases = []
    for m in ("Ts5", "BS5", "CK5", "Me4", "Pr7", "Pr8", "Pr9", "CFMR7osc"):
        # stiff relaxation oscillator: real dominant root
        cases.append(dict(id=f"vdp_mu100_{m}", method=m, problem="vanderpol",
                          params=[100.0], y0=[2.0, 0.0], t_span=[0.0, 60.0],
                          options=dict(rtol=1e-6, atol=1e-8,
                                       nfev_stiff_detect=1000)))
        # linear stiff decay
        cases.append(dict(id=f"linear_m2000_{m}", method=m, problem="linear",
                          params=[-2000.0], y0=[1.0, -2.0], t_span=[0.0, 3.0],
                          options=dict(rtol=1e-5, atol=1e-8,
                                       nfev_stiff_detect=600)))
        # forced oscillator: complex pair near the imaginary axis
        cases.append(dict(id=f"forced_osc_{m}", method=m, problem="forced_osc",
                          params=[], y0=[1.0, 11.0], t_span=[0.0, 150.0],
                          options=dict(rtol=1e-3, atol=1e-6,
                                       nfev_stiff_detect=400)))
        # non-stiff chaotic: diagnosis runs but stays silent
        cases.append(dict(id=f"lorenz_{m}", method=m, problem="lorenz63",
                          params=[10.0, 28.0, 8 / 3], y0=[1.0, 1.0, 1.0],
                          t_span=[0.0, 15.0],
                          options=dict(rtol=1e-6, atol=1e-8,
                                       nfev_stiff_detect=700)))
    out = []
    for c in cases:
        fun = make_fun(c["problem"], c["params"])
        res = {}
        for key, nsd in (("on", c["options"]["nfev_stiff_detect"]), ("off", 0)):
            opts = dict(c["options"], nfev_stiff_detect=nsd)
            with warnings.catch_warnings(record=True) as ws:
                warnings.simplefilter("always")
                sol = solve_ivp(fun, c["t_span"], c["y0"],
                                method=getattr(ref, c["method"]), **opts)
            res[key] = dict(nfev=int(sol.nfev), nfs=int(ref.NFS),
                            n_t=int(sol.t.size), status=int(sol.status),
                            flags=flags_of(ws),
                            y_final=[float(v).hex() for v in sol.y[:, -1]])
        assert res["on"]["n_t"] == res["off"]["n_t"]       # never changes t, y, h
        assert res["on"]["y_final"] == res["off"]["y_final"]
        c.update(nfev=res["on"]["nfev"], nfev_off=res["off"]["nfev"],
                 nfs=res["on"]["nfs"], n_t=res["on"]["n_t"],
                 status=res["on"]["status"], flags=res["on"]["flags"],
                 y_final=res["on"]["y_final"])
        out.append(c)
        print(c["id"], c["nfev"], c["nfev_off"], c["nfs"], c["n_t"], "flags", c["flags"])
    with open(OUT, "w") as fh:
        json.dump(dict(cases=out, reference=ref.__version__), fh, indent=0)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
