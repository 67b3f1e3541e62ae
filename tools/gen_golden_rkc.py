#!/usr/bin/env python
"""Dev-time tool: golden vectors of the UNMODIFIED reference's SSV2stab solver
(extensisq/sommeijer.py) -> tests/golden/rkc_golden.npz.  /root/reference is
only read here.  The reference has no test for SSV2stab; besides these vectors
the notebook table docs/Demo_SSV2stab.ipynb:350-356 is asserted below."""
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
from extensisq import sommeijer  # noqa: E402
from oracle.problems import heat2d_reaction, heat3d_notebook  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rkc_golden.npz")


def run(fun, span, y0, rho, t_eval=None, **opts):
    if rho is not None:
        opts["rho_jac"] = lambda t, y: float(rho)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = solve_ivp(fun, span, y0, method=ref.SSV2stab, t_eval=t_eval,
                        **opts)
    return sol, int(sommeijer.NFS), int(sommeijer.nfesig), int(sommeijer.maxm)


def main():
    arrays, meta = {}, []
    cases = []
    for nx in (32, 64, 128):
        cases.append(dict(id=f"rd2d_{nx}_rho", nx=nx, use_rho=True,
                          t_span=[0.0, 0.05], options=dict(rtol=1e-4, atol=1e-4)))
    cases.append(dict(id="rd2d_32_power", nx=32, use_rho=False,
                      t_span=[0.0, 0.05], options=dict(rtol=1e-4, atol=1e-4)))
    cases.append(dict(id="rd2d_32_power_constjac", nx=32, use_rho=False,
                      t_span=[0.0, 0.05],
                      options=dict(rtol=1e-5, atol=1e-5, const_jac=True)))
    cases.append(dict(id="rd2d_48_rho_constjac_tight", nx=48, use_rho=True,
                      t_span=[0.0, 0.1],
                      options=dict(rtol=1e-6, atol=1e-7, const_jac=True)))
    cases.append(dict(id="rd2d_32_teval", nx=32, use_rho=True,
                      t_span=[0.0, 0.05], t_eval=[0.0, 0.05, 11],
                      options=dict(rtol=1e-4, atol=1e-4)))
    cases.append(dict(id="rd2d_32_firststep_maxstep", nx=32, use_rho=True,
                      t_span=[0.0, 0.05],
                      options=dict(rtol=1e-4, atol=1e-4, first_step=1e-4,
                                   max_step=4e-3)))
    cases.append(dict(id="rd2d_24_loose", nx=24, use_rho=True,
                      t_span=[0.0, 0.3], options=dict(rtol=1e-2, atol=1e-2)))
    for c in cases:
        fun, y0, rho = heat2d_reaction(c["nx"])
        te = np.linspace(*c["t_eval"]) if c.get("t_eval") else None
        sol, nfs, nsig, smax = run(fun, c["t_span"], y0,
                                   rho if c["use_rho"] else None, te,
                                   **c["options"])
        m = dict(c, nfev=int(sol.nfev), nfs=nfs, nfesig=nsig, maxm=smax,
                 n_t=int(sol.t.size), status=int(sol.status), rho=float(rho))
        arrays[c["id"] + "/t"] = sol.t
        arrays[c["id"] + "/y"] = sol.y if te is not None else sol.y[:, -1]
        meta.append(m)
        print(c["id"], m["nfev"], nfs, nsig, smax, m["n_t"])
    # notebook table (first four tolerances): steps(rej) / f-evals / s-max
    fun, y0, rho = heat3d_notebook()
    table = {1e-1: (6, 1, 402, 132), 1e-2: (15, 4, 729, 85),
             1e-3: (27, 2, 786, 40), 1e-4: (57, 0, 1087, 26)}
    for tol, exp in table.items():
        sol, nfs, nsig, smax = run(fun, (0, 0.7), y0, rho, rtol=tol, atol=tol,
                                   const_jac=True)
        got = (sol.t.size - 1 + nfs, nfs, sol.nfev, smax)
        assert got == exp, (tol, got, exp)
        cid = f"heat3d_tol{tol:g}"
        meta.append(dict(id=cid, notebook=True, tol=tol, nfev=int(sol.nfev),
                         nfs=nfs, nfesig=nsig, maxm=smax, n_t=int(sol.t.size),
                         status=int(sol.status), rho=float(rho)))
        arrays[cid + "/y_sample"] = sol.y[::97, -1]     # decimated, keeps the file small
        print(cid, got)
    arrays["__meta__"] = np.array(json.dumps(
        dict(cases=meta, numpy=np.__version__, reference=ref.__version__)))
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
