#!/usr/bin/env python
"""Dev-time tool: golden vectors of the UNMODIFIED reference's SWAG solver
(extensisq/shampine.py) -> tests/golden/swag_golden.npz.  Same conventions as
tools/gen_golden.py; /root/reference is only read here."""
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

OUT = os.path.join(ROOT, "tests", "golden", "swag_golden.npz")
AREN_Y0 = [0.994, 0.0, 0.0, -2.00158510637908252240537862224]
AREN_T = 17.0652165601579625588917206249
AREN_MU = 0.012277471


def main():
    tol = dict(rtol=1e-8, atol=1e-10)
    cases = [
        dict(id="arenstorf_period", problem="arenstorf", params=[AREN_MU],
             y0=AREN_Y0, t_span=[0.0, AREN_T], options=tol,
             expect=(593, 16, 1207)),
        dict(id="arenstorf_teval", problem="arenstorf", params=[AREN_MU],
             y0=AREN_Y0, t_span=[0.0, AREN_T], t_eval=[0.0, AREN_T, 201],
             options=tol),
        dict(id="lorenz_T10", problem="lorenz63", params=[10.0, 28.0, 8 / 3],
             y0=[1.0, 1.0, 1.0], t_span=[0.0, 10.0], options=tol),
        dict(id="lorenz_T10_kmax5", problem="lorenz63",
             params=[10.0, 28.0, 8 / 3], y0=[1.0, 1.0, 1.0],
             t_span=[0.0, 10.0], options=dict(rtol=1e-6, atol=1e-8, k_max=5)),
        dict(id="lorenz_loose_teval", problem="lorenz63",
             params=[10.0, 28.0, 8 / 3], y0=[-3.0, 2.0, 21.0],
             t_span=[0.0, 4.0], t_eval=[0.0, 4.0, 81],
             options=dict(rtol=1e-4, atol=1e-6)),
        dict(id="rational_fwd", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0],
             options=dict(rtol=1e-3, atol=1e-6)),
        dict(id="rational_back", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 1.0],
             options=dict(rtol=1e-3, atol=1e-6)),
        dict(id="rational_teval", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0], t_eval=[5.0, 9.0, 41],
             options=dict(rtol=1e-6, atol=1e-9)),
        dict(id="rational_teval_back", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 1.0], t_eval=[5.0, 1.0, 23],
             options=dict(rtol=1e-6, atol=1e-9)),
        dict(id="rational_firststep", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0],
             options=dict(rtol=1e-6, atol=1e-9, first_step=0.1, max_step=0.5)),
        dict(id="rational_toosmall", problem="rational", params=[],
             y0=[1 / 3, 2 / 9], t_span=[5.0, 9.0],
             options=dict(rtol=1e-6, atol=1e-9, max_step=1e-20)),
        dict(id="duffing", problem="duffing", params=[], y0=[0.0, 0.0],
             t_span=[0.0, 20.0], t_eval=[0.0, 20.0, 201], options={}),
        dict(id="lorenz_atolvec", problem="lorenz63",
             params=[10.0, 28.0, 8 / 3], y0=[-3.0, 2.0, 21.0],
             t_span=[0.0, 3.0], atol_vec=[1e-9, 1e-6, 1e-12],
             options=dict(rtol=1e-7)),
    ]
    for mu in (0.1, 1.0, 10.0):
        cases.append(dict(id=f"vdp_mu{mu:g}", problem="vanderpol", params=[mu],
                          y0=[2.0, 0.0], t_span=[0.0, 20.0],
                          t_eval=[0.0, 20.0, 101], options=tol))
    # C2-style random lanes
    rng = np.random.default_rng(12345)
    n_l = 24
    y0s = np.stack([rng.uniform(-15, 15, n_l), rng.uniform(-20, 20, n_l),
                    rng.uniform(5, 40, n_l)], axis=1)
    prm = np.stack([rng.uniform(9, 11, n_l), rng.uniform(24, 32, n_l),
                    rng.uniform(2.4, 2.9, n_l)], axis=1)
    for i in range(4):
        cases.append(dict(id=f"c2_lane{i}", problem="lorenz63",
                          params=prm[i].tolist(), y0=y0s[i].tolist(),
                          t_span=[0.0, 5.0], options=tol))
    arrays, meta = {}, []
    for c in cases:
        fun = make_fun(c["problem"], c["params"])
        opts = dict(c["options"])
        if "atol_vec" in c:
            opts["atol"] = np.array(c["atol_vec"])
        te = np.linspace(*c["t_eval"]) if c.get("t_eval") else None
        ref.NFS[()] = 0      # SWAG.__init__ does not reset the global counter
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol = solve_ivp(fun, c["t_span"], c["y0"], method=ref.SWAG,
                            t_eval=te, **opts)
        nfs = int(ref.NFS)
        if "expect" in c:
            assert (sol.t.size - 1, nfs, sol.nfev) == c["expect"], \
                (sol.t.size - 1, nfs, sol.nfev)
        m = dict(c)
        m.pop("expect", None)
        m.update(nfev=int(sol.nfev), status=int(sol.status), nfs=nfs,
                 n_t=int(sol.t.size), message=sol.message)
        arrays[c["id"] + "/t"] = np.asarray(sol.t, float)
        arrays[c["id"] + "/y"] = np.asarray(sol.y, float)
        meta.append(m)
        print(c["id"], m["nfev"], nfs, m["n_t"], m["status"])
    arrays["__meta__"] = np.array(json.dumps(
        dict(cases=meta, numpy=np.__version__, reference=ref.__version__)))
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
