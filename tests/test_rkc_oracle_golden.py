"""The C restatement of SSV2stab (oracle/xsq_oracle_rkc.c) is pinned against
(a) golden vectors of the unmodified reference on 2-D reaction-diffusion grids
(tools/gen_golden_rkc.py) and (b) the reference's published notebook table
docs/Demo_SSV2stab.ipynb:350-356 (3-D heat problem: steps, rejected steps,
f-evals and s-max per tolerance).  The reference's own test-suite has no
SSV2stab test, so these are the only pins there are."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle.problems import heat2d_reaction, heat3d_notebook

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "rkc_golden.npz"))
CASES = json.loads(str(Z["__meta__"]))["cases"]
NOTEBOOK = {1e-1: (6, 1, 402, 132), 1e-2: (15, 4, 729, 85),
            1e-3: (27, 2, 786, 40), 1e-4: (57, 0, 1087, 26)}


def check_counts(c, r):
    assert r["status"] == 0
    assert (r["nfev"], r["n_rejected"], r["nfesig"], r["maxm"]) == \
        (c["nfev"], c["nfs"], c["nfesig"], c["maxm"])


@pytest.mark.parametrize("c", [c for c in CASES if not c.get("notebook")],
                         ids=lambda c: c["id"])
def test_c_oracle_rkc_matches_reference(c):
    fun, y0, rho = heat2d_reaction(c["nx"])
    te = np.linspace(*c["t_eval"]) if c.get("t_eval") else None
    r = CO.rkc_solve(y0, c["t_span"], rho=rho if c["use_rho"] else None,
                     t_eval=te, **c["options"])
    check_counts(c, r)
    yg = Z[c["id"] + "/y"]
    got = r["y"] if te is not None else r["y_final"]
    if te is None:
        assert r["n_accepted"] == c["n_t"] - 1
    assert np.abs(got - yg).max() <= 1e-11
    # the built-in C right-hand side and the Python one are the same function
    r2 = CO.rkc_solve(y0, c["t_span"], rho=rho if c["use_rho"] else None,
                      fun=fun, **c["options"])
    assert np.abs(r2["y_final"] - r["y_final"]).max() <= 1e-12


@pytest.mark.parametrize("c", [c for c in CASES if c.get("notebook")],
                         ids=lambda c: c["id"])
def test_c_oracle_rkc_reproduces_notebook_table(c):
    fun, y0, rho = heat3d_notebook()
    r = CO.rkc_solve(y0, (0, 0.7), rtol=c["tol"], atol=c["tol"],
                     const_jac=True, rho=rho, fun=fun)
    check_counts(c, r)
    steps, rej, nfev, smax = NOTEBOOK[c["tol"]]
    assert (r["n_accepted"] + r["n_rejected"], r["n_rejected"], r["nfev"],
            r["maxm"]) == (steps, rej, nfev, smax)
    assert np.abs(r["y_final"][::97] - Z[c["id"] + "/y_sample"]).max() <= 1e-11
