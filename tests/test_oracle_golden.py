"""The oracle is pinned against the reference: golden vectors produced by the
UNMODIFIED reference (tools/gen_golden.py, run in the build container where
/root/reference is importable) vs

* oracle/rk_oracle.py  -- NumPy restatement: same step counts and states equal
  to 1e-13 relative (bit-identical under the same NumPy/OpenBLAS; the
  tolerance only allows for a different BLAS build on another host);
* oracle/xsq_oracle.c  -- plain-C restatement with fma accumulation: same
  accepted / rejected / nfev counts and states within 1e-9 relative on the
  accuracy-limited cases, approximate counts on the stability-limited ones.
"""
import numpy as np
import pytest

from golden_util import (Golden, BUILTIN_PROBLEMS, case_options, case_span,
                         case_t_eval, stability_limited)
from oracle import rk_oracle as O
from oracle import c_oracle as CO
from oracle.problems import make_fun

G = Golden()
TABS = O.load_tableaux()
FAST = [c for c in G.cases if c.get("keep") != "counts"]
SLOW_IDS = ["lorenz_T100_BS5", "lorenz_T100_nostiff_BS5", "lorenz_T100_Ts5",
            "lorenz_T100_nostiff_CK5", "lorenz_T100_nostiff_Pr9"]


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("c", FAST, ids=[c["id"] for c in FAST])
def test_numpy_oracle_matches_reference(c):
    tab = TABS[c["method"]]
    fun = make_fun(c["problem"], c["params"])
    if c.get("forced"):
        r = O.rk_solve(tab, fun, case_span(c), c["y0"],
                       forced_h=G.arr(c["id"], "h"))
    else:
        r = O.rk_solve(tab, fun, case_span(c), c["y0"],
                       t_eval=case_t_eval(c), **case_options(c))
    assert r["n_rejected"] == c["nfs"]
    assert r["status"] == c["status"]
    assert r["nfev"] == c["nfev"]
    t, y = G.arr(c["id"], "t"), G.arr(c["id"], "y")
    assert r["t"].shape == t.shape and r["y"].shape == y.shape
    if t.size:
        assert _rel(r["t"], t) <= 1e-15
        assert _rel(r["y"], y) <= 1e-13
    if c["status"] == -1:
        assert r["message"] == c["message"]


@pytest.mark.parametrize("cid", SLOW_IDS)
def test_numpy_oracle_lorenz_T100_counts(cid):
    """BASELINE.json configs[0]: BS5, Lorenz-63, t in [0,100], rtol 1e-8, atol
    1e-10 -> 6822 accepted / 261 rejected."""
    c = G.by_id[cid]
    r = O.rk_solve(TABS[c["method"]], make_fun(c["problem"], c["params"]),
                   case_span(c), c["y0"], **case_options(c))
    assert r["n_rejected"] == c["nfs"]
    assert r["n_accepted"] == c["n_t"] - 1
    assert r["nfev"] == c["nfev"]      # with and without stiffness diagnosis
    assert _rel(r["y_final"], G.arr(cid, "y_final")) <= 1e-13
    if cid == "lorenz_T100_BS5":
        assert (r["n_accepted"], r["n_rejected"]) == (6822, 261)
        assert c["nfev"] == 49358


@pytest.mark.parametrize("c", FAST, ids=[c["id"] for c in FAST])
def test_c_oracle_matches_reference(c):
    tab = TABS[c["method"]]
    span = case_span(c)
    if c["problem"] in BUILTIN_PROBLEMS:
        kw = dict(rhs=c["problem"], params=[c["params"]])
    else:
        kw = dict(rhs=None, user_fn=make_fun(c["problem"], c["params"]))
    if c.get("forced"):
        h = G.arr(c["id"], "h")
        r = CO.rk_batch(tab, t_span=[span[0], span[0] + np.sign(span[1])],
                        y0=c["y0"], forced_h=h, **kw)
        assert r["n_accepted"][0] == h.size and r["nfev"][0] == c["nfev"]
        # "with a forced fixed step sequence, states agree to 1e-12 relative"
        assert _rel(r["y_final"][0], G.arr(c["id"], "y")[:, -1]) <= 1e-12
        assert _rel(r["t_final"], G.arr(c["id"], "t")[-1:]) <= 1e-15
        return
    te = case_t_eval(c)
    r = CO.rk_batch(tab, t_span=span, y0=c["y0"], t_eval=te, **kw,
                    **case_options(c))
    assert (r["status"][0] == 0) == (c["status"] == 0)
    if c["status"] == -1:
        assert r["status"][0] == -1          # TOO_SMALL_STEP
    n_t = c["n_t"]
    if stability_limited(c):
        assert abs(r["n_rejected"][0] - c["nfs"]) <= max(3, 0.15 * c["nfs"])
        assert abs(r["nfev"][0] - c["nfev"]) <= 0.05 * c["nfev"]
        tol = 1e-6
    else:
        assert r["n_rejected"][0] == c["nfs"]
        assert r["nfev"][0] == c["nfev"]
        if te is None:
            assert r["n_accepted"][0] == n_t - 1
        tol = 1e-9
    if te is None:
        if n_t > 1 or c["status"] == 0:
            assert _rel(r["y_final"][0], G.arr(c["id"], "y")[:, -1]) <= tol
    else:
        yg = G.arr(c["id"], "y")
        assert r["n_eval_done"][0] == yg.shape[1]
        assert _rel(r["y"][0][:, :yg.shape[1]], yg) <= tol
