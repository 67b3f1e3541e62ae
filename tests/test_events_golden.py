"""Event detection: scipy's solve_ivp event
machinery restated in oracle/rk_oracle.py (find_active_events, handle_events,
brentq) on top of the restated solvers, against golden vectors of the
unmodified reference driven by scipy (tools/gen_golden_events.py)."""
import json
import os

import numpy as np
import pytest

from oracle import rk_oracle as RO
from oracle.problems import make_fun, EVENT_SETS

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "events_golden.json")) as fh:
    CASES = json.load(fh)["cases"]
TABS = RO.load_tableaux()
TABS["CKdisc"] = RO.load_ckdisc()


def unhex(a):
    if not a:
        return np.zeros(0)
    return np.array([[float.fromhex(v) for v in row] for row in a]
                    if isinstance(a[0], list) else [float.fromhex(v) for v in a])


def ev_options(c):
    o = dict(c["options"])
    if isinstance(o.get("atol"), list):
        o["atol"] = np.array(o["atol"])
    return o


def ev_t_eval(c):
    return np.linspace(*c["t_eval"][:2], int(c["t_eval"][2])) if c["t_eval"] else None


def ev_list(c):
    fns, _ = EVENT_SETS[c["events"]]
    return [(g, tr, d) for g, tr, d in zip(fns, c["terminal"], c["direction"])]


def test_brentq_restatement_matches_scipy():
    sp = pytest.importorskip("scipy.optimize")
    rng = np.random.default_rng(0)
    eps = np.finfo(float).eps
    for k in range(600):
        a, b, c, d = rng.normal(size=4)
        f = [lambda x: (x - a) * (1 + 0.3 * np.sin(b * x)),
             lambda x: np.tanh(3 * (x - a)) + 0.1 * (x - a) ** 3,
             lambda x: np.exp(c * (x - a)) - 1,
             lambda x: (x - a) ** 3 * (1 + 0.1 * d) + 1e-3 * (x - a)][k % 4]
        lo, hi = a - abs(b) - 0.1, a + abs(c) + 0.2
        if k % 7 == 0:
            lo, hi = hi, lo
        r0, info = sp.brentq(f, lo, hi, xtol=4 * eps, rtol=4 * eps, full_output=True)
        r1, calls = RO.brentq(f, lo, hi)
        assert r0 == r1 and calls == info.function_calls


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_oracle_events_bit_identical_to_reference(c):
    r = RO.rk_solve(TABS[c["method"]], make_fun(c["problem"], c["params"]), c["t_span"],
                    c["y0"], t_eval=ev_t_eval(c), events=ev_list(c), **ev_options(c))
    assert r["status"] == c["status"]
    assert r["nfev"] == c["nfev"]
    assert r["n_rejected"] == c["nfs"]
    assert np.array_equal(r["t"], unhex(c["t"]))
    assert np.array_equal(r["y"], unhex(c["y"]).reshape(r["y"].shape))
    for k in range(len(c["terminal"])):
        assert np.array_equal(r["t_events"][k], unhex(c["t_events"][k]))
        ye = unhex(c["y_events"][k])
        assert np.array_equal(np.asarray(r["y_events"][k]).reshape(ye.shape), ye)
