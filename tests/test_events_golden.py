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


# SWAG cases pin the device path directly (there is no NumPy SWAG restatement)
RK_CASES = [c for c in CASES if c["method"] != "SWAG"]


@pytest.mark.parametrize("c", RK_CASES, ids=lambda c: c["id"])
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


# ---- the reference's own event tests (tests/test_ivp.py:369-470, 757-783),
# run on the restated solvers + restated scipy machinery --------------------
ALL = ["BS5", "Ts5", "CK5", "CKdisc", "Pr7", "Pr8", "Pr9", "CFMR7osc", "Me4"]


def sol_rational(t):                       # tests/test_ivp.py:31-32
    return np.asarray((t / (t + 10), 10 * t / (t + 10) ** 2))


def run_rational(method, span, y0, term, direc, solver, **kw):
    fns = EVENT_SETS["rational"][0]
    return solver(method, span, y0, [(g, a, b) for g, a, b in zip(fns, term, direc)], **kw)


def oracle_solver(method, span, y0, events, **kw):
    r = RO.rk_solve(TABS[method], make_fun("rational", []), span, y0, events=events, **kw)
    return dict(status=r["status"], t_events=r["t_events"],
                y_events=[np.asarray(y) for y in r["y_events"]], t=r["t"], y=r["y"])


def check_reference_event_test(method, solver):
    """Assertions of the reference's test_events, for any solver callable."""
    e1, e2, e3 = EVENT_SETS["rational"][0]
    y0 = [1 / 3, 2 / 9]
    # (the reference passes only the first two events in its first calls; here
    # the third, `t - 7.4`, rides along as a non-terminal event)
    res = run_rational(method, [5, 8], y0, [0, 0, 0], [0, 0, 1], solver)
    assert res["status"] == 0
    assert res["t_events"][0].size == 1 and res["t_events"][1].size == 1
    assert 5.3 < res["t_events"][0][0] < 5.7
    assert 7.3 < res["t_events"][1][0] < 7.7
    assert res["y_events"][0].shape == (1, 2) and res["y_events"][1].shape == (1, 2)
    assert np.isclose(e1(res["t_events"][0][0], res["y_events"][0][0]), 0)
    assert np.isclose(e2(res["t_events"][1][0], res["y_events"][1][0]), 0)
    res = run_rational(method, [5, 8], y0, [0, 0, 0], [1, 1, 1], solver)
    assert res["status"] == 0
    assert res["t_events"][0].size == 1 and res["t_events"][1].size == 0
    assert 5.3 < res["t_events"][0][0] < 5.7
    res = run_rational(method, [5, 8], y0, [0, 0, 0], [-1, -1, 1], solver)
    assert res["status"] == 0
    assert res["t_events"][0].size == 0 and res["t_events"][1].size == 1
    assert 7.3 < res["t_events"][1][0] < 7.7
    res = run_rational(method, [5, 8], y0, [0, 0, 1], [0, 0, 0], solver)
    assert res["status"] == 1
    assert res["t_events"][0].size == 1 and res["t_events"][1].size == 0
    assert res["t_events"][2].size == 1
    assert 5.3 < res["t_events"][0][0] < 5.7 and 7.3 < res["t_events"][2][0] < 7.5
    assert np.isclose(e3(res["t_events"][2][0], res["y_events"][2][0]), 0)
    assert np.allclose(sol_rational(res["t_events"][0][0]), res["y_events"][0][0],
                       rtol=1e-3, atol=1e-6)
    # backward direction
    res = run_rational(method, [8, 5], [4 / 9, 20 / 81], [0, 0, 0], [0, 0, 0], solver)
    assert res["status"] == 0
    assert res["t_events"][0].size == 1 and res["t_events"][1].size == 1
    assert 5.3 < res["t_events"][0][0] < 5.7 and 7.3 < res["t_events"][1][0] < 7.7
    res = run_rational(method, [8, 5], [4 / 9, 20 / 81], [0, 0, 1], [0, 0, 0], solver)
    assert res["status"] == 1
    assert res["t_events"][0].size == 0 and res["t_events"][2].size == 1
    assert 7.3 < res["t_events"][2][0] < 7.5


@pytest.mark.parametrize("method", ALL)
def test_reference_event_test_on_the_oracle(method):
    check_reference_event_test(method, oracle_solver)


@pytest.mark.parametrize("method", ALL)
def test_reference_t_eval_early_event_on_the_oracle(method):
    # tests/test_ivp.py:757-783: terminal event before the first t_eval point
    te = np.linspace(7.5, 9, 16)
    r = RO.rk_solve(TABS[method], make_fun("rational", []), [5, 9], [1 / 3, 2 / 9],
                    t_eval=te, events=[(EVENT_SETS["early"][0][0], 1, 0)])
    assert r["status"] == 1
    assert r["t"].size == 0 and r["y"].size == 0
    assert r["t_events"][0].size == 1 and r["t_events"][0][0] == 7


@pytest.mark.parametrize("c", RK_CASES, ids=lambda c: c["id"])
def test_c_oracle_events_in_device_arithmetic_against_the_reference(c):
    """events_after_step / brentq_c of oracle/xsq_oracle.c (the C restatement of
    the kernels' event handling, bit-identical to the kernel sources:
    tests/test_kernel_host.py) against the reference's golden runs: same status,
    and -- where the kernels' arithmetic takes the same steps -- event times to
    1e-9, event states to 1e-7, the t_eval output cut at the same terminal event."""
    from oracle import c_oracle as CO
    opts = ev_options(c)
    fns, _ = EVENT_SETS[c["events"]]
    te = ev_t_eval(c)
    with CO.device_math():
        o = CO.rk_events_batch(TABS[c["method"]], None, c["t_span"], [c["y0"]], list(fns),
                               c["terminal"], c["direction"], 64, t_eval=te,
                               user_fn=make_fun(c["problem"], c["params"]), **opts)
    assert int(o["status"][0]) == c["status"]
    if int(o["nfev"][0]) != c["nfev"] or int(o["n_rejected"][0]) != c["nfs"]:
        pytest.skip("the kernels' arithmetic takes a different step sequence on this case")
    for k in range(len(c["terminal"])):
        tg = unhex(c["t_events"][k])
        assert int(o["event_counts"][0, k]) >= tg.size
        assert np.allclose(o["t_events"][0, k, :tg.size], tg, rtol=1e-9, atol=1e-9)
        assert np.isnan(o["t_events"][0, k, tg.size:]).all()
        if tg.size:
            ye = unhex(c["y_events"][k]).reshape(tg.size, -1)
            assert np.allclose(o["y_events"][0, k, :tg.size], ye, rtol=1e-7, atol=1e-7)
    if te is not None:
        yg = unhex(c["y"])
        n_done = int(o["n_eval_done"][0])
        assert n_done == (yg.shape[1] if yg.ndim == 2 else 0)
        if n_done:
            assert np.allclose(o["y"][0][:, :n_done], yg, rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("c", [c for c in CASES if c["method"] == "SWAG"], ids=lambda c: c["id"])
def test_c_swag_oracle_events_against_the_reference(c):
    """The same for SWAG (oracle/xsq_oracle_swag.c with events, device arithmetic)
    against the reference's golden runs of SWAG driven by scipy's event loop."""
    from oracle import c_oracle as CO
    if c["problem"] not in CO.RHS_IDS:
        pytest.skip("the SWAG C oracle takes built-in right-hand sides")
    opts = ev_options(c)
    fns, _ = EVENT_SETS[c["events"]]
    te = ev_t_eval(c)
    with CO.device_math():
        o = CO.swag_events_batch(c["problem"], c["t_span"], [c["y0"]], list(fns), c["terminal"],
                                 c["direction"], 64, params=[c["params"]] if c["params"] else None,
                                 t_eval=te, **opts)
    assert int(o["status"][0]) == c["status"]
    if int(o["nfev"][0]) != c["nfev"]:
        pytest.skip("the kernels' arithmetic takes a different step sequence on this case")
    for k in range(len(c["terminal"])):
        tg = unhex(c["t_events"][k])
        assert int(o["event_counts"][0, k]) >= tg.size
        assert np.allclose(o["t_events"][0, k, :tg.size], tg, rtol=1e-9, atol=1e-9)
        if tg.size:
            ye = unhex(c["y_events"][k]).reshape(tg.size, -1)
            assert np.allclose(o["y_events"][0, k, :tg.size], ye, rtol=1e-7, atol=1e-7)
