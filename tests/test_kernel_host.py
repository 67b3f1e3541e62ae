"""The CUDA kernel SOURCES (extensisq_b200/csrc: ens_init_body, rk_fast_body,
rk_persistent_body, stiff_queue_body) compiled for the host by g++
(tests/kernel_host/) and run over small ensembles, against the C oracle in
device arithmetic: every output must be bit identical.  Checks the real kernel
logic -- controller, _reassess_stepsize short cuts, stiffness bookkeeping, probe
queue and slot paths, dense output -- without a GPU; the GPU tests
(tests/test_gpu_exact.py, tests/test_gpu_fast.py) repeat it on the device."""
import os
import sys

import numpy as np
import pytest

import extensisq_b200 as xb
from oracle import c_oracle as CO
from oracle import rk_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_host"))
import emu  # noqa: E402

TABS = O.load_tableaux()
METHODS = [xb.Ts5, xb.BS5, xb.CK5, xb.Me4, xb.Pr7, xb.Pr8, xb.Pr9, xb.CFMR7osc]
GENERIC = [xb.Ts5, xb.CK5, xb.Me4, xb.Pr7, xb.Pr8, xb.Pr9]
KEYS_F = ("t_final", "y_final", "h_next")
KEYS_I = ("n_accepted", "n_rejected", "nfev", "status", "stiff_flags")


def lanes(prob, N, seed=12345):
    rng = np.random.default_rng(seed)
    if prob == "lorenz63":
        y0 = np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 40, N)], 1)
        prm = np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1)
        return y0, prm, (0.0, 6.0)
    if prob == "vanderpol":
        mu = 10.0 ** (-1 + 3 * np.arange(N) / max(N - 1, 1))
        return np.tile([2.0, 0.0], (N, 1)), mu[:, None], (0.0, 6.0)
    y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) + rng.uniform(-1e-3, 1e-3, (N, 4))
    return y0, np.full((N, 1), 0.012277471), (0.0, 17.0652165601579625588917206249)


def oracle(prob, span, y0, m, prm, **kw):
    kw.pop("fast", None)
    kw.pop("queue_records", None)
    with CO.device_math():
        return CO.rk_batch(TABS[m.__name__], prob, span, y0, params=prm, n_threads=CO.max_threads(), **kw)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def same(g, o, what, dense=False):
    for k in KEYS_I:
        assert np.array_equal(g[k], o[k]), (what, k, np.flatnonzero(g[k] != o[k])[:5])
    for k in KEYS_F:
        assert np.array_equal(bits(g[k]), bits(o[k])), (what, k)
    if dense:
        a, b = g["y"], o["y"]
        ok = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
        assert ok.all(), (what, "y(t_eval)", np.argwhere(~ok)[:3])


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_kernel_source_equals_oracle_bit_for_bit(m, prob):
    y0, prm, span = lanes(prob, 96)
    for stiff in (5000, 300, 0):
        kw = dict(rtol=1e-8, atol=1e-10, nfev_stiff_detect=stiff)
        o = oracle(prob, span, y0, m, prm, **kw)
        g = emu.solve(prob, span, y0, m, prm, fast=True, **kw)
        assert g["used_fast"] == (m in GENERIC)
        same(g, o, (m.__name__, prob, stiff, "fast"))
        if m in GENERIC:                      # the generic kernel on the same input
            g2 = emu.solve(prob, span, y0, m, prm, fast=False, **kw)
            assert not g2["used_fast"]
            same(g2, o, (m.__name__, prob, stiff, "generic"))


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_kernel_source_dense_output(m):
    y0, prm, _ = lanes("lorenz63", 48, seed=7)
    t_eval = np.linspace(0.0, 4.0, 131)
    kw = dict(rtol=1e-7, atol=1e-9, t_eval=t_eval)
    o = oracle("lorenz63", (0.0, 4.0), y0, m, prm, **kw)
    g = emu.solve("lorenz63", (0.0, 4.0), y0, m, prm, **kw)
    same(g, o, (m.__name__, "t_eval"), dense=True)


@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_ckdisc_kernel_source_equals_oracle_bit_for_bit(prob):
    """Lane::attempt_ckdisc (cash.py:245-416: assessment after stages 1 and 3,
    fallback solutions, twiddle / quit adaptation, Horner or cubic output)
    against ck_solve_one of oracle/xsq_oracle.c in device arithmetic."""
    tab = O.load_ckdisc()
    y0, prm, span = lanes(prob, 96)
    for kw in (dict(rtol=1e-8, atol=1e-10), dict(rtol=1e-4, atol=1e-6),
               dict(rtol=1e-6, atol=1e-8, t_eval=np.linspace(span[0], span[1], 57)),
               dict(rtol=1e-5, atol=1e-7, max_step=0.02 * (span[1] - span[0])),
               dict(rtol=1e-3, atol=1e-3, max_steps=40)):
        with CO.device_math():
            o = CO.rk_batch(tab, prob, span, y0, params=prm, n_threads=CO.max_threads(), **kw)
        g = emu.solve(prob, span, y0, xb.CKdisc, prm, **kw)
        assert o["n_rejected"].sum() > 0
        same(g, o, ("CKdisc", prob, sorted(kw)), dense="t_eval" in kw)
    # backward in time
    with CO.device_math():
        o = CO.rk_batch(tab, prob, (span[1] * 0.1, 0.0), y0, params=prm, rtol=1e-6, atol=1e-8)
    same(emu.solve(prob, (span[1] * 0.1, 0.0), y0, xb.CKdisc, prm, rtol=1e-6, atol=1e-8), o,
         ("CKdisc", prob, "backward"))


@pytest.mark.parametrize("kw", [
    dict(rtol=1e-3, atol=1e-6), dict(rtol=1e-11, atol=1e-13),
    dict(rtol=1e-6, atol=[1e-9, 1e-7, 1e-8]), dict(rtol=1e-8, atol=1e-10, max_step=0.02),
    dict(rtol=1e-8, atol=1e-10, first_step=1e-3),
    dict(rtol=1e-6, atol=1e-8, max_step=0.3, first_step=0.25),
    dict(rtol=1e-6, atol=1e-8, sc_params=(0.6, -0.2, 0.0, 0.9)),
    dict(rtol=1e-6, atol=1e-8, sc_params=(0.7, -0.4, 0.1, 0.8)),
], ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items() if k != "atol"))
def test_kernel_source_options(kw):
    y0, prm, _ = lanes("lorenz63", 64, seed=21)
    for m in (xb.Ts5, xb.BS5, xb.Pr7):
        for span in ((0.0, 3.0), (0.4, 0.0)):
            o = oracle("lorenz63", span, y0, m, prm, **kw)
            same(emu.solve("lorenz63", span, y0, m, prm, fast=True, **kw), o, (m.__name__, kw, span))
            same(emu.solve("lorenz63", span, y0, m, prm, fast=False, **kw), o, (m.__name__, kw, span))


def test_kernel_source_failures_and_edge_spans():
    # explicit method on a very stiff lane with absurd tolerances: rejected steps, tiny steps
    y0 = np.tile([2.0, 0.0], (8, 1))
    mu = np.full((8, 1), 1e9)
    for m in (xb.Ts5, xb.Pr9, xb.CFMR7osc):
        kw = dict(rtol=1e-12, atol=1e-14)
        o = oracle("vanderpol", (0.0, 1e-5), y0, m, mu, **kw)
        same(emu.solve("vanderpol", (0.0, 1e-5), y0, m, mu, **kw), o, (m.__name__, "stiff"))
    # overflow
    y0 = np.tile([1e200, 1e200, 1e200], (8, 1))
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (8, 1))
    for m in (xb.Ts5, xb.BS5):
        o = oracle("lorenz63", (0.0, 1.0), y0, m, prm, rtol=1e-8, atol=1e-10)
        g = emu.solve("lorenz63", (0.0, 1.0), y0, m, prm, rtol=1e-8, atol=1e-10)
        assert (g["status"] != 0).all()
        same(g, o, (m.__name__, "overflow"))
    # zero-length span, span shorter than the first step, time far from zero
    y0, prm, _ = lanes("lorenz63", 16, seed=5)
    for span in ((1.0, 1.0), (0.0, 1e-9), (1e6, 1e6 + 2.0), (-3.0, -1.0)):
        o = oracle("lorenz63", span, y0, xb.Ts5, prm, rtol=1e-7, atol=1e-9)
        same(emu.solve("lorenz63", span, y0, xb.Ts5, prm, rtol=1e-7, atol=1e-9), o, span)


def test_kernel_source_probe_queue_and_slots_agree():
    y0, prm, _ = lanes("lorenz63", 80, seed=11)
    kw = dict(rtol=1e-8, atol=1e-10, nfev_stiff_detect=240)
    o = oracle("lorenz63", (0.0, 8.0), y0, xb.Ts5, prm, **kw)
    assert o["nfev"].sum() > 0
    for q in (0, 7, 1000, -1):
        for fast in (True, False):
            g = emu.solve("lorenz63", (0.0, 8.0), y0, xb.Ts5, prm, fast=fast, queue_records=q, **kw)
            same(g, o, ("queue", q, fast))


@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_swag_kernel_source_equals_oracle_bit_for_bit(prob):
    """swag_persistent_body (xsq_swag_core.cuh) on the host against
    oracle/xsq_oracle_swag.c in device arithmetic, with and without t_eval."""
    y0, prm, span = lanes(prob, 96)
    for te in (None, np.linspace(span[0], span[1], 57)):
        for kw in (dict(rtol=1e-8, atol=1e-10), dict(rtol=1e-4, atol=1e-7, k_max=5),
                   dict(rtol=1e-6, atol=1e-9, max_step=0.05 * (span[1] - span[0]))):
            okw = dict(kw, t_eval=te)
            with CO.device_math():
                o = CO.swag_batch(prob, span, y0, params=prm, n_threads=CO.max_threads(), **okw)
            g = emu.solve(prob, span, y0, xb.SWAG, prm, **okw)
            for k in ("n_accepted", "n_rejected", "nfev", "status"):
                assert np.array_equal(g[k], o[k]), (prob, kw, k)
            for k in ("t_final", "y_final"):
                assert np.array_equal(bits(g[k]), bits(o[k])), (prob, kw, k)
            if te is not None:
                ok = (bits(g["y"]) == bits(o["y"])) | (np.isnan(g["y"]) & np.isnan(o["y"]))
                assert ok.all(), (prob, kw, "y(t_eval)")


# ---- events (scipy's `events=`): the kernels as NVRTC builds them, on the host ----
def _lorenz_event_lanes(N, seed=11):
    rng = np.random.default_rng(seed)
    y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
    prm = np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1)
    return y0, prm


def _same_events(a, b, what, keys):
    for k in keys:
        x, y = a[k], b[k]
        ok = (x == y) | ((x != x) & (y != y))
        assert ok.all(), (what, k, np.argwhere(~ok)[:4])


@pytest.mark.parametrize("m", [xb.Ts5, xb.BS5, xb.Pr8, xb.CKdisc], ids=lambda m: m.__name__)
def test_event_kernel_sources_in_lane_queue_and_fast_agree(m):
    """xsq_rk_core.cuh / xsq_rk_fast.cuh compiled with XSQ_EVENTS_N = 3 (three
    Lorenz section functions): roots located inside the lane (no queue), by
    event_queue_body from the queued steps (rk_persistent), and -- generic pair,
    no t_eval -- by rk_fast with event hooks (terminal occurrences through the
    out-of-line in-lane solver).  Every event time and
    state, count and trajectory output must be equal bit for bit, and the
    trajectory itself must be the one the C oracle computes without events."""
    y0, prm = _lorenz_event_lanes(72)
    te = np.linspace(0.0, 5.0, 41)
    keys = ("t_events", "y_events", "event_counts", "y_final", "t_final", "h_next", "nfev",
            "n_accepted", "n_rejected", "status", "stiff_flags")
    generic = m in (xb.Ts5, xb.Pr8)
    for term, kw in (([0, 0, 0], {}), ([0, 0, 0], dict(nfev_stiff_detect=0)), ([0, 2, 0], {}), ([2, 0, 3], dict(nfev_stiff_detect=0)),
                     ([0, 0, 0], dict(t_eval=te)), ([3, 0, 0], dict(t_eval=te))):
        if m is xb.CKdisc:
            kw = dict(kw, nfev_stiff_detect=0)
        base = dict(rtol=1e-7, atol=1e-9, max_event_records=12, events=(term, [1, 0, -1]), **kw)
        a = emu.solve("lorenz63", (0.0, 5.0), y0, m, prm, event_queue_records=0, **base)
        b = emu.solve("lorenz63", (0.0, 5.0), y0, m, prm, event_queue_records=700, **base)
        c = emu.solve("lorenz63", (0.0, 5.0), y0, m, prm, event_queue_records=-1, fast=False, **base)
        d = emu.solve("lorenz63", (0.0, 5.0), y0, m, prm, event_queue_records=-1, fast=True, **base)
        assert not a["used_fast"] and not b["used_fast"] and not c["used_fast"]
        assert d["used_fast"] == (generic and "t_eval" not in kw)
        assert a["event_counts"].sum() > (2 if any(term) else 5) * len(y0)
        if any(term):
            assert (a["status"] == 1).sum() > len(y0) // 4      # lanes stopped by the event
        k2 = keys + (("y",) if "t_eval" in kw else ())
        others = [(b, "overflowing queue"), (c, "queue"), (d, "fast")]
        if not any(term):
            # the build NVRTC makes for event sets without a terminal event
            e = emu.solve("lorenz63", (0.0, 5.0), y0, m, prm, event_queue_records=-1, fast=True,
                          no_terminal_build=True, **base)
            assert e["used_fast"] == d["used_fast"]
            others.append((e, "fast, no-terminal build"))
        for other, name in others:
            _same_events(a, other, (m.__name__, term, sorted(kw), name), k2)
        # ... and against the INDEPENDENT restatement of the event handling in the C
        # oracle (events_after_step, brentq_c, dense_build / dense_eval of
        # oracle/xsq_oracle.c in device arithmetic): every event record, count,
        # status and trajectory output bit for bit
        tab_o = O.load_ckdisc() if m is xb.CKdisc else TABS[m.__name__]
        okw = {k: v for k, v in base.items() if k not in ("events", "max_event_records")}
        with CO.device_math():
            oe = CO.rk_events_batch(tab_o, "lorenz63", (0.0, 5.0), y0, "lorenz_sections", term,
                                    [1, 0, -1], 12, params=prm, n_threads=CO.max_threads(), **okw)
        _same_events(a, oe, (m.__name__, term, sorted(kw), "C oracle with events"), k2)
        if term == [0, 0, 0]:
            # events never change t, y, h: the plain solve of the oracle
            okw = {k: v for k, v in base.items() if k not in ("events", "max_event_records")}
            with CO.device_math():
                tab = O.load_ckdisc() if m is xb.CKdisc else TABS[m.__name__]
                o = CO.rk_batch(tab, "lorenz63", (0.0, 5.0), y0, params=prm,
                                n_threads=CO.max_threads(), **okw)
            if m is xb.BS5:
                # the extra stages of BS5's interpolant on steps with an event count in
                # nfev (bogacki.py:372-388), as in the reference: everything but nfev
                for k in ("n_accepted", "n_rejected", "status"):
                    assert np.array_equal(a[k], o[k]), (k,)
                for k in KEYS_F:
                    assert np.array_equal(bits(a[k]), bits(o[k])), (k,)
                assert (a["nfev"] >= o["nfev"]).all() and (a["nfev"] > o["nfev"]).any()
            else:
                same(a, o, (m.__name__, "events vs plain oracle"), dense="t_eval" in kw)


def test_event_kernel_sources_against_the_numpy_restatement():
    """Event times and states of the emulated kernels against the restated
    reference driven by scipy's event loop (oracle/rk_oracle.py, bit-identical
    to the reference on the event goldens)."""
    from oracle.problems import EVENT_SETS, make_fun
    y0, prm = _lorenz_event_lanes(10, seed=5)
    term, direc = [0, 4, 0], [1, 0, -1]
    kw = dict(rtol=1e-7, atol=1e-9)
    g = emu.solve("lorenz63", (0.0, 6.0), y0, xb.Ts5, prm, events=(term, direc),
                  max_event_records=32, **kw)
    fns = EVENT_SETS["lorenz_sections"][0]
    checked = 0
    for i in range(len(y0)):
        o = O.rk_solve(TABS["Ts5"], make_fun("lorenz63", prm[i]), (0.0, 6.0), y0[i],
                       events=[(f, a, b) for f, a, b in zip(fns, term, direc)], **kw)
        assert g["status"][i] == o["status"]
        if g["nfev"][i] != o["nfev"]:
            continue                     # a flipped accept/reject decision (other arithmetic)
        checked += 1
        for k in range(3):
            tg = o["t_events"][k]
            assert g["event_counts"][i, k] >= tg.size
            assert np.allclose(g["t_events"][i, k, :tg.size], tg, rtol=1e-9, atol=1e-9)
            if tg.size:
                assert np.allclose(g["y_events"][i, k, :tg.size], o["y_events"][k], rtol=1e-7, atol=1e-7)
        assert abs(g["t_final"][i] - o["t_final"]) <= 1e-9
    assert checked >= 8 and (g["status"] == 1).any()


def test_swag_event_kernel_source_equals_the_c_oracle_with_events():
    """SwagLane::finish_step (events on SWAG's interpolant, xsq_swag_core.cuh built
    with XSQ_EVENTS_N) against the C restatement (oracle/xsq_oracle_swag.c): event
    times and states, counts, terminal stops, t_eval output, trajectory -- bit
    for bit."""
    y0, prm = _lorenz_event_lanes(72)
    te = np.linspace(0.0, 5.0, 41)
    for term, kw in (([0, 0, 0], {}), ([0, 2, 0], {}), ([2, 0, 3], dict(k_max=5)),
                     ([0, 0, 0], dict(t_eval=te)), ([3, 0, 0], dict(t_eval=te))):
        base = dict(rtol=1e-7, atol=1e-9, **kw)
        a = emu.solve("lorenz63", (0.0, 5.0), y0, xb.SWAG, prm, events=(term, [1, 0, -1]),
                      max_event_records=12, **base)
        with CO.device_math():
            o = CO.swag_events_batch("lorenz63", (0.0, 5.0), y0, "lorenz_sections", term, [1, 0, -1],
                                     12, params=prm, n_threads=CO.max_threads(), **base)
        assert o["event_counts"].sum() > 2 * len(y0)
        if any(term):
            assert (o["status"] == 1).sum() > len(y0) // 4
        keys = ("t_events", "y_events", "event_counts", "y_final", "t_final", "nfev", "n_accepted",
                "n_rejected", "status") + (("y",) if "t_eval" in kw else ())
        _same_events(a, o, ("SWAG", term, sorted(kw)), keys)


# ---- Runge-Kutta-Nystrom methods ---------------------------------------------------
@pytest.mark.parametrize("m", [xb.Fi4N, xb.Fi5N, xb.Mu5Nmb], ids=lambda m: m.__name__)
@pytest.mark.parametrize("prob", ["vanderpol", "arenstorf"])
def test_nystrom_kernel_source_equals_oracle_bit_for_bit(m, prob):
    """The Nystrom variant of rk_persistent (Lane::stage / solution / error for
    tab::NYSTROMV, common.py:1279-1309) on the host against rkn_stage and the
    Nystrom block of oracle/xsq_oracle.c in device arithmetic.  (MR6NN takes
    velocity independent problems only: covered on the GPU with a Kepler user
    right-hand side, tests/test_gpu_rkn.py.)"""
    tab = O.load_tableaux_rkn()[m.__name__]
    y0, prm, span = lanes(prob, 64)
    if prob == "vanderpol":
        keep = prm[:, 0] <= 30.0
        y0, prm = y0[keep], prm[keep]
    for kw in (dict(rtol=1e-8, atol=1e-10), dict(rtol=1e-4, atol=1e-6)):
        with CO.device_math():
            o = CO.rk_batch(tab, prob, span, y0, params=prm, n_threads=CO.max_threads(),
                            nfev_stiff_detect=0, **kw)
        g = emu.solve(prob, span, y0, m, prm, nfev_stiff_detect=0, **kw)
        assert o["n_rejected"].sum() > 0 and (o["status"] == 0).all()
        same(g, o, (m.__name__, prob, sorted(kw)))


def test_event_kernels_random_options_against_the_c_oracle():
    """80 seeded random settings -- method (Ts5, BS5, Pr8, CKdisc, SWAG), span
    (forward / backward), tolerances, max_step, first_step, t_eval, terminal counts
    and directions, record capacity, BS5 interpolant, stiffness diagnosis, step
    budget, event queue size (exact / overflowing / none), fast kernel allowed or
    not: the emulated kernels and the C oracle with events must agree bit for
    bit on every output.  (This search found the one disagreement fixed in the
    oracle: h_next of a lane that runs out of its step budget.)"""
    rng = np.random.default_rng(2026)
    methods = [xb.Ts5, xb.BS5, xb.Pr8, xb.CKdisc, xb.SWAG]
    for _ in range(80):
        m = methods[rng.integers(len(methods))]
        N = int(rng.integers(4, 16))
        y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
        prm = np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1)
        back = rng.random() < 0.25          # (Lorenz backwards blows up: short spans only)
        T = float(rng.uniform(0.05, 0.4)) if back else float(rng.uniform(0.5, 3.0))
        span = (T, 0.0) if back else (0.0, T)
        term = [int(rng.integers(0, 4)) if rng.random() < 0.4 else 0 for _ in range(3)]
        direc = [int(rng.integers(-1, 2)) for _ in range(3)]
        kw = dict(rtol=float(10 ** rng.uniform(-9, -3)), atol=float(10 ** rng.uniform(-11, -5)))
        if rng.random() < 0.3:
            kw["max_step"] = float(rng.uniform(0.01, 0.3))
        if rng.random() < 0.3:
            kw["first_step"] = float(rng.uniform(1e-4, 1e-2))
        if rng.random() < 0.4:
            te = np.unique(rng.uniform(0, T, int(rng.integers(1, 30))))
            kw["t_eval"] = te[::-1].copy() if back else te
        if rng.random() < 0.3:
            kw["max_steps"] = int([3000, 200][rng.integers(2)])
        cap = int(rng.integers(1, 10))
        swag = m is xb.SWAG
        if not swag:
            if m is xb.BS5 and rng.random() < 0.7:
                kw["interpolant"] = ["best", "low", "free"][rng.integers(3)]
            kw["nfev_stiff_detect"] = 0 if m is xb.CKdisc else int([0, 300, 5000][rng.integers(3)])
        q = int([-1, -1, 0, 40][rng.integers(4)])
        what = (m.__name__, span, term, direc, sorted(kw), cap, q)
        a = emu.solve("lorenz63", span, y0, m, prm, events=(term, direc), max_event_records=cap,
                      event_queue_records=q, fast=bool(rng.random() < 0.7), **kw)
        with CO.device_math():
            if swag:
                o = CO.swag_events_batch("lorenz63", span, y0, "lorenz_sections", term, direc, cap,
                                         params=prm, n_threads=CO.max_threads(), **kw)
            else:
                tab = O.load_ckdisc() if m is xb.CKdisc else TABS[m.__name__]
                o = CO.rk_events_batch(tab, "lorenz63", span, y0, "lorenz_sections", term, direc, cap,
                                       params=prm, n_threads=CO.max_threads(), **kw)
        keys = ["t_events", "y_events", "event_counts", "y_final", "t_final", "nfev", "n_accepted",
                "n_rejected", "status"] + (["y"] if "t_eval" in kw else []) + \
            ([] if swag else ["h_next", "stiff_flags"])
        _same_events(a, o, what, keys)
