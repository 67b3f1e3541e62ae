"""GPU parity of event detection (scipy's solve_ivp `events=` on the device;
SURVEY.md section 8f rank 3), through the C ABI, against the golden vectors
of the reference + scipy and the NumPy restatement on seeded ensembles."""
import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle import rk_oracle as RO
from oracle.problems import CUDA_SOURCES, EVENT_SETS, make_fun
from test_events_golden import (ALL, CASES, TABS, check_reference_event_test, ev_list,
                                ev_options, ev_t_eval, unhex)

pytestmark = pytest.mark.gpu
BUILTIN_PROBLEMS = {"lorenz63", "vanderpol", "arenstorf"}
_RHS, _EV = {}, {}


def rhs_for(problem):
    if problem in BUILTIN_PROBLEMS:
        return problem
    if problem not in _RHS:
        n, p, src = CUDA_SOURCES[problem]
        _RHS[problem] = xb.DeviceRHS.from_source(src, "rhs", n, p)
    return _RHS[problem]


def events_for(name, terminal, direction):
    if name not in _EV:
        py, src = EVENT_SETS[name]
        _EV[name] = xb.DeviceEvents.from_source(src, "event", len(py))
    return _EV[name].with_attributes(terminal=terminal, direction=direction)


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_events_vs_reference_golden(c):
    o = ev_options(c)
    te = ev_t_eval(c)
    ev = events_for(c["events"], c["terminal"], c["direction"])
    r = xb.solve_ivp_batched(rhs_for(c["problem"]), c["t_span"], [c["y0"]],
                             getattr(xb, c["method"]),
                             params=[c["params"]] if c["params"] else None, t_eval=te,
                             events=ev, max_event_records=32, max_steps=100000, **o)
    torch.cuda.synchronize()
    assert int(r.status[0]) == c["status"]
    rtol = o.get("rtol", 1e-3)
    # the step sequence is the reference's (same counts), so event times agree
    # to rounding amplified by the root finder's conditioning, far below rtol
    if c["method"] == "SWAG":        # variable order: one flipped order choice moves nfev
        assert abs(int(r.nfev[0]) - c["nfev"]) <= 0.02 * c["nfev"] + 2
        tol = 1e-9 if int(r.nfev[0]) == c["nfev"] else 10 * rtol
    else:
        assert int(r.nfev[0]) == c["nfev"] and int(r.n_rejected[0]) == c["nfs"]
        tol = 1e-9
    ytol = 1e-8 if tol == 1e-9 else 10 * rtol
    cnt = r.event_counts.cpu().numpy()[0]
    t_ev = r.t_events.cpu().numpy()[0]
    y_ev = r.y_events.cpu().numpy()[0]
    for k in range(len(c["terminal"])):
        tg = unhex(c["t_events"][k])
        yg = unhex(c["y_events"][k])
        if c["status"] == 0:
            assert cnt[k] == tg.size, (k, cnt[k], tg.size)
        # after a terminal stop scipy drops the roots behind it; the count keeps them
        assert cnt[k] >= tg.size
        assert np.allclose(t_ev[k, :tg.size], tg, rtol=tol, atol=tol), (k, t_ev[k, :tg.size], tg)
        if tg.size:
            assert np.allclose(y_ev[k, :tg.size], yg.reshape(tg.size, -1), rtol=ytol, atol=ytol)
        if c["status"] == 0:
            assert np.isnan(t_ev[k, tg.size:]).all()
    t_g, y_g = unhex(c["t"]), unhex(c["y"])
    if te is None:
        assert abs(float(r.t_final[0]) - t_g[-1]) <= tol * max(1.0, abs(t_g[-1]))
        assert np.allclose(r.y_final.cpu().numpy()[0], y_g[:, -1], rtol=ytol, atol=ytol)
    else:
        # t_eval output stops at the terminal event (ivp.py: t = roots[-1])
        assert int(r.n_eval_done[0]) == t_g.size
        y = r.y.cpu().numpy()[0][:, :t_g.size]
        assert np.allclose(y, y_g.reshape(y.shape), rtol=ytol, atol=ytol)
    assert rtol > 0


def test_events_ensemble_poincare_section_vs_oracle():
    """48 Lorenz lanes, upward crossings of z = 27 (non-terminal) and the 4th
    zero of x (terminal): per-lane counts, times and states."""
    N = 48
    rng = np.random.default_rng(5)
    y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (N, 1))
    kw = dict(rtol=1e-7, atol=1e-9)
    term, direc = [0, 4, 0], [1, 0, -1]
    ev = events_for("lorenz_sections", term, direc)
    r = xb.solve_ivp_batched("lorenz63", (0.0, 6.0), y0, xb.Ts5, params=prm, events=ev,
                             max_event_records=32, **kw)
    torch.cuda.synchronize()
    cnt = r.event_counts.cpu().numpy()
    t_ev = r.t_events.cpu().numpy()
    status = r.status.cpu().numpy()
    fns = EVENT_SETS["lorenz_sections"][0]
    same = 0
    for i in range(N):
        o = RO.rk_solve(TABS["Ts5"], make_fun("lorenz63", prm[i]), (0.0, 6.0), y0[i],
                        events=[(g, a, b) for g, a, b in zip(fns, term, direc)], **kw)
        assert status[i] == o["status"]
        if int(r.nfev[i]) != o["nfev"]:
            continue                     # a flipped accept/reject decision (chaotic system)
        same += 1
        for k in range(3):
            tg = o["t_events"][k]
            assert cnt[i, k] >= tg.size
            assert np.allclose(t_ev[i, k, :tg.size], tg, rtol=1e-8, atol=1e-8)
        assert abs(float(r.t_final[i]) - o["t_final"]) <= 1e-8
        assert np.allclose(r.y_final[i].cpu().numpy(), o["y_final"], rtol=1e-6, atol=1e-6)
    assert same >= 0.9 * N
    assert (status == 1).any() and (cnt[:, 0] > 0).all()


def test_events_leave_the_trajectory_untouched_and_validate_arguments():
    N = 256
    rng = np.random.default_rng(1)
    y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (N, 1))
    kw = dict(params=prm, rtol=1e-6, atol=1e-9)
    ev = events_for("lorenz_sections", [0, 0, 0], [0, 0, 0])
    a = xb.solve_ivp_batched("lorenz63", (0.0, 3.0), y0, xb.Pr8, events=ev, **kw)
    b = xb.solve_ivp_batched("lorenz63", (0.0, 3.0), y0, xb.Pr8, **kw)
    torch.cuda.synchronize()
    # (the kernel with events is compiled at run time by NVRTC, the other one
    # by nvcc: same source and flags, not necessarily the same schedule)
    assert torch.equal(a.nfev, b.nfev) and torch.equal(a.n_rejected, b.n_rejected)
    assert torch.allclose(a.y_final, b.y_final, rtol=1e-9, atol=1e-9)
    assert (a.status == 0).all() and (a.event_counts.sum(1) > 0).all()
    # more occurrences than records: counted, the first `capacity` are kept
    c = xb.solve_ivp_batched("lorenz63", (0.0, 3.0), y0, xb.Pr8, events=ev,
                             max_event_records=2, **kw)
    torch.cuda.synchronize()
    assert torch.equal(c.event_counts, a.event_counts)
    assert torch.equal(c.t_events[:, :, :2].nan_to_num(-1.0), a.t_events[:, :, :2].nan_to_num(-1.0))
    with pytest.raises(ValueError):
        xb.solve_ivp_batched("lorenz63", (0.0, 1.0), y0, xb.Ts5, events=ev,
                             forced_steps=[0.01, 0.01], **kw)
    with pytest.raises(ValueError):
        xb.DeviceEvents.from_source("x", "event", 2, terminal=[-1, 0])


# ---- the reference's own event tests (tests/test_ivp.py:369-470, 757-783) ------
def gpu_solver(method, span, y0, events, **kw):
    ev = events_for("rational", [e[1] for e in events], [e[2] for e in events])
    r = xb.solve_ivp_batched(rhs_for("rational"), span, [y0], getattr(xb, method), events=ev, **kw)
    torch.cuda.synchronize()
    cnt = r.event_counts.cpu().numpy()[0]
    t_ev, y_ev = r.t_events.cpu().numpy()[0], r.y_events.cpu().numpy()[0]
    keep = [int(np.isfinite(t_ev[k]).sum()) for k in range(len(events))]
    assert all(k <= c for k, c in zip(keep, cnt))
    return dict(status=int(r.status[0]), t_events=[t_ev[k, :n] for k, n in enumerate(keep)],
                y_events=[y_ev[k, :n] for k, n in enumerate(keep)])


@pytest.mark.parametrize("method", ALL + ["SWAG"])
def test_reference_event_test_on_the_device(method):
    check_reference_event_test(method, gpu_solver)


@pytest.mark.parametrize("method", ALL + ["SWAG"])
def test_reference_t_eval_early_event_on_the_device(method):
    te = np.linspace(7.5, 9, 16)
    ev = events_for("early", [1], [0])
    r = xb.solve_ivp_batched(rhs_for("rational"), [5, 9], [[1 / 3, 2 / 9]], getattr(xb, method),
                             t_eval=te, events=ev)
    torch.cuda.synchronize()
    assert int(r.status[0]) == 1 and r.message(0) == "A termination event occurred."
    assert int(r.n_eval_done[0]) == 0
    assert float(r.t_events[0, 0, 0]) == 7.0 and float(r.t_final[0]) == 7.0


def test_events_through_the_host_buffer_c_entry_point():
    """xsq_rk_solve_host with events: host arrays in and out, unused records NaN,
    same numbers as the device-buffer path."""
    import ctypes as C
    from extensisq_b200 import _lib
    lib = _lib.load()
    N, cap, ne = 200, 8, 3
    rng = np.random.default_rng(4)
    y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (N, 1))
    ev = events_for("lorenz_sections", [0, 2, 0], [1, 0, 0])
    y0_soa, prm_soa = np.ascontiguousarray(y0.T), np.ascontiguousarray(prm.T)
    atol = np.array([1e-9])
    out = dict(t_final=np.empty(N), y_final=np.empty((3, N)),
               t_events=np.zeros((N, ne, cap)), y_events=np.zeros((N, ne, cap, 3)))
    ints = {k: np.empty(N, np.int32) for k in ("n_accepted", "n_rejected", "nfev", "status")}
    ev_count = np.empty((N, ne), np.int32)
    term = (C.c_int32 * ne)(*ev.terminal)
    direc = (C.c_int32 * ne)(*ev.direction)
    a = _lib.XsqRkArgs()
    a.struct_size = C.sizeof(_lib.XsqRkArgs)
    a.method, a.rhs, a.n_state, a.n_param = _lib.METHOD_IDS["Ts5"], 0, 3, 3
    a.n_lanes = N
    a.y0, a.params = y0_soa.ctypes.data, prm_soa.ctypes.data
    a.t0, a.t_bound, a.rtol = 0.0, 3.0, 1e-6
    a.atol = atol.ctypes.data_as(C.POINTER(C.c_double))
    a.n_atol = 1
    a.max_step, a.max_steps, a.nfev_stiff_detect = float("inf"), 1000000, 5000
    a.t_final, a.y_final = out["t_final"].ctypes.data, out["y_final"].ctypes.data
    for k, v in ints.items():
        setattr(a, k, v.ctypes.data)
    a.events, a.n_event_fns = ev.handle, ne
    a.ev_terminal = C.cast(term, C.POINTER(C.c_int32))
    a.ev_direction = C.cast(direc, C.POINTER(C.c_int32))
    a.ev_capacity = cap
    a.t_events, a.y_events = out["t_events"].ctypes.data, out["y_events"].ctypes.data
    a.ev_count = ev_count.ctypes.data
    assert lib.xsq_rk_solve_host(C.byref(a), 0) == 0, lib.xsq_last_error_detail().decode()
    r = xb.solve_ivp_batched("lorenz63", (0.0, 3.0), y0, xb.Ts5, params=prm, rtol=1e-6, atol=1e-9,
                             events=ev, max_event_records=cap, max_steps=1000000)
    torch.cuda.synchronize()
    assert np.array_equal(ints["status"], r.status.cpu().numpy()) and (ints["status"] == 1).any()
    assert np.array_equal(ev_count, r.event_counts.cpu().numpy())
    assert np.array_equal(out["t_events"], r.t_events.cpu().numpy(), equal_nan=True)
    assert np.array_equal(out["y_events"], r.y_events.cpu().numpy(), equal_nan=True)
    assert np.isnan(out["t_events"]).any()                    # unused records
    assert np.array_equal(out["y_final"].T, r.y_final.cpu().numpy())


def test_events_with_a_user_tableau():
    """A user RungeKutta subclass (uploaded tableau, NVRTC kernel) with events:
    Heun's 2(1) pair on the rational problem, against the restated reference."""
    from oracle.problems import make_fun

    class Heun(xb.RungeKutta):
        n_stages, order, order_secondary = 2, 2, 1
        A = np.array([[0.0, 0.0], [1.0, 0.0]])
        B = np.array([0.5, 0.5])
        C = np.array([0.0, 1.0])
        E = np.array([-0.5, 0.5, 0.0])
        P = np.array([[1.0, -0.5], [0.0, 0.5], [0.0, 0.0]])

    tab = RO.Tableau(dict(name="Heun", n_stages=2, order=2, order_secondary=1, sc_params="standard",
                          A=[[v.hex() for v in r] for r in Heun.A.tolist()],
                          B=[v.hex() for v in Heun.B.tolist()], C=[v.hex() for v in Heun.C.tolist()],
                          E=[v.hex() for v in Heun.E.tolist()],
                          P=[[v.hex() for v in r] for r in Heun.P.tolist()]))
    fns = EVENT_SETS["rational"][0]
    term, direc = [0, 0, 1], [0, 0, 0]
    o = RO.rk_solve(tab, make_fun("rational", []), [5, 8], [1 / 3, 2 / 9], rtol=1e-4, atol=1e-7,
                    events=[(g, a, b) for g, a, b in zip(fns, term, direc)])
    ev = events_for("rational", term, direc)
    r = xb.solve_ivp_batched(rhs_for("rational"), [5, 8], [[1 / 3, 2 / 9]], Heun, rtol=1e-4, atol=1e-7,
                             events=ev, max_steps=100000)
    torch.cuda.synchronize()
    assert int(r.status[0]) == o["status"] == 1
    assert int(r.nfev[0]) == o["nfev"]
    for k in range(3):
        tg = o["t_events"][k]
        assert np.allclose(r.t_events.cpu().numpy()[0, k, :tg.size], tg, rtol=1e-9, atol=1e-9)
    assert abs(float(r.t_final[0]) - 7.4) < 1e-12


@pytest.mark.parametrize("method", ["Ts5", "BS5", "Pr8", "CKdisc"])
def test_event_queue_and_in_lane_root_solves_are_bit_identical(method, monkeypatch):
    """Steps whose sign changes cannot end the trajectory are queued and their
    roots located by event_queue_body after the persistent kernel; what does not
    fit in the queue (or may be terminal) is solved inside the lane.  Same
    dense_build / brentq / dense_eval code on the same stages: every event time
    and state, count and final state must be equal bit for bit whether the queue
    is off (0 records), overflows (a few records) or takes everything."""
    N = 3000
    rng = np.random.default_rng(11)
    y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
    prm = np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1)
    te = np.linspace(0.0, 5.0, 41)
    # without t_eval the default run (queue large enough for every record) takes
    # the fast kernel (rk_fast + event hooks) for the generic pairs, with and
    # without the stiffness diagnosis, terminal occurrences included; the runs
    # with a limited queue take rk_persistent: the comparison covers both kernels
    for term, kw in (([0, 0, 0], {}), ([0, 0, 0], dict(nfev_stiff_detect=0)), ([0, 2, 0], {}),
                     ([2, 0, 3], dict(nfev_stiff_detect=0)), ([0, 0, 0], dict(t_eval=te)),
                     ([3, 0, 0], dict(t_eval=te))):
        ev = events_for("lorenz_sections", term, [1, 0, -1])
        runs = []
        for q in ("0", "1500", None):
            if q is None:
                monkeypatch.delenv("XSQ_EVENT_QUEUE_RECORDS", raising=False)
            else:
                monkeypatch.setenv("XSQ_EVENT_QUEUE_RECORDS", q)
            r = xb.solve_ivp_batched("lorenz63", (0.0, 5.0), y0, getattr(xb, method), params=prm,
                                     rtol=1e-7, atol=1e-9, events=ev, max_event_records=12, **kw)
            torch.cuda.synchronize()
            runs.append({k: getattr(r, k).cpu().numpy() for k in
                         ("t_events", "y_events", "event_counts", "y_final", "t_final", "status",
                          "nfev", "n_accepted", "n_rejected", "h_next", "stiff_flags") +
                         (("y",) if "t_eval" in kw else ())})
        assert runs[0]["event_counts"].sum() > (2 if any(term) else 5) * N
        if any(term):
            assert (runs[0]["status"] == 1).sum() > N // 4      # lanes stopped by the event
        for other in runs[1:]:
            for k, a in runs[0].items():
                b = other[k]
                same = (a == b) | ((a != a) & (b != b))
                assert same.all(), (method, term, k, np.argwhere(~same)[:4])
