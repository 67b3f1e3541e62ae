"""CPU tests of the host side: tableau classes (same attribute protocol and
algebraic properties the reference checks in tests/test_rk.py:14-72), argument
validation with the reference's exception types/messages, and ensemble
sharding over ranks (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import extensisq_b200 as xb
from extensisq_b200 import batched

METHODS = [xb.Ts5, xb.BS5, xb.CK5, xb.Me4, xb.Pr7, xb.Pr8, xb.Pr9, xb.CFMR7osc]


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_coefficient_properties(m):
    # reference tests/test_rk.py:45-72
    s = m.n_stages
    assert m.A.shape == (s, s) and m.B.shape == (s,) and m.C.shape == (s,)
    assert m.E.shape == (s + 1,) and m.P.shape[0] == s + 1
    assert abs(m.B.sum() - 1) < 1e-15
    assert abs(m.E.sum()) < 1e-15
    assert np.all(np.abs(m.A.sum(axis=1) - m.C) < 1e-13)
    assert np.all(np.triu(m.A) == 0)
    # interpolant continuity, same formulation/tolerances as the reference
    Ps = m.P.sum(axis=1)                       # C0 end
    Ps[:s] -= m.B
    assert np.all(np.abs(Ps) < 1e-12)
    Ps = m.P.sum(axis=0)                       # C1 start
    Ps[0] -= 1
    assert np.all(np.abs(Ps) < 1e-12)
    dPs = (m.P * (np.arange(m.P.shape[1]) + 1)).sum(axis=1)   # C1 end
    dPs[-1] -= 1
    assert np.all(np.abs(dPs) < 2e-12)


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_low_order_conditions(m):
    """Order conditions up to order 4 for (B, C, A) and for the embedded
    weights B + E[:-1] up to its own order (cf. tests/test_rk.py:14-42)."""
    A, C = m.A, m.C
    for b, order in ((m.B, m.order), (m.B + m.E[:-1], m.order_secondary)):
        if m.E[-1] != 0 and b is not m.B:
            continue            # FSAL pairs use the extra stage; checked on GPU
        tol = m.n_stages * 1e-14
        conds = [(b.sum(), 1.0)]
        if order >= 2:
            conds.append((b @ C, 1 / 2))
        if order >= 3:
            conds += [(b @ C**2, 1 / 3), (b @ (A @ C), 1 / 6)]
        if order >= 4:
            conds += [(b @ C**3, 1 / 4), (b @ (C * (A @ C)), 1 / 8),
                      (b @ (A @ C**2), 1 / 12), (b @ (A @ (A @ C)), 1 / 24)]
        for got, want in conds:
            assert abs(got - want) < tol


def test_class_attributes_are_read_only_and_named_like_the_reference():
    assert [m.__name__ for m in METHODS] == ["Ts5", "BS5", "CK5", "Me4",
                                             "Pr7", "Pr8", "Pr9", "CFMR7osc"]
    assert xb.Ts5.sc_params == "G" and xb.Pr7.sc_params == "S"
    assert xb.BS5.sc_params == "standard" and xb.BS5.n_extra_stages == 3
    assert (xb.Pr9.n_stages, xb.Pr9.order, xb.Pr9.order_secondary) == (17, 9, 7)
    with pytest.raises(ValueError):
        xb.Ts5.A[1, 0] = 0.0


def test_user_tableau_validation():
    class Bad(xb.RungeKutta):
        n_stages = 2
        order, order_secondary = 2, 1
        A = np.array([[0, 1.0], [1.0, 0]])
        B = np.array([0.5, 0.5])
        C = np.array([0, 1.0])
        E = np.array([0.5, -0.5, 0])
    with pytest.raises(ValueError, match="lower triangular"):
        Bad.validate()


def test_sc_params_parsing():
    assert batched._sc_tuple("G") == (0.7, -0.4, 0.0, 0.9)
    assert batched._sc_tuple((0.6, -0.25, 0.1, 0.85)) == (0.6, -0.25, 0.1, 0.85)
    with pytest.raises(ValueError, match="sc_params should be a tuple"):
        batched._sc_tuple("W")


def test_shard_bounds_cover_all_lanes():
    for n, w in ((10, 3), (7, 8), (10_000_000, 8), (0, 2), (5, 1)):
        spans = [xb.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and b >= a
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        xb.shard_bounds(4, 2, 2)


def _gather_worker(rank, world, port, n_lanes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = xb.shard_bounds(n_lanes, rank, world)
    from extensisq_b200.batched import BatchedOdeResult
    lanes = torch.arange(lo, hi, dtype=torch.float64)
    # a shard's result as the solver returns it, events included ([N, n_events,
    # capacity] and [N, n_events, capacity, n])
    res = BatchedOdeResult(
        t=None, y=None, t_final=lanes.clone(), y_final=lanes[:, None] * torch.ones(1, 3, dtype=torch.float64),
        h_next=lanes.clone(), n_accepted=torch.arange(lo, hi, dtype=torch.int32),
        n_rejected=torch.zeros(hi - lo, dtype=torch.int32), nfev=torch.zeros(hi - lo, dtype=torch.int32),
        status=torch.zeros(hi - lo, dtype=torch.int32),
        t_events=lanes[:, None, None] * torch.ones(1, 2, 4, dtype=torch.float64),
        y_events=lanes[:, None, None, None] * torch.ones(1, 2, 4, 3, dtype=torch.float64),
        event_counts=torch.arange(lo, hi, dtype=torch.int32)[:, None] * torch.ones(1, 2, dtype=torch.int32))
    mine = res.lane_tensors()
    assert "y" not in mine and "t_events" in mine
    full = xb.gather_result(mine, n_lanes)
    ok = (torch.equal(full["n_accepted"],
                      torch.arange(n_lanes, dtype=torch.int32))
          and full["y_final"].shape == (n_lanes, 3)
          and torch.equal(full["y_final"][:, 1],
                          torch.arange(n_lanes, dtype=torch.float64))
          and full["t_events"].shape == (n_lanes, 2, 4)
          and full["y_events"].shape == (n_lanes, 2, 4, 3)
          and torch.equal(full["y_events"][:, 1, 2, 0], torch.arange(n_lanes, dtype=torch.float64))
          and torch.equal(full["event_counts"][:, 1], torch.arange(n_lanes, dtype=torch.int32)))
    root = xb.gather_result(mine, n_lanes, dst=0)
    ok = ok and ((root["n_accepted"] is not None) == (rank == 0))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_result_world_size_2_gloo():
    """The N>1 path: lanes are sharded, no data-path collective, one final
    gather."""
    world, n_lanes = 2, 11
    with mp.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_gather_worker, args=(world, port, n_lanes, out),
                 nprocs=world, join=True)
        assert out[0] and out[1]


def test_ckdisc_class_mirrors_the_reference_attributes():
    # cash.py:184-236
    import extensisq_b200 as xb
    ck = xb.CKdisc
    assert (ck.n_stages, ck.order, ck.order_secondary) == (6, 5, 4)
    assert ck.max_factor == 5 and ck.min_factor == 1 / 5
    assert np.array_equal(ck.A, xb.CK5.A) and np.array_equal(ck.C, xb.CK5.C)
    assert np.array_equal(ck.B, xb.CK5.B) and np.array_equal(ck.P, xb.CK5.P)
    assert ck.E.shape == (7,) and ck.E[-1] == 0.0
    assert ck.B_assess.shape == (2, 6) and ck.E_assess.shape == (2, 6)
    assert ck.B_fallback.shape == (2, 6) and ck.E_fallback.shape == (2, 6)
    assert np.array_equal(ck.C_fallback, ck.C[[1, 3]])
    assert ck._xsq_method == 8 and "CKdisc" not in xb.BUILTIN
    assert ck.stbrad is None and ck.tanang is None      # no stiffness diagnosis


def test_device_events_and_sens_forward_argument_checks_need_no_gpu():
    """Validation mirrors scipy's prepare_events and sensitivity.py:135-156 and
    happens on the host before anything is launched."""
    import extensisq_b200 as xb
    src = "__device__ double event(int k, double t, const double* y, const double* p) { return y[0]; }"
    ev = xb.DeviceEvents.from_source(src, "event", 2, terminal=[True, 3], direction=[-2.5, 0])
    assert ev.terminal == [1, 3] and ev.direction == [-1, 0] and ev.handle >= 1
    assert ev.with_attributes(terminal=[0, 0]).handle == ev.handle
    with pytest.raises(ValueError):
        xb.DeviceEvents.from_source(src, "event", 2, terminal=[-1, 0])
    with pytest.raises(ValueError):
        xb.DeviceEvents.from_source(src, "event", 2, terminal=[1])
    with pytest.raises(ValueError):                      # 9 event functions: XSQ_MAX_EVENTS is 8
        xb.DeviceEvents.from_source(src, "event", 9)
    with pytest.raises(AssertionError):                  # dy0dp must be (ny, np)
        xb.sens_forward("", (0.0, 1.0), [[1.0, 1.0, 1.0]], np.zeros((2, 3)), [1.0, 2.0, 3.0])
    with pytest.raises(ValueError):                      # ny (np + 1) > 1024 states per warp
        xb.sens_forward("", (0.0, 1.0), [[1.0] * 300], np.zeros((300, 3)), [1.0, 2.0, 3.0])
    with pytest.raises(AssertionError):                  # t_eval must end at t_span[1]
        xb.sens_forward("", (0.0, 1.0), [[1.0, 1.0]], np.zeros((2, 1)), [1.0], t_eval=[0.0, 0.5])
    with pytest.raises(AssertionError):                  # rtol must be a float
        xb.sens_forward("", (0.0, 1.0), [[1.0, 1.0]], np.zeros((2, 1)), [1.0], rtol=1)


def test_batched_ode_solution_orders_and_scatters_the_requested_times(monkeypatch):
    """BatchedOdeSolution (dense_output=True) without a GPU: the solve it repeats
    is replaced by a stub that validates t_eval like the real one (ivp.py:600-612)
    and returns y = t; any order, repeated points, scalars, both directions."""
    import numpy as np
    import torch
    import extensisq_b200.batched as B

    class R:
        pass

    def fake(fun, t_span, y0, method, t_eval=None, **kw):
        t0, tf = map(float, t_span)
        te = B._as_device(t_eval, "cpu")
        assert te.ndim == 1 and bool(((te >= min(t0, tf)) & (te <= max(t0, tf))).all())
        d = te[1:] - te[:-1]
        assert bool((d > 0).all()) if tf > t0 else bool((d < 0).all())
        r = R()
        r.y = torch.zeros(3, 2, te.numel(), dtype=torch.float64) + te
        return r

    monkeypatch.setattr(B, "solve_ivp_batched", fake)
    for t_span in ((5.0, 9.0), (5.0, 1.0)):
        sol = B.BatchedOdeSolution(None, t_span, None, None, {})
        tc = np.linspace(*t_span)
        assert np.array_equal(sol(tc)[1, 0].numpy(), tc)
        tm = (t_span[0] + t_span[1]) / 2
        assert sol(tm).shape == (3, 2) and float(sol(tm)[0, 1]) == tm
        tq = np.array([t_span[1], tm, t_span[0], tm, tc[7]])
        assert np.array_equal(sol(tq)[2, 1].numpy(), tq)
        assert np.array_equal(sol(torch.tensor(tq))[0, 0].numpy(), tq)
        with pytest.raises(ValueError):
            sol(max(t_span) + 1.0)
        with pytest.raises(ValueError):
            sol(np.zeros((2, 2)) + tm)
