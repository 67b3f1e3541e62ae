"""rk_fast (xsq_rk_fast.cuh, the ensemble hot path) against rk_persistent (the
generic kernel, selected with XSQ_NO_FAST=1): every output must be bit
identical -- the two kernels spend their instructions differently but perform
the same arithmetic, operation by operation."""
import os

import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from test_gpu_rk import lorenz_lanes, vdp_lanes, arenstorf_lanes, to_np

pytestmark = pytest.mark.gpu

GENERIC = [xb.Ts5, xb.CK5, xb.Me4, xb.Pr7, xb.Pr8, xb.Pr9]
PROBLEMS = {"lorenz63": (lorenz_lanes, (0.0, 6.0)),
            "vanderpol": (vdp_lanes, (0.0, 8.0)),
            "arenstorf": (arenstorf_lanes, (0.0, 12.0))}


def solve(rhs, span, y0, m, prm, fast, **kw):
    old = os.environ.get("XSQ_NO_FAST")
    os.environ["XSQ_NO_FAST"] = "0" if fast else "1"
    try:
        res = xb.solve_ivp_batched(rhs, span, y0, m, params=prm, **kw)
        r = to_np(res)
        r["stiff_flags"] = (res.stiff_flags.cpu().numpy() if res.stiff_flags is not None
                            else np.zeros(len(y0), np.int32))
    finally:
        if old is None:
            del os.environ["XSQ_NO_FAST"]
        else:
            os.environ["XSQ_NO_FAST"] = old
    return r


def same(a, b):
    for k in ("t_final", "y_final", "h_next", "n_accepted", "n_rejected", "nfev", "status",
              "stiff_flags"):
        x, y = a[k], b[k]
        if x.dtype.kind == "f":
            assert np.array_equal(x.view(np.uint64), y.view(np.uint64)), k
        else:
            assert np.array_equal(x, y), k


@pytest.mark.parametrize("m", GENERIC, ids=lambda m: m.__name__)
@pytest.mark.parametrize("prob", sorted(PROBLEMS))
@pytest.mark.parametrize("stiff", [0, 5000, 300])
def test_fast_kernel_is_bit_identical_to_generic_kernel(m, prob, stiff):
    lanes, span = PROBLEMS[prob]
    y0, prm = lanes(1500)
    kw = dict(rtol=1e-8, atol=1e-10, nfev_stiff_detect=stiff)
    a = solve(prob, span, y0, m, prm, True, **kw)
    b = solve(prob, span, y0, m, prm, False, **kw)
    assert a["n_accepted"].min() > 10
    same(a, b)


@pytest.mark.parametrize("kw", [
    dict(rtol=1e-3, atol=1e-6),
    dict(rtol=1e-12, atol=1e-14),
    dict(rtol=1e-6, atol=[1e-8, 1e-9, 1e-7]),
    dict(rtol=1e-8, atol=1e-10, first_step=1e-4),
    dict(rtol=1e-8, atol=1e-10, max_step=0.01),
    dict(rtol=1e-6, atol=1e-9, max_step=0.3, first_step=0.25),
], ids=lambda kw: "-".join(f"{k}{v}" for k, v in kw.items() if k not in ("atol",)))
def test_fast_kernel_options(kw):
    y0, prm = lorenz_lanes(700, seed=3)
    for m in (xb.Ts5, xb.Pr8):
        a = solve("lorenz63", (0.0, 3.0), y0, m, prm, True, **kw)
        b = solve("lorenz63", (0.0, 3.0), y0, m, prm, False, **kw)
        assert (a["status"] == 0).all()
        same(a, b)


def test_fast_kernel_backward_and_zero_span():
    y0, prm = lorenz_lanes(300, seed=5)
    # (backward in time the Lorenz system expands: keep the spans short)
    for span in ((0.4, 0.0), (1.0, 1.0), (-1.0, -1.3)):
        a = solve("lorenz63", span, y0, xb.Ts5, prm, True, rtol=1e-7, atol=1e-9)
        b = solve("lorenz63", span, y0, xb.Ts5, prm, False, rtol=1e-7, atol=1e-9)
        same(a, b)


def test_fast_kernel_failures_match():
    """TOO_SMALL_STEP (finite-time blow-up y' = y^2 in the Van der Pol slot is
    not available, so: absurd tolerance on a stiff lane) and overflow."""
    y0 = np.tile([2.0, 0.0], (64, 1))
    mu = np.full((64, 1), 1e9)                 # explicit methods crawl, rejected steps pile up
    for m in (xb.Ts5, xb.Pr9):
        a = solve("vanderpol", (0.0, 1e-3), y0, m, mu, True, rtol=1e-12, atol=1e-14)
        b = solve("vanderpol", (0.0, 1e-3), y0, m, mu, False, rtol=1e-12, atol=1e-14)
        same(a, b)
    y0 = np.tile([1e200, 1e200, 1e200], (64, 1))
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (64, 1))
    a = solve("lorenz63", (0.0, 1.0), y0, xb.Ts5, prm, True, rtol=1e-8, atol=1e-10)
    b = solve("lorenz63", (0.0, 1.0), y0, xb.Ts5, prm, False, rtol=1e-8, atol=1e-10)
    assert (a["status"] != 0).any()
    same(a, b)


def test_fast_kernel_probe_queue_overflow_path():
    """With a tiny probe queue the slots take over (both kernels): same nfev / flags."""
    y0, prm = lorenz_lanes(2000, seed=11)
    outs = []
    for q in ("0", "64", None):
        if q is None:
            os.environ.pop("XSQ_STIFF_QUEUE_RECORDS", None)
        else:
            os.environ["XSQ_STIFF_QUEUE_RECORDS"] = q
        try:
            outs.append(solve("lorenz63", (0.0, 8.0), y0, xb.Ts5, prm, True, rtol=1e-8,
                              atol=1e-10, nfev_stiff_detect=240))
        finally:
            os.environ.pop("XSQ_STIFF_QUEUE_RECORDS", None)
    ref = solve("lorenz63", (0.0, 8.0), y0, xb.Ts5, prm, False, rtol=1e-8, atol=1e-10,
                nfev_stiff_detect=240)
    for o in outs:
        same(o, ref)
