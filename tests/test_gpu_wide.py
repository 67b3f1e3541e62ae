"""Warp-per-system kernels for USER right-hand sides with 16 < n <= 1024
(the reference takes any n, common.py:187-217; north star: "one warp per
medium system").  The user's device function returns one component of the
derivative; component i lives in slot i / 32 of lane i % 32
(extensisq_b200/csrc/xsq_rhs.cuh, WideSystem).

Problem: the 1-D Brusselator with diffusion on N = n/2 grid points,
y = [u_0, v_0, u_1, v_1, ...], Dirichlet boundary u = 1, v = 3.

Oracle: the C restatement (oracle/xsq_oracle.c) with the right-hand side as a
NumPy callback that evaluates the same expression tree; for the adaptive runs
in the kernels' own arithmetic AND summation order (device_math with
warp_strided), so the comparison is exact."""
import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle import c_oracle as CO
from oracle import rk_oracle as O

pytestmark = pytest.mark.gpu

TABS = O.load_tableaux()

SRC = """
#define NS %d
__device__ double brus(int i, double t, const double* y, const double* p) {
    const int j = i >> 1;
    const double u = y[2 * j], v = y[2 * j + 1], c = p[0], A = p[1], B = p[2];
    const double uuv = u * u * v;
    if (i & 1) {
        const double vl = j > 0 ? y[i - 2] : 3.0, vr = j < NS / 2 - 1 ? y[i + 2] : 3.0;
        return B * u - uuv + c * (vl - 2.0 * v + vr);
    }
    const double ul = j > 0 ? y[i - 2] : 1.0, ur = j < NS / 2 - 1 ? y[i + 2] : 1.0;
    return A + uuv - (B + 1.0) * u + c * (ul - 2.0 * u + ur);
}
"""
_RHS = {}


def rhs_for(n):
    if n not in _RHS:
        _RHS[n] = xb.DeviceRHS.from_source(SRC % n, "brus", n, 3)
    return _RHS[n]


def brus_numpy(t, y, p):
    """Same operations in the same order as the device function."""
    c, A, B = p[0], p[1], p[2]
    u, v = y[0::2], y[1::2]
    ul = np.concatenate(([1.0], u[:-1]))
    ur = np.concatenate((u[1:], [1.0]))
    vl = np.concatenate(([3.0], v[:-1]))
    vr = np.concatenate((v[1:], [3.0]))
    uuv = u * u * v
    out = np.empty_like(y)
    out[0::2] = A + uuv - (B + 1.0) * u + c * (ul - 2.0 * u + ur)
    out[1::2] = B * u - uuv + c * (vl - 2.0 * v + vr)
    return out


def lanes(n, N):
    ng = n // 2
    x = (np.arange(ng) + 1.0) / (ng + 1.0)
    y0 = np.empty((N, n))
    for s in range(N):
        y0[s, 0::2] = 1.0 + np.sin(2.0 * np.pi * x) * (1.0 + 0.05 * s)
        y0[s, 1::2] = 3.0
    c = 0.02 * (ng + 1.0) ** 2 * (1.0 + 0.1 * np.arange(N) / max(N - 1, 1))
    prm = np.stack([c, np.ones(N), np.full(N, 3.0)], axis=1)
    return y0, prm


def to_np(res):
    torch.cuda.synchronize()
    return {k: (getattr(res, k).cpu().numpy() if getattr(res, k) is not None else None)
            for k in ("y", "t_final", "y_final", "n_accepted", "n_rejected", "nfev",
                      "status", "stiff_flags")}


@pytest.mark.parametrize("n,m", [(64, xb.Ts5), (64, xb.Pr8), (100, xb.Ts5), (100, xb.BS5),
                                 (256, xb.CK5), (1024, xb.Ts5)],
                         ids=lambda v: getattr(v, "__name__", str(v)))
def test_forced_steps_bit_exact(n, m):
    N = 6
    y0, prm = lanes(n, N)
    hmax = 0.5 / prm[:, 0].max()                 # inside the stability region
    hs = hmax * (0.6 + 0.4 * np.sin(0.37 * np.arange(40)) ** 2)
    span = (0.0, 1.0)
    res = to_np(xb.solve_ivp_batched(rhs_for(n), span, y0, m, params=prm, forced_steps=hs))
    ref = CO.rk_batch(TABS[m.__name__], None, span, y0, params=prm, forced_h=hs,
                      user_fn=brus_numpy, user_fn_params=True)
    assert np.array_equal(res["n_accepted"], ref["n_accepted"])
    assert np.array_equal(res["nfev"], ref["nfev"])
    assert np.array_equal(res["t_final"], ref["t_final"])
    assert np.array_equal(res["y_final"], ref["y_final"])          # bit-exact


@pytest.mark.parametrize("n,m,T", [(64, xb.Ts5, 2.0), (64, xb.BS5, 2.0), (100, xb.Ts5, 1.0),
                                   (100, xb.Me4, 0.5), (256, xb.CK5, 0.3), (1024, xb.Ts5, 0.02)],
                         ids=lambda v: getattr(v, "__name__", str(v)))
def test_adaptive_identical_to_device_math_oracle(n, m, T):
    """Accepted / rejected / nfev per system and the final states, exactly."""
    N = 5
    y0, prm = lanes(n, N)
    kw = dict(rtol=1e-6, atol=1e-8)
    res = to_np(xb.solve_ivp_batched(rhs_for(n), (0.0, T), y0, m, params=prm, **kw))
    with CO.device_math(warp_strided=True):
        ref = CO.rk_batch(TABS[m.__name__], None, (0.0, T), y0, params=prm,
                          user_fn=brus_numpy, user_fn_params=True, **kw)
    assert (res["status"] == 0).all() and (ref["status"] == 0).all()
    for k in ("n_accepted", "n_rejected", "nfev"):
        assert np.array_equal(res[k], ref[k]), (k, res[k], ref[k])
    assert np.array_equal(res["y_final"], ref["y_final"])
    assert res["n_accepted"].min() > 20


def test_stiffness_probes_of_a_wide_system_match_the_oracle():
    """The probe runs inside the persistent kernel for warp-per-system right-hand
    sides (two slots per thread); padded slots (n = 100) must stay out of it."""
    n, N = 100, 4
    y0, prm = lanes(n, N)
    kw = dict(rtol=1e-5, atol=1e-7, nfev_stiff_detect=300)
    res = to_np(xb.solve_ivp_batched(rhs_for(n), (0.0, 3.0), y0, xb.Ts5, params=prm, **kw))
    with CO.device_math(warp_strided=True):
        ref = CO.rk_batch(TABS["Ts5"], None, (0.0, 3.0), y0, params=prm,
                          user_fn=brus_numpy, user_fn_params=True, **kw)
    for k in ("n_accepted", "n_rejected", "nfev", "stiff_flags"):
        assert np.array_equal(res[k], ref[k]), (k, res[k], ref[k])
    assert np.array_equal(res["y_final"], ref["y_final"])
    assert (res["nfev"] > 6 * (res["n_accepted"] + res["n_rejected"]) + 4).all()   # probes ran
    assert (res["stiff_flags"] != 0).any()       # diffusion limited: diagnosed stiff


def test_dense_output_of_a_wide_system():
    n, N = 100, 3
    y0, prm = lanes(n, N)
    te = np.linspace(0.0, 1.0, 23)
    kw = dict(rtol=1e-6, atol=1e-8, t_eval=te)
    res = to_np(xb.solve_ivp_batched(rhs_for(n), (0.0, 1.0), y0, xb.Ts5, params=prm, **kw))
    with CO.device_math(warp_strided=True):
        ref = CO.rk_batch(TABS["Ts5"], None, (0.0, 1.0), y0, params=prm,
                          user_fn=brus_numpy, user_fn_params=True, **kw)
    assert np.array_equal(res["n_accepted"], ref["n_accepted"])
    assert res["y"].shape == (N, n, te.size)
    assert np.abs(res["y"] - ref["y"]).max() <= 1e-13
    assert np.array_equal(res["y"][:, :, 0], y0)


def test_swag_runs_a_wide_system():
    n, N = 64, 4
    y0, prm = lanes(n, N)
    kw = dict(rtol=1e-6, atol=1e-8)
    res = to_np(xb.solve_ivp_batched(rhs_for(n), (0.0, 1.0), y0, xb.SWAG, params=prm, **kw))
    ref = to_np(xb.solve_ivp_batched(rhs_for(n), (0.0, 1.0), y0, xb.Pr8, params=prm,
                                     rtol=1e-10, atol=1e-12))
    assert (res["status"] == 0).all()
    assert np.abs(res["y_final"] - ref["y_final"]).max() <= 2e-5


def test_limits_of_the_wide_path():
    with pytest.raises(Exception):
        xb.DeviceRHS.from_source(SRC % 2048, "brus", 2048, 3)
    ev = xb.DeviceEvents.from_source(
        "__device__ double ev(int k, double t, const double* y, const double* p) { return y[0] - 1.5; }",
        "ev", 1, terminal=[False], direction=[0])
    y0, prm = lanes(64, 2)
    with pytest.raises(Exception):
        xb.solve_ivp_batched(rhs_for(64), (0.0, 1.0), y0, xb.Ts5, params=prm, events=ev)
