"""Adaptive runs of the CUDA kernels against the C oracle in DEVICE ARITHMETIC
(oracle/xsq_oracle.c with xsq_oracle_set_device_math(1): the kernel's seeded
reciprocal for err/scale and its table-driven log2/exp2 in the controller and
in h_start, restated operation by operation).  The bar is bit equality of
every output on 100 % of the lanes -- accepted / rejected / nfev counts, final
time and state, the next step size, the stiffness flags and the dense output --
for every built-in method and right-hand side, at BASELINE.json's shapes:

* C2: Lorenz-63 lanes of the benchmark ensemble itself, t in [0, 100];
* C3: Pr8 / Pr9 on the Van der Pol mu-sweep up to mu = 100 with 1000 t_eval points.

What the device arithmetic costs against the REFERENCE's arithmetic (pow, true
division) is measured separately, C oracle against C oracle and against the
NumPy restatement that is bit-identical to the reference
(test_cost_of_device_arithmetic*): the fractions are printed and asserted.
Reference: common.py:249-287 (controller), ivp.py:711-728 (t_eval slicing).
"""
import os
import sys

import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle import c_oracle as CO
from oracle import rk_oracle as O
from test_gpu_rk import lorenz_lanes, vdp_lanes, arenstorf_lanes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABS = O.load_tableaux()
METHODS = [xb.Ts5, xb.BS5, xb.CK5, xb.Me4, xb.Pr7, xb.Pr8, xb.Pr9, xb.CFMR7osc]
THREADS = max(1, len(os.sched_getaffinity(0)))
KEYS_F = ("t_final", "y_final", "h_next")
KEYS_I = ("n_accepted", "n_rejected", "nfev", "status")


def gpu(rhs, span, y0, m, prm, **kw):
    res = xb.solve_ivp_batched(rhs, span, y0, m, params=prm, **kw)
    torch.cuda.synchronize()
    out = {k: getattr(res, k).cpu().numpy() for k in KEYS_F + KEYS_I}
    out["stiff_flags"] = res.stiff_flags.cpu().numpy()
    out["y"] = res.y.cpu().numpy() if res.y is not None else None
    return out


def oracle(rhs, span, y0, m, prm, device_math=True, **kw):
    kw = dict(kw)
    kw.setdefault("n_threads", THREADS)
    if device_math:
        with CO.device_math():
            return CO.rk_batch(TABS[m.__name__], rhs, span, y0, params=prm, **kw)
    return CO.rk_batch(TABS[m.__name__], rhs, span, y0, params=prm, **kw)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_identical(g, o, what, dense=False):
    for k in KEYS_I + ("stiff_flags",):
        bad = np.flatnonzero(g[k] != o[k])
        assert bad.size == 0, (what, k, bad[:5], g[k][bad[:5]], o[k][bad[:5]])
    for k in KEYS_F:
        bad = np.flatnonzero((bits(g[k]) != bits(o[k])).reshape(len(g[k]), -1).any(axis=1))
        assert bad.size == 0, (what, k, bad[:5], g[k][bad[:5]], o[k][bad[:5]])
    if dense:
        a, b = g["y"], o["y"]
        same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (what, "y(t_eval)", np.argwhere(~same)[:5],
                            np.nanmax(np.abs(a - b)))


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_adaptive_run_is_bit_identical_to_oracle_in_device_arithmetic(m, prob):
    lanes, span = {"lorenz63": (lorenz_lanes, (0.0, 12.0)),
                   "vanderpol": (vdp_lanes, (0.0, 20.0)),
                   "arenstorf": (arenstorf_lanes, (0.0, 17.0652165601579625588917206249))}[prob]
    y0, prm = lanes(768)
    for stiff in (5000, 300):
        kw = dict(rtol=1e-8, atol=1e-10, nfev_stiff_detect=stiff)
        g = gpu(prob, span, y0, m, prm, **kw)
        o = oracle(prob, span, y0, m, prm, **kw)
        assert g["n_accepted"].min() > 20
        assert_identical(g, o, (m.__name__, prob, stiff))


@pytest.mark.parametrize("m", METHODS, ids=lambda m: m.__name__)
def test_dense_output_is_bit_identical_to_oracle_in_device_arithmetic(m):
    y0, prm = lorenz_lanes(384, seed=7)
    t_eval = np.linspace(0.0, 6.0, 257)
    kw = dict(rtol=1e-7, atol=1e-9, t_eval=t_eval)
    g = gpu("lorenz63", (0.0, 6.0), y0, m, prm, **kw)
    o = oracle("lorenz63", (0.0, 6.0), y0, m, prm, **kw)
    assert_identical(g, o, (m.__name__, "t_eval"), dense=True)


@pytest.mark.parametrize("kw", [
    dict(rtol=1e-3, atol=1e-6), dict(rtol=1e-11, atol=1e-13),
    dict(rtol=1e-6, atol=[1e-9, 1e-7, 1e-8]), dict(rtol=1e-8, atol=1e-10, max_step=0.02),
    dict(rtol=1e-8, atol=1e-10, first_step=1e-3),
    dict(rtol=1e-6, atol=1e-8, sc_params=(0.6, -0.2, 0.0, 0.9)),
    dict(rtol=1e-6, atol=1e-8, sc_params=(0.7, -0.4, 0.1, 0.8)),     # alpha term: generic kernel
], ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items() if k != "atol"))
def test_options_bit_identical(kw):
    y0, prm = lorenz_lanes(512, seed=21)
    for m in (xb.Ts5, xb.BS5, xb.Pr7):
        g = gpu("lorenz63", (0.0, 5.0), y0, m, prm, **kw)
        o = oracle("lorenz63", (0.0, 5.0), y0, m, prm, **kw)
        assert_identical(g, o, (m.__name__, kw))


# ---- CKdisc ------------------------------------------------------------------
@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_ckdisc_bit_identical_to_oracle_in_device_arithmetic(prob):
    """rk_persistent<CKdisc> (Lane::attempt_ckdisc) against ck_solve_one of the C
    oracle in device arithmetic: the variable order step of cash.py:245-416 --
    assessments, fallback solutions, twiddle / quit adaptation -- reproduced
    bit for bit on every lane, dense output through both interpolants included."""
    tab = O.load_ckdisc()
    lanes, span = {"lorenz63": (lorenz_lanes, (0.0, 12.0)),
                   "vanderpol": (vdp_lanes, (0.0, 20.0)),
                   "arenstorf": (arenstorf_lanes, (0.0, 17.0652165601579625588917206249))}[prob]
    y0, prm = lanes(768)
    for kw in (dict(rtol=1e-8, atol=1e-10), dict(rtol=1e-4, atol=1e-6),
               dict(rtol=1e-6, atol=1e-8, t_eval=np.linspace(span[0], span[1], 200))):
        g = gpu(prob, span, y0, xb.CKdisc, prm, **kw)
        with CO.device_math():
            o = CO.rk_batch(tab, prob, span, y0, params=prm, n_threads=THREADS, **kw)
        assert g["n_rejected"].sum() > 0
        assert_identical(g, o, ("CKdisc", prob, sorted(kw)), dense="t_eval" in kw)


def test_ckdisc_nonsmooth_ensemble_bit_identical_and_cost_of_arithmetic():
    """DETEST F2 (docs/Cash_Karp.ipynb; the right-hand side switches at every
    integer t, which is what CKdisc is for) as a user right-hand side: 96 lanes
    bit-identical to the C oracle in device arithmetic (a Python callback plays
    the right-hand side there).  How often the reference's own arithmetic takes
    the same step sequence is printed: that fraction is a property of the
    problem (steps land on the discontinuities), not of the kernel."""
    from oracle.problems import CUDA_SOURCES, make_fun
    tab = O.load_ckdisc()
    n, p, src = CUDA_SOURCES["detest_f2"]
    rhs = xb.DeviceRHS.from_source(src, "rhs", n, p)
    N = 96
    y0 = np.random.default_rng(7).uniform(60.0, 140.0, (N, 1))
    kw = dict(rtol=1e-6, atol=1e-8)
    g = gpu(rhs, (0.0, 6.0), y0, xb.CKdisc, None, max_steps=100000, **kw)
    fun = make_fun("detest_f2", [])
    with CO.device_math():
        o = CO.rk_batch(tab, None, (0.0, 6.0), y0, user_fn=fun, **kw)
    assert (g["status"] == 0).all() and g["n_rejected"].min() > 0
    assert_identical(g, o, ("CKdisc", "detest_f2"))
    r = CO.rk_batch(tab, None, (0.0, 6.0), y0, user_fn=fun, **kw)
    same = (r["nfev"] == g["nfev"]) & (r["n_rejected"] == g["n_rejected"])
    print(f"\nCKdisc DETEST F2: {same.sum()} of {N} lanes take the same step sequence in the "
          f"reference's arithmetic; nfev per lane {g['nfev'].mean():.1f} vs {r['nfev'].mean():.1f}")
    assert abs(g["nfev"].mean() / r["nfev"].mean() - 1) < 0.05


# ---- user tableau + user right-hand side ---------------------------------------------
def test_user_tableau_and_user_rhs_bit_identical_to_oracle():
    """docs/Demo_own_RK.ipynb: a user RungeKutta subclass (Heun-Euler pair, no P:
    cubic Hermite dense output) on a user right-hand side -- both compiled at run
    time -- against the C oracle in device arithmetic with the same tableau."""
    from oracle.problems import CUDA_SOURCES, mass_spring_damper

    class Heun(xb.RungeKutta):
        n_stages = 2
        order = 2
        order_secondary = 1
        C = np.array([0, 1.])
        A = np.array([[0, 0], [1., 0]])
        B = np.array([1 / 2, 1 / 2])
        E = np.array([1., 0, 0])
        E[:-1] -= B

    class HeunTab:
        name = "Heun"
        n_stages, order, order_secondary = 2, 2, 1
        A, B, C, E = Heun.A, Heun.B, Heun.C, Heun.E
        P = None
        sc_params = "standard"
    n, p, src = CUDA_SOURCES["mass_spring_damper"]
    msd = xb.DeviceRHS.from_source(src, "rhs", n, p)
    rng = np.random.default_rng(8)
    y0 = np.stack([rng.uniform(-1, 1, 64), rng.uniform(-2, 0, 64)], 1)
    te = np.linspace(0, 16, 33)
    for kw in (dict(atol=0.05), dict(rtol=1e-5, atol=1e-7), dict(rtol=1e-4, atol=1e-6, t_eval=te)):
        g = gpu(msd, (0.0, 16.0), y0, Heun, None, **kw)
        with CO.device_math():
            o = CO.rk_batch(HeunTab, None, (0.0, 16.0), y0, user_fn=mass_spring_damper, **kw)
        assert_identical(g, o, ("Heun", sorted(kw)), dense="t_eval" in kw)


# ---- events ------------------------------------------------------------------------
@pytest.mark.parametrize("m", [xb.Ts5, xb.BS5, xb.Pr8, xb.CKdisc], ids=lambda m: m.__name__)
def test_events_bit_identical_to_oracle_in_device_arithmetic(m):
    """scipy's `events=` on the device (event queue + event_queue_body, rk_fast
    with event hooks, the in-lane solver for terminal occurrences) against the
    independent C restatement of ivp.py's event handling in the kernels'
    arithmetic (oracle/xsq_oracle.c events_after_step / brentq_c, pinned to the
    reference's golden runs by tests/test_events_golden.py): event times and
    states, counts, terminal stops, the t_eval output cut at the event, and the
    trajectory outputs, bit for bit on every lane."""
    from oracle.problems import EVENT_SETS
    tab = O.load_ckdisc() if m is xb.CKdisc else TABS[m.__name__]
    _, src = EVENT_SETS["lorenz_sections"]
    ev0 = xb.DeviceEvents.from_source(src, "event", 3)
    y0, prm = lorenz_lanes(768, seed=31)
    te = np.linspace(0.0, 6.0, 61)
    cap = 16
    for term, kw in (([0, 0, 0], {}), ([0, 3, 0], {}), ([2, 0, 4], dict(nfev_stiff_detect=0)),
                     ([0, 0, 0], dict(t_eval=te)), ([3, 0, 0], dict(t_eval=te))):
        direc = [1, 0, -1]
        base = dict(rtol=1e-7, atol=1e-9, **kw)
        ev = ev0.with_attributes(terminal=term, direction=direc)
        res = xb.solve_ivp_batched("lorenz63", (0.0, 6.0), y0, m, params=prm, events=ev,
                                   max_event_records=cap, **base)
        torch.cuda.synchronize()
        g = {k: getattr(res, k).cpu().numpy() for k in KEYS_F + KEYS_I + ("t_events", "y_events",
                                                                          "event_counts")}
        g["stiff_flags"] = res.stiff_flags.cpu().numpy()
        g["y"] = res.y.cpu().numpy() if res.y is not None else None
        with CO.device_math():
            o = CO.rk_events_batch(tab, "lorenz63", (0.0, 6.0), y0, "lorenz_sections", term, direc,
                                   cap, params=prm, n_threads=THREADS, **base)
        assert g["event_counts"].sum() > 2 * len(y0)
        if any(term):
            assert (g["status"] == 1).sum() > len(y0) // 4
        assert_identical(g, o, (m.__name__, term, sorted(kw)), dense="t_eval" in kw)
        for k in ("t_events", "y_events", "event_counts"):
            a, b = g[k], o[k]
            same = (a == b) | ((a != a) & (b != b))
            assert same.all(), (m.__name__, term, sorted(kw), k, np.argwhere(~same)[:4])


def test_swag_events_bit_identical_to_oracle_in_device_arithmetic():
    """SWAG with events (roots on SwagDenseOutput inside the lane) against
    oracle/xsq_oracle_swag.c with events in device arithmetic: bit for bit."""
    from oracle.problems import EVENT_SETS
    _, src = EVENT_SETS["lorenz_sections"]
    ev0 = xb.DeviceEvents.from_source(src, "event", 3)
    y0, prm = lorenz_lanes(512, seed=33)
    te = np.linspace(0.0, 6.0, 61)
    for term, kw in (([0, 0, 0], {}), ([0, 3, 0], {}), ([3, 0, 0], dict(t_eval=te))):
        direc = [1, 0, -1]
        base = dict(rtol=1e-7, atol=1e-9, **kw)
        res = xb.solve_ivp_batched("lorenz63", (0.0, 6.0), y0, xb.SWAG, params=prm,
                                   events=ev0.with_attributes(terminal=term, direction=direc),
                                   max_event_records=16, **base)
        torch.cuda.synchronize()
        with CO.device_math():
            o = CO.swag_events_batch("lorenz63", (0.0, 6.0), y0, "lorenz_sections", term, direc, 16,
                                     params=prm, n_threads=THREADS, **base)
        if any(term):
            assert (o["status"] == 1).sum() > len(y0) // 4
        for k in ("n_accepted", "n_rejected", "nfev", "status", "event_counts", "t_final", "y_final",
                  "t_events", "y_events") + (("y",) if "t_eval" in kw else ()):
            a, b = getattr(res, k).cpu().numpy(), o[k]
            same = (a == b) | ((a != a) & (b != b))
            assert same.all(), ("SWAG", term, sorted(kw), k, np.argwhere(~same)[:4])


# ---- C4: the perturbed Arenstorf ensemble of bench.py ----------------------------
def test_c4_collision_orbits():
    """The first 16 384 lanes of bench.py's C4 ensemble (Arenstorf initial state
    +- 1e-3, one period) with Pr8.  About one lane in a thousand hits the Moon's
    singularity: the step size falls below the spacing of t and the reference
    gives up with 'required step size is less than spacing between numbers'
    (common.py:233-234) -- status -1 here.  Same lanes, same time of failure,
    same bits as the C oracle in device arithmetic."""
    N = 16384
    rng = np.random.default_rng(2024)
    y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) + \
        rng.uniform(-1e-3, 1e-3, (1_000_000, 4))[:N]
    prm = np.full((N, 1), 0.012277471)
    span = (0.0, 17.0652165601579625588917206249)
    kw = dict(rtol=1e-8, atol=1e-10)
    g = gpu("arenstorf", span, y0, xb.Pr8, prm, **kw)
    o = oracle("arenstorf", span, y0, xb.Pr8, prm, **kw)
    assert_identical(g, o, ("Pr8", "C4"))
    failed = g["status"] != 0
    print(f"\nC4 Pr8: {failed.sum()} of {N} lanes end with status -1 (collision orbits)")
    assert (g["status"][failed] == -1).all() and 0 < failed.sum() < 0.01 * N
    # every failed lane is at the Moon: x = 1 - mu, y = 0
    assert np.abs(g["y_final"][failed, 0] - (1 - 0.012277471)).max() < 1e-6


# ---- C3 at its shape ---------------------------------------------------------
@pytest.mark.parametrize("m", [xb.Pr8, xb.Pr9], ids=lambda m: m.__name__)
def test_c3_vanderpol_mu_sweep_1000_points(m):
    """BASELINE.json configs[2]: mu log-uniform on [0.1, 100], y0 = (2, 0),
    t in [0, 20], 1000 t_eval points, rtol 1e-8, atol 1e-10 -- 4096 lanes spread
    over the whole 1 M-lane sweep (every 244th lane + both ends)."""
    N_FULL, N = 1_000_000, 4096
    idx = np.unique(np.round(np.linspace(0, N_FULL - 1, N)).astype(np.int64))
    mu = 10.0 ** (-1 + 3 * idx / (N_FULL - 1))
    assert mu[0] == 0.1 and abs(mu[-1] - 100.0) < 1e-12
    y0 = np.tile([2.0, 0.0], (len(mu), 1))
    t_eval = np.linspace(0.0, 20.0, 1000)
    kw = dict(rtol=1e-8, atol=1e-10, t_eval=t_eval)
    g = gpu("vanderpol", (0.0, 20.0), y0, m, mu[:, None], **kw)
    o = oracle("vanderpol", (0.0, 20.0), y0, m, mu[:, None], **kw)
    assert (g["status"] == 0).all()
    assert_identical(g, o, (m.__name__, "C3"), dense=True)
    # the stiff end really is in the sample: rejected steps pile up there
    assert g["n_rejected"][-64:].mean() > 5 * g["n_rejected"][:64].mean()
    # and against the reference's arithmetic (pow, true division): solutions at
    # t_eval within 10 x rtol (BASELINE.json north_star), counts reported
    r = oracle("vanderpol", (0.0, 20.0), y0, m, mu[:, None], device_math=False, **kw)
    scale = 1e-10 + 1e-8 * np.abs(r["y"])
    worst = np.max(np.abs(g["y"] - r["y"]) / scale)
    same = (g["n_accepted"] == r["n_accepted"]) & (g["n_rejected"] == r["n_rejected"])
    print(f"\nC3 {m.__name__}: counts identical to reference arithmetic on "
          f"{same.mean():.4f} of {len(mu)} lanes; total attempts "
          f"{(g['n_accepted'] + g['n_rejected']).sum()} vs {(r['n_accepted'] + r['n_rejected']).sum()}; "
          f"max |dy| / (atol + rtol |y|) at t_eval = {worst:.2f}")
    tot_g = float((g["n_accepted"] + g["n_rejected"]).sum())
    tot_r = float((r["n_accepted"] + r["n_rejected"]).sum())
    assert abs(tot_g - tot_r) / tot_r < 1e-3


# ---- C2 at its shape ---------------------------------------------------------
def test_c2_bench_ensemble_t100_exact_and_distribution():
    """The first 10 240 lanes of the benchmark's own shard (bench.make_lanes,
    rank 0), t in [0, 100], Ts5 and CK5, reference defaults: bit-identical to
    the oracle in device arithmetic, lane by lane, at the benchmark's horizon
    (chaos amplifies any difference, so this is the strongest statement there
    is); and against the reference's arithmetic the mean accepted / rejected
    steps per lane agree within 0.1 % (hard part 1c)."""
    sys.path.insert(0, ROOT)
    import bench
    N = 10240
    y0_full, prm_full = bench.make_lanes(bench.LANES_PER_GPU, 0)
    y0, prm = y0_full[:N], prm_full[:N]
    y0_n, prm_n = bench.make_lanes(N, 0)
    assert np.array_equal(y0, y0_n) and np.array_equal(prm, prm_n)   # a true subset
    for m in (xb.Ts5, xb.CK5):
        kw = dict(rtol=bench.RTOL, atol=bench.ATOL)
        g = gpu("lorenz63", (0.0, bench.T_END), y0, m, prm, **kw)
        o = oracle("lorenz63", (0.0, bench.T_END), y0, m, prm, **kw)
        assert (g["status"] == 0).all()
        assert_identical(g, o, (m.__name__, "C2 T=100"))
        r = oracle("lorenz63", (0.0, bench.T_END), y0, m, prm, device_math=False, **kw)
        da = g["n_accepted"].mean() / r["n_accepted"].mean() - 1
        dr = g["n_rejected"].mean() / r["n_rejected"].mean() - 1
        # at T = 100 the two arithmetics follow different (equally valid) chaotic
        # trajectories, so the lane means are independent samples: their
        # difference has standard error sqrt((var_g + var_r) / N)
        se_a = np.sqrt((g["n_accepted"].var() + r["n_accepted"].var()) / N) / r["n_accepted"].mean()
        se_r = np.sqrt((g["n_rejected"].var() + r["n_rejected"].var()) / N) / r["n_rejected"].mean()
        print(f"\nC2 {m.__name__} T=100, {N} lanes: accepted/lane {g['n_accepted'].mean():.2f} "
              f"(reference arithmetic {r['n_accepted'].mean():.2f}, {da:+.2e}, s.e. {se_a:.1e}), "
              f"rejected/lane {g['n_rejected'].mean():.2f} ({r['n_rejected'].mean():.2f}, "
              f"{dr:+.2e}, s.e. {se_r:.1e})")
        assert abs(da) < max(1e-3, 4 * se_a) and abs(dr) < max(1e-3, 4 * se_r)


def test_cost_of_device_arithmetic_vs_numpy_restatement():
    """C oracle in device arithmetic against the NumPy restatement that is
    bit-identical to the reference: fraction of lanes with identical
    accepted / rejected counts inside the predictability horizon."""
    y0, prm = lorenz_lanes(96, seed=99)
    tab = TABS["Ts5"]
    with CO.device_math():
        d = CO.rk_batch(tab, "lorenz63", (0.0, 10.0), y0, params=prm, rtol=1e-8, atol=1e-10,
                        n_threads=THREADS)
    same = 0
    for i in range(len(y0)):
        r = O.rk_solve(tab, O.lorenz63(*prm[i]), (0.0, 10.0), y0[i], rtol=1e-8, atol=1e-10)
        same += int(r["n_accepted"] == d["n_accepted"][i] and r["n_rejected"] == d["n_rejected"][i]
                    and r["nfev"] == d["nfev"][i])
    print(f"\ndevice arithmetic vs reference (NumPy restatement), Lorenz T=10: "
          f"{same}/{len(y0)} lanes with identical accepted/rejected/nfev")
    assert same >= 0.99 * len(y0)


# ---- SWAG ------------------------------------------------------------------------
@pytest.mark.parametrize("prob", ["lorenz63", "vanderpol", "arenstorf"])
def test_swag_bit_identical_to_oracle_in_device_arithmetic(prob):
    """swag_persistent against oracle/xsq_oracle_swag.c in device arithmetic
    (h_start's tolerance power and the step-reduction power through the kernel's
    own log2 / exp2): accepted / failed / nfev counts, final states and the
    dense output equal bit for bit on every lane (shampine.py:180-480)."""
    lanes, span = {"lorenz63": (lorenz_lanes, (0.0, 10.0)),
                   "vanderpol": (vdp_lanes, (0.0, 20.0)),
                   "arenstorf": (arenstorf_lanes, (0.0, 17.0652165601579625588917206249))}[prob]
    y0, prm = lanes(640)
    for te in (None, np.linspace(span[0], span[1], 200)):
        kw = dict(rtol=1e-8, atol=1e-10, t_eval=te)
        res = xb.solve_ivp_batched(prob, span, y0, xb.SWAG, params=prm, **kw)
        torch.cuda.synchronize()
        with CO.device_math():
            o = CO.swag_batch(prob, span, y0, params=prm, n_threads=THREADS, **kw)
        for k in ("n_accepted", "n_rejected", "nfev", "status"):
            g = getattr(res, k).cpu().numpy()
            bad = np.flatnonzero(g != o[k])
            assert bad.size == 0, (prob, k, bad[:5], g[bad[:5]], o[k][bad[:5]])
        for k in ("t_final", "y_final"):
            g = getattr(res, k).cpu().numpy()
            assert np.array_equal(bits(g), bits(o[k])), (prob, k)
        if te is not None:
            g = res.y.cpu().numpy()
            ok = (bits(g) == bits(o["y"])) | (np.isnan(g) & np.isnan(o["y"]))
            assert ok.all(), (prob, "y(t_eval)", np.argwhere(~ok)[:4])
