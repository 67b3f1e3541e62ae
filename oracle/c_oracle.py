"""TEST INFRASTRUCTURE -- ctypes wrapper of ``oracle/libxsq_oracle.so`` (the
plain-C restatement in ``oracle/xsq_oracle.c``).  Not part of the product."""
import ctypes as C
import os

import numpy as np

from . import rk_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxsq_oracle.so")
MAXS, MAXPOL = 18, 8

RHS_IDS = {"lorenz63": 0, "vanderpol": 1, "arenstorf": 2, "nbody32": 3}
RHS_FN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double),
                     C.POINTER(C.c_double), C.POINTER(C.c_double))


class OTab(C.Structure):
    _fields_ = [
        ("s", C.c_int32), ("order", C.c_int32), ("order2", C.c_int32),
        ("npol", C.c_int32), ("variant", C.c_int32), ("pad", C.c_int32),
        ("A", (C.c_double * MAXS) * MAXS), ("B", C.c_double * MAXS),
        ("C", C.c_double * MAXS), ("E", C.c_double * (MAXS + 1)),
        ("P", (C.c_double * MAXPOL) * (MAXS + 1)),
        ("E_pre", C.c_double * MAXS), ("B_scale_pre", C.c_double * MAXS),
        ("C_extra", C.c_double * 3), ("A_extra", (C.c_double * MAXS) * 3),
        ("Plow", (C.c_double * MAXPOL) * (MAXS + 2)),
        ("Pbest", (C.c_double * MAXPOL) * (MAXS + 4)),
        ("npol_low", C.c_int32), ("npol_best", C.c_int32),
        ("sc", C.c_double * 4),
        ("stbrad", C.c_double), ("tanang", C.c_double),
        ("B_assess", (C.c_double * MAXS) * 2), ("E_assess", (C.c_double * MAXS) * 2),
        ("B_fallback", (C.c_double * MAXS) * 2), ("E_fallback", (C.c_double * MAXS) * 2),
        ("C_fallback", C.c_double * 2), ("ck_max_factor", C.c_double),
        ("ck_min_factor", C.c_double), ("ck_safety", C.c_double),
        ("Ap", (C.c_double * MAXS) * MAXS), ("Bp", C.c_double * MAXS),
        ("Ep", C.c_double * (MAXS + 1)),
    ]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle`")
        _lib = C.CDLL(LIB_PATH)
        _lib.xsq_oracle_tab_size.restype = C.c_size_t
        assert _lib.xsq_oracle_tab_size() == C.sizeof(OTab)
    return _lib


def max_threads():
    return load().xsq_oracle_max_threads()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


_rcp_bits = None


def set_device_math(on):
    """Switch the C oracle to the kernels' own arithmetic (seeded reciprocal
    for err/scale, table-driven log2/exp2 in the controller and in h_start):
    an adaptive device run is then reproduced BIT FOR BIT.  The reciprocal
    seed of MUFU.RCP64H is a table dumped on a B200 (rcp64h_delta.bin.z)."""
    global _rcp_bits
    lib = load()
    if on and _rcp_bits is None:
        import zlib
        raw = zlib.decompress(open(os.path.join(_HERE, "rcp64h_delta.bin.z"), "rb").read())
        _rcp_bits = np.frombuffer(raw, dtype=np.uint8).copy()
        assert _rcp_bits.size == (1 << 20) // 8
        lib.xsq_oracle_set_rcp_table(_rcp_bits.ctypes.data_as(C.c_void_p))
    rc = lib.xsq_oracle_set_device_math(1 if on else 0)
    assert rc == 0


def devmath(fn, x):
    """Element-wise restated device function: 'rcp_scale', 'log2', 'exp2', 'rcp64h'."""
    lib = load()
    set_device_math(True)
    set_device_math(False)            # the table stays loaded
    x = np.ascontiguousarray(x, dtype=float)
    out = np.empty_like(x)
    rc = lib.xsq_oracle_devmath({"rcp_scale": 0, "log2": 1, "exp2": 2, "rcp64h": 3}[fn],
                                _dp(x), _dp(out), C.c_int64(x.size))
    assert rc == 0
    return out


def set_warp_strided(on):
    """Sum over the components in the order of a warp-per-system kernel with
    component c in lane c % 32 (xsq_rhs.cuh WideSystem).  Returns the old mode."""
    return load().xsq_oracle_set_warp_strided(1 if on else 0)


class device_math:
    """with c_oracle.device_math(): ...   (restores the reference arithmetic)"""

    def __init__(self, warp_strided=False):
        self.warp_strided = warp_strided

    def __enter__(self):
        set_device_math(True)
        if self.warp_strided:
            set_warp_strided(True)

    def __exit__(self, *exc):
        set_device_math(False)
        set_warp_strided(False)


def _fill2(dst, src):
    src = np.asarray(src)
    for i in range(src.shape[0]):
        for j in range(src.shape[1]):
            dst[i][j] = float(src[i, j])


def make_tab(tab, sc_params=None):
    """tab: oracle.rk_oracle.Tableau or any object with the reference's class
    attributes (n_stages, order, order_secondary, A, B, C, E[, P])."""
    t = OTab()
    s = tab.n_stages
    t.s, t.order, t.order2 = s, tab.order, tab.order_secondary
    name = getattr(tab, "name", getattr(tab, "__name__", ""))
    t.variant = {"BS5": 1, "CFMR7osc": 2, "CKdisc": 3}.get(name, 0)
    _fill2(t.A, tab.A)
    for i in range(s):
        t.B[i] = float(tab.B[i])
        t.C[i] = float(tab.C[i])
    for i in range(s + 1):
        t.E[i] = float(tab.E[i])
    P = getattr(tab, "P", None)
    if isinstance(P, np.ndarray):
        t.npol = P.shape[1]
        _fill2(t.P, P)
    else:
        t.npol = 0
    if t.variant == 1:
        for i in range(6):
            t.E_pre[i] = float(tab.E_pre[i])
            t.B_scale_pre[i] = float(tab.B_scale_pre[i])
        for i in range(3):
            t.C_extra[i] = float(tab.C_extra[i])
        _fill2(t.A_extra, np.asarray(tab.A_extra))
        _fill2(t.Plow, tab.Plow)
        _fill2(t.Pbest, tab.Pbest)
        t.npol_low, t.npol_best = tab.Plow.shape[1], tab.Pbest.shape[1]
    if t.variant == 3:                      # cash.py:184-236
        for k in ("B_assess", "E_assess", "B_fallback", "E_fallback"):
            _fill2(getattr(t, k), np.asarray(getattr(tab, k)))
        for i in range(2):
            t.C_fallback[i] = float(tab.C_fallback[i])
        t.ck_max_factor = float(tab.max_factor)
        t.ck_min_factor = float(tab.min_factor)
        t.ck_safety = float(tab.safety)
    if getattr(tab, "variant", "") == "nystrom":     # common.py:1207-1309
        t.variant = 4
        _fill2(t.Ap, tab.Ap)
        for i in range(s):
            t.Bp[i] = float(tab.Bp[i])
        for i in range(s + 1):
            t.Ep[i] = float(tab.Ep[i])
    scp = sc_params if sc_params is not None else tab.sc_params
    if isinstance(scp, str):
        scp = O.SC_PRESETS[scp]
    for i in range(4):
        t.sc[i] = float(scp[i])
    sb, ta = getattr(tab, "stbrad", None), getattr(tab, "tanang", None)
    t.stbrad = float(sb) if isinstance(sb, (int, float)) else 0.0
    t.tanang = float(ta) if isinstance(ta, (int, float)) else 0.0
    return t


def rk_batch(tab, rhs, t_span, y0, params=None, rtol=1e-3, atol=1e-6,
             first_step=None, max_step=np.inf, sc_params=None,
             interpolant=None, t_eval=None, forced_h=None, max_steps=0,
             n_threads=1, user_fn=None, n_param=None, nfev_stiff_detect=5000,
             user_fn_params=False):
    """Integrate N lanes with the C oracle.  y0 [N, n], params [N, p].
    `rhs`: built-in name, or None with `user_fn` = python callable
    f(t, y) -> dy (slow; single thread).  Returns a dict of numpy arrays."""
    lib = load()
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=float)))
    N, n = y0.shape
    rtol_v, atol_v = O.validate_tol(float(rtol), atol, y0[0])
    atol_v = np.ascontiguousarray(np.broadcast_to(atol_v, (n,)), dtype=float)
    if params is not None:
        params = np.ascontiguousarray(np.asarray(params, dtype=float))
        if params.ndim == 1:
            params = params.reshape(N, -1)
        p = params.shape[1]
    else:
        p = 0
    t = make_tab(tab, sc_params)
    te = (np.ascontiguousarray(np.asarray(t_eval, dtype=float))
          if t_eval is not None else None)
    n_eval = te.size if te is not None else 0
    hf = (np.ascontiguousarray(np.asarray(forced_h, dtype=float))
          if forced_h is not None else None)
    y_eval = np.empty((N, n, n_eval)) if n_eval else None
    t_final = np.empty(N)
    y_final = np.empty((N, n))
    h_next = np.empty(N)
    n_acc = np.empty(N, np.int32)
    n_rej = np.empty(N, np.int32)
    nfev = np.empty(N, np.int32)
    status = np.empty(N, np.int32)
    n_done = np.empty(N, np.int32)
    stiff = np.zeros(N, np.int32)
    if rhs is not None:
        rid, cb = RHS_IDS[rhs], RHS_FN()
    else:
        rid = -1

        def _cb(tt, yp, pp, dyp):
            yv = np.ctypeslib.as_array(yp, (n,))
            if user_fn_params:
                out = user_fn(tt, yv, np.ctypeslib.as_array(pp, (p,)))
            else:
                out = user_fn(tt, yv)
            np.ctypeslib.as_array(dyp, (n,))[:] = out
        cb = RHS_FN(_cb)
    ip = {None: 0, "free": 1, "low": 2, "best": 3}[interpolant]
    rc = lib.xsq_oracle_rk_batch(
        C.byref(t), C.c_int(rid), cb, C.c_int(n), C.c_int(p), C.c_int64(N),
        _dp(y0), _dp(params), C.c_double(t_span[0]), C.c_double(t_span[1]),
        C.c_double(float(rtol_v)), _dp(atol_v),
        C.c_double(first_step if first_step is not None else 0.0),
        C.c_double(max_step), t.sc, C.c_int(ip), _dp(te), C.c_int(n_eval),
        _dp(y_eval), _dp(hf), C.c_int(hf.size if hf is not None else 0),
        C.c_int(max_steps), _dp(t_final), _dp(y_final), _dp(h_next),
        _ip(n_acc), _ip(n_rej), _ip(nfev), _ip(status), _ip(n_done),
        C.c_int(n_threads), C.c_int(nfev_stiff_detect), _ip(stiff))
    if rc != 0:
        raise RuntimeError("xsq_oracle_rk_batch failed")
    return dict(t=te, y=y_eval, t_final=t_final, y_final=y_final,
                h_next=h_next, n_accepted=n_acc, n_rejected=n_rej, nfev=nfev,
                status=status, n_eval_done=n_done, stiff_flags=stiff)


EVENT_FN = C.CFUNCTYPE(C.c_double, C.c_int, C.c_double, C.POINTER(C.c_double),
                       C.POINTER(C.c_double))
EVENT_SETS = {"lorenz_sections": (0, 3)}       # built into xsq_oracle.c: id, number of functions


def rk_events_batch(tab, rhs, t_span, y0, events, terminal, direction, capacity, params=None,
                    rtol=1e-3, atol=1e-6, first_step=None, max_step=np.inf, sc_params=None,
                    interpolant=None, t_eval=None, max_steps=0, n_threads=1,
                    nfev_stiff_detect=5000, user_fn=None):
    """rk_batch with scipy's `events=` in the kernels' arithmetic (xsq_oracle.c
    events_after_step / brentq_c; call inside ``with device_math():``).
    `events`: the name of a built-in set (EVENT_SETS) or a list of Python callables
    g(t, y) (single thread); `rhs=None` with `user_fn(t, y)` as in rk_batch.  Returns rk_batch's dict plus t_events
    [N, n_events, capacity], y_events [N, n_events, capacity, n] (NaN where unused)
    and event_counts [N, n_events]."""
    lib = load()
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=float)))
    N, n = y0.shape
    rtol_v, atol_v = O.validate_tol(float(rtol), atol, y0[0])
    atol_v = np.ascontiguousarray(np.broadcast_to(atol_v, (n,)), dtype=float)
    if params is not None:
        params = np.ascontiguousarray(np.asarray(params, dtype=float))
        if params.ndim == 1:
            params = params.reshape(N, -1)
        p = params.shape[1]
    else:
        p = 0
    t = make_tab(tab, sc_params)
    te = (np.ascontiguousarray(np.asarray(t_eval, dtype=float)) if t_eval is not None else None)
    n_eval = te.size if te is not None else 0
    y_eval = np.empty((N, n, n_eval)) if n_eval else None
    t_final, y_final, h_next = np.empty(N), np.empty((N, n)), np.empty(N)
    ints = [np.empty(N, np.int32) for _ in range(5)]
    n_acc, n_rej, nfev, status, n_done = ints
    stiff = np.zeros(N, np.int32)
    if isinstance(events, str):
        ev_set, ne = EVENT_SETS[events]
        gcb = EVENT_FN()
    else:
        ev_set, ne = -1, len(events)

        def _g(k, tt, yp, pp):
            return float(events[k](tt, np.ctypeslib.as_array(yp, (n,))))
        gcb = EVENT_FN(_g)
    if rhs is not None:
        rid, fcb = RHS_IDS[rhs], RHS_FN()
    else:
        rid = -1

        def _f(tt, yp, pp, dyp):
            np.ctypeslib.as_array(dyp, (n,))[:] = user_fn(tt, np.ctypeslib.as_array(yp, (n,)))
        fcb = RHS_FN(_f)
    term = np.ascontiguousarray(np.asarray(terminal, dtype=np.int32))
    direc = np.ascontiguousarray(np.asarray(direction, dtype=np.int32))
    assert term.size == ne and direc.size == ne
    t_ev = np.full((N, ne, capacity), np.nan)
    y_ev = np.full((N, ne, capacity, n), np.nan)
    ev_cnt = np.zeros((N, ne), np.int32)
    ip = {None: 0, "free": 1, "low": 2, "best": 3}[interpolant]
    rc = lib.xsq_oracle_rk_events_batch(
        C.byref(t), C.c_int(rid), fcb, C.c_int(n), C.c_int(p), C.c_int64(N),
        _dp(y0), _dp(params), C.c_double(t_span[0]), C.c_double(t_span[1]),
        C.c_double(float(rtol_v)), _dp(atol_v),
        C.c_double(first_step if first_step is not None else 0.0),
        C.c_double(max_step), t.sc, C.c_int(ip), _dp(te), C.c_int(n_eval),
        _dp(y_eval), C.c_int(max_steps), _dp(t_final), _dp(y_final), _dp(h_next),
        _ip(n_acc), _ip(n_rej), _ip(nfev), _ip(status), _ip(n_done),
        C.c_int(n_threads), C.c_int(nfev_stiff_detect), _ip(stiff),
        C.c_int(ev_set), gcb, C.c_int(ne), _ip(term), _ip(direc), C.c_int(capacity),
        _dp(t_ev), _dp(y_ev), _ip(ev_cnt))
    if rc != 0:
        raise RuntimeError("xsq_oracle_rk_events_batch failed (device arithmetic only)")
    return dict(t=te, y=y_eval, t_final=t_final, y_final=y_final, h_next=h_next,
                n_accepted=n_acc, n_rejected=n_rej, nfev=nfev, status=status,
                n_eval_done=n_done, stiff_flags=stiff, t_events=t_ev, y_events=y_ev,
                event_counts=ev_cnt)


def swag_batch(rhs, t_span, y0, params=None, rtol=1e-3, atol=1e-6,
               first_step=None, max_step=np.inf, k_max=12, t_eval=None,
               max_steps=0, n_threads=1, user_fn=None):
    """Integrate N lanes with the C restatement of SWAG
    (oracle/xsq_oracle_swag.c).  Same conventions as rk_batch."""
    lib = load()
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=float)))
    N, n = y0.shape
    rtol_v, atol_v = O.validate_tol(float(rtol), atol, y0[0])
    atol_v = np.ascontiguousarray(np.broadcast_to(atol_v, (n,)), dtype=float)
    if params is not None:
        params = np.ascontiguousarray(np.asarray(params, dtype=float))
        if params.ndim == 1:
            params = params.reshape(N, -1)
        p = params.shape[1]
    else:
        p = 0
    te = (np.ascontiguousarray(np.asarray(t_eval, dtype=float))
          if t_eval is not None else None)
    n_eval = te.size if te is not None else 0
    y_eval = np.empty((N, n, n_eval)) if n_eval else None
    t_final = np.empty(N)
    y_final = np.empty((N, n))
    ints = [np.empty(N, np.int32) for _ in range(6)]
    n_acc, n_fail, nfev, status, n_done, k_final = ints
    if rhs is not None:
        rid, cb = RHS_IDS[rhs], RHS_FN()
    else:
        rid = -1

        def _cb(tt, yp, pp, dyp):
            out = np.asarray(user_fn(tt, np.ctypeslib.as_array(yp, (n,))),
                             dtype=float)
            for i in range(n):
                dyp[i] = out[i]
        cb = RHS_FN(_cb)
    rc = lib.xsq_oracle_swag_batch(
        C.c_int(rid), cb, C.c_int(n), C.c_int(p), C.c_int64(N), _dp(y0),
        _dp(params), C.c_double(t_span[0]), C.c_double(t_span[1]),
        C.c_double(float(rtol_v)), _dp(atol_v),
        C.c_double(first_step if first_step is not None else 0.0),
        C.c_double(max_step), C.c_int(k_max), _dp(te), C.c_int(n_eval),
        _dp(y_eval), C.c_int(max_steps), _dp(t_final), _dp(y_final),
        _ip(n_acc), _ip(n_fail), _ip(nfev), _ip(status), _ip(n_done),
        _ip(k_final), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError("xsq_oracle_swag_batch failed")
    return dict(t=te, y=y_eval, t_final=t_final, y_final=y_final,
                n_accepted=n_acc, n_rejected=n_fail, nfev=nfev, status=status,
                n_eval_done=n_done, k_final=k_final)


def swag_events_batch(rhs, t_span, y0, events, terminal, direction, capacity, params=None,
                      rtol=1e-3, atol=1e-6, first_step=None, max_step=np.inf, k_max=12,
                      t_eval=None, max_steps=0, n_threads=1):
    """swag_batch with scipy's `events=` on SWAG's interpolant, in the kernels'
    arithmetic (xsq_oracle_swag.c; call inside ``with device_math():``).  `events`:
    the name of a built-in set (EVENT_SETS) or a list of callables g(t, y)."""
    lib = load()
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=float)))
    N, n = y0.shape
    rtol_v, atol_v = O.validate_tol(float(rtol), atol, y0[0])
    atol_v = np.ascontiguousarray(np.broadcast_to(atol_v, (n,)), dtype=float)
    if params is not None:
        params = np.ascontiguousarray(np.asarray(params, dtype=float))
        if params.ndim == 1:
            params = params.reshape(N, -1)
        p = params.shape[1]
    else:
        p = 0
    te = (np.ascontiguousarray(np.asarray(t_eval, dtype=float)) if t_eval is not None else None)
    n_eval = te.size if te is not None else 0
    y_eval = np.empty((N, n, n_eval)) if n_eval else None
    t_final, y_final = np.empty(N), np.empty((N, n))
    ints = [np.empty(N, np.int32) for _ in range(6)]
    n_acc, n_fail, nfev, status, n_done, k_final = ints
    if isinstance(events, str):
        ev_set, ne = EVENT_SETS[events]
        gcb = EVENT_FN()
    else:
        ev_set, ne = -1, len(events)

        def _g(k, tt, yp, pp):
            return float(events[k](tt, np.ctypeslib.as_array(yp, (n,))))
        gcb = EVENT_FN(_g)
    term = np.ascontiguousarray(np.asarray(terminal, dtype=np.int32))
    direc = np.ascontiguousarray(np.asarray(direction, dtype=np.int32))
    t_ev = np.full((N, ne, capacity), np.nan)
    y_ev = np.full((N, ne, capacity, n), np.nan)
    ev_cnt = np.zeros((N, ne), np.int32)
    rc = lib.xsq_oracle_swag_events_batch(
        C.c_int(RHS_IDS[rhs]), RHS_FN(), C.c_int(n), C.c_int(p), C.c_int64(N), _dp(y0),
        _dp(params), C.c_double(t_span[0]), C.c_double(t_span[1]),
        C.c_double(float(rtol_v)), _dp(atol_v),
        C.c_double(first_step if first_step is not None else 0.0),
        C.c_double(max_step), C.c_int(k_max), _dp(te), C.c_int(n_eval),
        _dp(y_eval), C.c_int(max_steps), _dp(t_final), _dp(y_final),
        _ip(n_acc), _ip(n_fail), _ip(nfev), _ip(status), _ip(n_done),
        _ip(k_final), C.c_int(n_threads), C.c_int(ev_set), gcb, C.c_int(ne), _ip(term),
        _ip(direc), C.c_int(capacity), _dp(t_ev), _dp(y_ev), _ip(ev_cnt))
    if rc != 0:
        raise RuntimeError("xsq_oracle_swag_events_batch failed (device arithmetic only)")
    return dict(t=te, y=y_eval, t_final=t_final, y_final=y_final,
                n_accepted=n_acc, n_rejected=n_fail, nfev=nfev, status=status,
                n_eval_done=n_done, k_final=k_final, t_events=t_ev, y_events=y_ev,
                event_counts=ev_cnt)


VEC_RHS_FN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double),
                         C.POINTER(C.c_double), C.c_int64, C.c_void_p)


def rkc_solve(y0, t_span, rtol=1e-3, atol=1e-6, first_step=None,
              max_step=np.inf, const_jac=False, rho=None, t_eval=None,
              max_steps=0, fun=None):
    """SSV2stab through the C restatement (oracle/xsq_oracle_rkc.c).
    ``fun=None`` uses the built-in 2-D reaction-diffusion RHS (len(y0) must be
    a square); otherwise ``fun(t, y) -> dy`` is called back from C.
    ``rho``: constant playing the role of rho_jac, None -> power iteration."""
    lib = load()
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=float))
    n = y0.size
    te = (np.ascontiguousarray(np.asarray(t_eval, dtype=float))
          if t_eval is not None else None)
    n_eval = te.size if te is not None else 0
    y_eval = np.empty((n_eval, n)) if n_eval else None
    t_final = C.c_double()
    y_final = np.empty(n)
    counters = np.zeros(6, np.int32)
    if fun is None:
        kind, cb = 1, VEC_RHS_FN()
    else:
        kind = 0

        def _cb(t, yp, dyp, nn, ctx):
            yv = np.ctypeslib.as_array(yp, (n,))
            np.ctypeslib.as_array(dyp, (n,))[:] = fun(t, yv)
        cb = VEC_RHS_FN(_cb)
    rc = lib.xsq_oracle_rkc_solve(
        C.c_int(kind), cb, C.c_int64(n), _dp(y0), C.c_double(t_span[0]),
        C.c_double(t_span[1]), C.c_double(rtol), C.c_double(atol),
        C.c_double(first_step if first_step is not None else 0.0),
        C.c_double(max_step), C.c_int(1 if const_jac else 0),
        C.c_double(rho if rho is not None else 0.0), _dp(te), C.c_int(n_eval),
        _dp(y_eval), C.c_int(max_steps), C.byref(t_final), _dp(y_final),
        _ip(counters))
    if rc != 0:
        raise RuntimeError("xsq_oracle_rkc_solve failed")
    return dict(t=te, y=y_eval.T if y_eval is not None else None,
                t_final=t_final.value, y_final=y_final,
                n_accepted=int(counters[0]), n_rejected=int(counters[1]),
                nfev=int(counters[2]), nfesig=int(counters[3]),
                maxm=int(counters[4]), status=int(counters[5]))
