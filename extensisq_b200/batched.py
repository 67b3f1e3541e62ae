"""``solve_ivp_batched`` -- the batched, device-resident counterpart of
``scipy.integrate.solve_ivp(fun, t_span, y0, method=Cls, t_eval=..., **opts)``
for the reference's explicit Runge-Kutta classes.

Host logic only: argument validation with the reference's semantics and
exception types (``extensisq/common.py:30-54, 166-185, 187-214``; scipy
``_ivp/common.py:10-23``; ``_ivp/ivp.py:600-620`` for ``t_eval``), layout
conversion to the SoA buffers the kernel wants, and the ctypes call into
``libxsq.so``.  PyTorch is the memory / stream provider.  No arithmetic of the
method happens here and there is no CPU path.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .tableaux import RungeKutta, RungeKuttaNystrom, SWAG

__all__ = ["DeviceRHS", "BatchedOdeResult", "solve_ivp_batched", "NFS"]

# The reference keeps a module-global failed-step counter reset by every
# constructor (common.py:14, 220).  For ensembles it holds the SUM over lanes of
# the last solve; per-lane values are in BatchedOdeResult.n_rejected.
NFS = np.array(0)


class DeviceRHS:
    """A right-hand side that exists as device code.

    ``DeviceRHS.builtin("lorenz63")`` or ``DeviceRHS.from_source(src, "rhs",
    n_state, n_param)`` where ``src`` defines
    ``__device__ void rhs(double t, const double* y, const double* p,
    double* dy)``.  Replaces the Python callable ``fun`` of common.py:187."""

    def __init__(self, handle, n_state, n_param, name):
        self.handle, self.n_state, self.n_param, self.name = (
            handle, n_state, n_param, name)

    _builtin_cache = {}

    @classmethod
    def builtin(cls, name):
        if name in cls._builtin_cache:
            return cls._builtin_cache[name]
        lib = _lib.load()
        h, ns, npar = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.xsq_rhs_builtin(name.encode(), C.byref(h), C.byref(ns),
                                       C.byref(npar)))
        r = cls(h.value, ns.value, npar.value, name)
        cls._builtin_cache[name] = r
        return r

    @classmethod
    def from_source(cls, cuda_src, entry, n_state, n_param=0):
        lib = _lib.load()
        h = C.c_int32()
        _lib.check(lib.xsq_rhs_register_source(
            cuda_src.encode(), entry.encode(), int(n_state), int(n_param),
            C.byref(h)))
        return cls(h.value, int(n_state), int(n_param), f"user:{entry}")


class DeviceEvents:
    """The ``events=`` argument of scipy's ``solve_ivp`` as device code.

    ``DeviceEvents.from_source(src, "event", n_events, terminal=[...],
    direction=[...])`` where ``src`` defines
    ``__device__ double event(int k, double t, const double* y, const double* p)``
    returning the value of event function ``k``.  ``terminal[k]``: False / 0
    (never), True / 1 (first occurrence) or n (n-th occurrence); ``direction[k]``
    as in scipy (ivp.py prepare_events / find_active_events)."""

    def __init__(self, handle, n_events, terminal, direction, name):
        self.handle, self.n_events, self.name = handle, n_events, name
        self.terminal = [int(v) for v in terminal]
        self.direction = [int(np.sign(v)) for v in direction]
        if len(self.terminal) != n_events or len(self.direction) != n_events:
            raise ValueError("terminal / direction need one entry per event")
        if any(v < 0 for v in self.terminal):
            raise ValueError("The `terminal` attribute of each event must be a "
                             "boolean or positive integer.")

    @classmethod
    def from_source(cls, cuda_src, entry, n_events, terminal=None, direction=None):
        lib = _lib.load()
        h = C.c_int32()
        _lib.check(lib.xsq_events_register_source(
            cuda_src.encode(), entry.encode(), int(n_events), C.byref(h)))
        return cls(h.value, int(n_events), terminal or [0] * n_events,
                   direction or [0] * n_events, f"events:{entry}")

    def with_attributes(self, terminal=None, direction=None):
        """Same compiled functions, other terminal / direction attributes."""
        return DeviceEvents(self.handle, self.n_events,
                            self.terminal if terminal is None else terminal,
                            self.direction if direction is None else direction,
                            self.name)


@dataclass
class BatchedOdeResult:
    """Per-lane results; mirrors scipy's OdeResult field names where they
    exist (ivp.py:758-760)."""
    t: object                   # t_eval tensor [n_eval] or None
    y: object                   # [N, n, n_eval] or None
    t_final: torch.Tensor       # [N]
    y_final: torch.Tensor       # [N, n]
    h_next: torch.Tensor        # [N]
    n_accepted: torch.Tensor    # int32 [N]
    n_rejected: torch.Tensor    # int32 [N]   (the reference's NFS)
    nfev: torch.Tensor          # int32 [N]
    status: torch.Tensor        # int32 [N]   xsq_lane_status
    n_eval_done: object = None  # int32 [N] or None
    stiff_flags: object = None  # int32 [N]: 1 stiff (real root), 2 stiff
    #                             (complex pair), 4 oscillatory + many failures
    t_events: object = None     # [N, n_events, capacity], NaN where unused
    y_events: object = None     # [N, n_events, capacity, n]
    event_counts: object = None  # int32 [N, n_events] occurrences found
    njev: int = 0
    nlu: int = 0
    sol: object = None          # BatchedOdeSolution when dense_output=True

    @property
    def success(self):
        return self.status >= 0

    def message(self, lane):
        return _lib.LANE_MESSAGES[int(self.status[lane])]

    def lane_tensors(self):
        """Every per-lane result tensor (dim 0 = lanes) by name, e.g. as the
        argument of ``gather_result`` after a sharded solve."""
        names = ("y", "t_final", "y_final", "h_next", "n_accepted", "n_rejected", "nfev",
                 "status", "n_eval_done", "stiff_flags", "t_events", "y_events",
                 "event_counts")
        return {k: getattr(self, k) for k in names if getattr(self, k) is not None}


class BatchedOdeSolution:
    """``res.sol`` of ``solve_ivp(..., dense_output=True)`` (ivp.py:730-741,
    scipy's OdeSolution) for an ensemble: ``sol(t)`` is the methods' own dense
    output -- Horner / BS5 low / best / cubic / SWAG's interpolant, the one
    ``t_eval`` is served from -- at any time(s) inside the integrated span.

    Keeping the interpolant of every step of 10^6 lanes would take more memory
    than the GPU has, so nothing is stored: the step sequence of a lane is a
    deterministic function of its inputs, and a call repeats the solve with the
    requested times as ``t_eval``.  The values are therefore exactly those a
    stored interpolant would give (``sol(t_eval) == res.y`` bit for bit, and
    ``sol(t_final) == y_final``).  Lanes that ended early (terminal event,
    failure) return NaN beyond their end, like ``res.y``."""

    def __init__(self, fun, t_span, y0, method, kwargs):
        self._args = (fun, t_span, y0, method)
        self._kw = kwargs
        self.t_min, self.t_max = min(t_span), max(t_span)
        self.ascending = t_span[1] >= t_span[0]

    def __call__(self, t):
        t_np = np.asarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t,
                          dtype=float)
        scalar = t_np.ndim == 0
        pts = np.atleast_1d(t_np)
        if pts.ndim != 1:
            raise ValueError("`t` must be a float or a 1-D array.")
        if pts.size and (pts.min() < self.t_min or pts.max() > self.t_max):
            raise ValueError("`t` outside of the integrated span (no extrapolation on "
                             "the device).")
        uniq, inverse = np.unique(pts, return_inverse=True)
        order = uniq if self.ascending else uniq[::-1].copy()
        fun, t_span, y0, method = self._args
        r = solve_ivp_batched(fun, t_span, y0, method, t_eval=order, **self._kw)
        idx = inverse if self.ascending else (len(uniq) - 1 - inverse)
        y = r.y[:, :, torch.as_tensor(idx, device=r.y.device)]
        return y[:, :, 0] if scalar else y


def _as_device(x, device, dtype=torch.float64):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(device,
                                                         non_blocking=True)


def _upload_user_tableau(cls):
    cls.validate()
    lib = _lib.load()
    t = _lib.XsqTableau()
    s = cls.n_stages
    if s > _lib.XSQ_MAX_STAGES - 1:
        raise ValueError(f"n_stages > {_lib.XSQ_MAX_STAGES - 1} not supported")
    t.n_stages, t.order, t.order_secondary = s, cls.order, cls.order_secondary
    for i in range(s):
        for j in range(s):
            t.A[i][j] = float(cls.A[i, j])
        t.B[i] = float(cls.B[i])
        t.C[i] = float(cls.C[i])
    for i in range(s + 1):
        t.E[i] = float(cls.E[i])
    if isinstance(cls.P, np.ndarray):
        if cls.P.shape[1] > _lib.XSQ_MAX_POLY:
            raise ValueError("P has too many columns")
        t.n_poly = cls.P.shape[1]
        for i in range(s + 1):
            for k in range(t.n_poly):
                t.P[i][k] = float(cls.P[i, k])
    else:
        t.n_poly = 0
    kb = _sc_tuple(cls.sc_params)
    for i in range(4):
        t.sc_params[i] = kb[i]
    # stiffness detection needs both; NotImplemented disables it (common.py:155)
    ok = all(isinstance(getattr(cls, k), (int, float)) for k in ("stbrad", "tanang"))
    t.stbrad = float(cls.stbrad) if ok else 0.0
    t.tanang = float(cls.tanang) if ok else 0.0
    _lib.check(lib.xsq_tableau_load(C.byref(t)))


_SC = {"G": (0.7, -0.4, 0, 0.9), "S": (0.6, -0.2, 0, 0.9),
       "standard": (1, 0, 0, 0.9)}          # common.py:167-169


def _sc_tuple(sc_params):
    if isinstance(sc_params, str) and sc_params in _SC:
        return tuple(float(v) for v in _SC[sc_params])
    if isinstance(sc_params, tuple) and len(sc_params) == 4:
        return tuple(float(v) for v in sc_params)
    raise ValueError('sc_params should be a tuple of length 4 or one '
                     'of the strings "G", "S", "W" or "standard"')


def solve_ivp_batched(fun, t_span, y0, method, t_eval=None, params=None,
                      rtol=1e-3, atol=1e-6, first_step=None, max_step=np.inf,
                      sc_params=None, interpolant=None, k_max=None,
                      nfev_stiff_detect=5000, max_steps=None, events=None,
                      max_event_records=16,
                      forced_steps=None, device=None, stream=None,
                      dense_output=False, **extraneous):
    """Integrate N independent systems ``y' = fun(t, y; params_i)``.

    Parameters follow ``solve_ivp`` / ``RungeKutta.__init__`` (common.py:187):

    fun : DeviceRHS or str (built-in name)
    t_span : (t0, tf), shared by all lanes
    y0 : [N, n] float64 tensor/array (a 1-D [n] input is one lane)
    method : a tableau class from this package or a RungeKutta subclass
    t_eval : [n_eval] times inside t_span, sorted along the direction
    params : [N, p] per-lane parameters of the RHS
    rtol : float;  atol : float or [n];  first_step, max_step : float
        (first_step may also be a [N] tensor, one first step per lane: pass the
        ``h_next`` of the result that ended at ``t_span[0]`` to continue a solve)
    sc_params : "G" | "S" | "standard" | (kb1, kb2, a, g)
    interpolant : BS5 only, 'best' | 'low' | 'free' (bogacki.py:217)
    k_max : SWAG only, maximum order 1..12 (shampine.py:99-103)
    events : DeviceEvents or None -- scipy's `events=` (terminal, direction,
        root finding with brentq on the method's dense output, ivp.py); the
        result carries t_events / y_events / event_counts per lane, lanes stopped
        by a terminal event have status 1 and t_final / y_final at the event
    max_event_records : occurrences kept per event function and lane
    nfev_stiff_detect : stiffness diagnosis every this many evaluations or on
        >= 10 failed steps in 40 (common.py:150-164, 370-516); 0 turns it off.
        Instead of warnings the result carries per-lane ``stiff_flags``
    max_steps : attempted-step budget per lane (GPU safety net, no reference
        analogue)
    forced_steps : [k] sequence of |h|; takes exactly these steps, accepting
        each (parity mode of BASELINE.json's north_star)
    dense_output : bool -- as in ``solve_ivp``: the result carries ``sol``, a
        callable :class:`BatchedOdeSolution` (``sol(t)`` -> [N, n, len(t)])

    Returns a :class:`BatchedOdeResult` with device tensors.
    """
    global NFS
    if extraneous:
        import warnings
        warnings.warn("The following arguments have no effect for a chosen "
                      f"solver: {', '.join(f'`{k}`' for k in extraneous)}.",
                      stacklevel=2)
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("extensisq_b200 needs a CUDA device; there is no "
                           "CPU fallback")
    if isinstance(fun, str):
        fun = DeviceRHS.builtin(fun)
    if not isinstance(fun, DeviceRHS):
        raise TypeError("`fun` must be a DeviceRHS or the name of a built-in "
                        "right-hand side; Python callables cannot run on the "
                        "device")
    is_swag = isinstance(method, type) and issubclass(method, SWAG)
    if not (isinstance(method, type) and
            (issubclass(method, RungeKutta) or is_swag)):
        raise ValueError("`method` must be one of the tableau classes, SWAG or "
                         "a RungeKutta subclass")
    if is_swag:
        k_max = method.k_max if k_max is None else k_max
        if not (isinstance(k_max, int) and 0 < k_max < 13):  # shampine.py:102
            raise ValueError("`k_max` should be an integer between 1 and 12.")
        if forced_steps is not None or sc_params is not None or \
                interpolant is not None:
            raise ValueError("forced_steps / sc_params / interpolant do not "
                             "apply to SWAG")
    elif k_max is not None:
        raise ValueError("`k_max` only applies to SWAG")
    if events is not None:
        if not isinstance(events, DeviceEvents):
            raise ValueError("`events` must be a DeviceEvents (device code; a "
                             "Python callable cannot run inside the kernel)")
        if forced_steps is not None:
            raise ValueError("events do not apply to forced_steps")
        if not (isinstance(max_event_records, int) and max_event_records > 0):
            raise ValueError("`max_event_records` must be a positive integer")
    is_rkn = not is_swag and issubclass(method, RungeKuttaNystrom)
    if is_rkn:
        # RungeKuttaNystrom.__init__, common.py:1240-1277
        if method._xsq_method is None:
            raise ValueError("user defined Runge-Kutta-Nystrom tableaux are not supported")
        if fun.n_state % 2:
            raise AssertionError('This method is for second order problems'
                                 ' and `fun` should have signature: [v, a] = fun(t, [x, v]).')
        if t_eval is not None or events is not None or interpolant is not None:
            raise ValueError("Runge-Kutta-Nystrom methods return the final state only on the "
                             "device (no t_eval / events / interpolant)")
        if not method.velocity_dependent and fun.name in ("vanderpol", "arenstorf"):
            raise AssertionError("This method is for velocity independent ODEs, "
                                 "but `fun` seems velocity dependent.")
        nfev_stiff_detect = 0      # the rectangular-domain diagnosis (common.py:1322) is host-only
    is_ckdisc = not is_swag and getattr(method, "_xsq_method", None) == _lib.METHOD_IDS["CKdisc"]
    if is_ckdisc:
        # CKdisc.__init__(fun, t0, y0, t_bound, **extraneous) passes
        # nfev_stiff_detect=0 itself (cash.py:238-240) and has its own step rule
        if forced_steps is not None or interpolant is not None:
            raise ValueError("forced_steps / interpolant do not apply to CKdisc")
        nfev_stiff_detect = 0
    if device is None:
        device = (y0.device if isinstance(y0, torch.Tensor) and y0.is_cuda
                  else torch.device("cuda", torch.cuda.current_device()))
    device = torch.device(device)
    t0, tf = map(float, t_span)
    n, p = fun.n_state, fun.n_param

    # --- validation, same order/messages as the reference -----------------
    if not isinstance(rtol, float):
        raise ValueError("`rtol` must be a float.")
    if rtol < 0:
        raise ValueError("`rtol` must be positive.")
    atol_np = np.atleast_1d(np.asarray(atol, dtype=float))
    if atol_np.ndim > 1 or atol_np.size not in (1, n):
        raise ValueError("`atol` has wrong shape.")
    if np.any(atol_np < 0):
        raise ValueError("`atol` must be positive.")
    if max_step <= 0:
        raise ValueError("`max_step` must be positive.")
    first_lanes = None
    if first_step is not None and np.ndim(first_step) > 0:
        # one first step per lane: resume with the h_next of a previous solve (the
        # reference's manual stepping continues with the controller's proposal)
        if forced_steps is not None:
            raise ValueError("a per-lane `first_step` does not apply to forced_steps")
        first_lanes = first_step
        first_step = None
    if first_step is not None and forced_steps is None:
        if first_step <= 0:
            raise ValueError("`first_step` must be positive.")
        if first_step > abs(tf - t0):
            raise ValueError("`first_step` exceeds bounds.")
    if not (isinstance(nfev_stiff_detect, int) and nfev_stiff_detect >= 0):
        raise ValueError("`nfev_stiff_detect` must be a non-negative integer.")
    sc = _sc_tuple(sc_params) if sc_params is not None else None
    if interpolant not in (None, "best", "low", "free"):
        raise ValueError("interpolant should be one of: 'best', 'low', 'free'")

    # Everything below -- host->device copies, transposes, result allocation and
    # the kernels -- is enqueued on ONE stream: the caller's `stream`, made
    # current for the duration of the call after it has waited for the work
    # already queued on the previously current stream (the inputs may have been
    # produced there).  Allocations made under it belong to it, so the caching
    # allocator cannot recycle a buffer the kernel still uses.
    run_stream = stream if stream is not None else torch.cuda.current_stream(device)
    if stream is not None:
        run_stream.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.device(device), torch.cuda.stream(run_stream):
        y0_t = _as_device(y0, device)
        if y0_t.ndim == 1:
            y0_t = y0_t[None, :]
        if y0_t.ndim != 2 or y0_t.shape[1] != n:
            raise ValueError(f"`y0` must have shape [N, {n}]")
        N = y0_t.shape[0]
        y0_soa = y0_t.t().contiguous()                     # [n, N]
        if p > 0:
            if params is None:
                raise ValueError(f"`params` [N, {p}] required by {fun.name}")
            prm = _as_device(params, device)
            if prm.ndim == 1:
                prm = prm[:, None] if p == 1 else prm[None, :]
            if prm.shape[0] == 1 and N > 1:
                prm = prm.expand(N, p)
            if prm.shape != (N, p):
                raise ValueError(f"`params` must have shape [N, {p}]")
            prm_soa = prm.t().contiguous()
        else:
            prm_soa = None

        n_eval = 0
        te = None
        if t_eval is not None:                              # ivp.py:600-612
            te = _as_device(t_eval, device)
            if te.ndim != 1:
                raise ValueError("`t_eval` must be 1-dimensional.")
            if te.numel() > 0:
                lo, hi = min(t0, tf), max(t0, tf)
                if bool((te < lo).any()) or bool((te > hi).any()):
                    raise ValueError("Values in `t_eval` are not within "
                                     "`t_span`.")
                d = te[1:] - te[:-1]
                if (tf > t0 and bool((d <= 0).any())) or \
                        (tf < t0 and bool((d >= 0).any())):
                    raise ValueError("Values in `t_eval` are not properly "
                                     "sorted.")
            n_eval = te.numel()
            te = te.contiguous()
        hf = None
        if forced_steps is not None:
            hf = _as_device(forced_steps, device).contiguous()
            if hf.ndim != 1 or hf.numel() < 1:
                raise ValueError("`forced_steps` must be a non-empty 1-D "
                                 "sequence")

        if is_swag:
            mid = 0
        elif method._xsq_method is None:
            _upload_user_tableau(method)
            mid = _lib.XSQ_METHOD_USER
        else:
            mid = method._xsq_method

        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        # rows padded to a multiple of 4 doubles: the kernel writes aligned
        # 32-byte groups (include/xsq.h); the result exposes [:, :, :n_eval]
        pitch = (n_eval + 3) // 4 * 4
        y_eval_buf = torch.empty((N, n, pitch), **f64) if n_eval else None
        y_eval = y_eval_buf[:, :, :n_eval] if n_eval else None
        t_final = torch.empty(N, **f64)
        y_final = torch.empty((n, N), **f64)
        h_next = torch.empty(N, **f64)
        n_acc = torch.empty(N, **i32)
        n_rej = torch.empty(N, **i32)
        nfev = torch.empty(N, **i32)
        status = torch.empty(N, **i32)
        n_done = torch.empty(N, **i32) if n_eval else None
        stiff = torch.zeros(N, **i32)
        t_ev = y_ev = ev_cnt = None
        if events is not None:
            cap = int(max_event_records)
            t_ev = torch.full((N, events.n_events, cap), float("nan"), **f64)
            y_ev = torch.full((N, events.n_events, cap, n), float("nan"), **f64)
            ev_cnt = torch.zeros((N, events.n_events), **i32)

        a = _lib.XsqRkArgs()
        a.struct_size = C.sizeof(_lib.XsqRkArgs)
        a.method, a.rhs = mid, fun.handle
        a.n_state, a.n_param = n, p
        a.interpolant = _lib.INTERPOLANTS[interpolant]
        a.n_lanes = N
        a.y0 = y0_soa.data_ptr()
        a.params = prm_soa.data_ptr() if prm_soa is not None else None
        a.t0, a.t_bound = t0, tf
        a.rtol = rtol
        atol_c = (C.c_double * atol_np.size)(*atol_np.tolist())
        a.atol = C.cast(atol_c, C.POINTER(C.c_double))
        a.n_atol = atol_np.size
        a.use_sc_params = 1 if sc is not None else 0
        if sc is not None:
            for i in range(4):
                a.sc_params[i] = sc[i]
        a.first_step = float(first_step) if first_step is not None else 0.0
        if first_lanes is not None:
            fl = _as_device(first_lanes, device).to(torch.float64).contiguous()
            if fl.shape != (N,):
                raise ValueError("a per-lane `first_step` must have one entry per lane")
            if not bool((fl > 0).all()):
                raise ValueError("`first_step` must be positive.")
            a.first_step_lanes = fl.data_ptr()
        a.max_step = float(max_step)
        a.t_eval = te.data_ptr() if n_eval else None
        a.n_eval = n_eval
        a.max_steps = int(max_steps) if max_steps else 0
        a.y_eval = y_eval_buf.data_ptr() if n_eval else None
        a.h_forced = hf.data_ptr() if hf is not None else None
        a.n_forced = hf.numel() if hf is not None else 0
        a.t_final = t_final.data_ptr()
        a.y_final = y_final.data_ptr()
        a.h_next = h_next.data_ptr()
        a.n_accepted = n_acc.data_ptr()
        a.n_rejected = n_rej.data_ptr()
        a.nfev = nfev.data_ptr()
        a.status = status.data_ptr()
        a.n_eval_done = n_done.data_ptr() if n_eval else None
        a.nfev_stiff_detect = 0 if is_swag else nfev_stiff_detect
        a.stiff_flags = stiff.data_ptr()
        if events is not None:
            ne = events.n_events
            term_c = (C.c_int32 * ne)(*events.terminal)
            dir_c = (C.c_int32 * ne)(*events.direction)
            a.events, a.n_event_fns = events.handle, ne
            a.ev_terminal = C.cast(term_c, _lib._ip)
            a.ev_direction = C.cast(dir_c, _lib._ip)
            a.ev_capacity = int(max_event_records)
            a.t_events = t_ev.data_ptr()
            a.y_events = y_ev.data_ptr()
            a.ev_count = ev_cnt.data_ptr()
        st = run_stream
        if is_swag:
            _lib.check(lib.xsq_swag_solve(C.byref(a), k_max,
                                          C.c_void_p(st.cuda_stream)))
        else:
            _lib.check(lib.xsq_rk_solve(C.byref(a),
                                        C.c_void_p(st.cuda_stream)))
        # the kernel runs on `st`; tensors above stay referenced by the result

    res = BatchedOdeResult(
        t=te, y=y_eval, t_final=t_final, y_final=y_final.t(), h_next=h_next,
        n_accepted=n_acc, n_rejected=n_rej, nfev=nfev, status=status,
        n_eval_done=n_done, stiff_flags=stiff, t_events=t_ev, y_events=y_ev,
        event_counts=ev_cnt)
    res._keepalive = (y0_soa, prm_soa, hf, atol_c)
    if dense_output:
        if forced_steps is not None:
            raise ValueError("dense_output does not apply to forced_steps")
        res.sol = BatchedOdeSolution(fun, (t0, tf), y0, method, dict(
            params=params, rtol=rtol, atol=atol, first_step=first_lanes if first_lanes is not None
            else first_step, max_step=max_step, sc_params=sc_params, interpolant=interpolant,
            k_max=k_max if is_swag else None, nfev_stiff_detect=nfev_stiff_detect,
            max_steps=max_steps, events=events, max_event_records=max_event_records,
            device=device, stream=stream))
    return res


def trim_memory(device=None):
    """Return the scratch the library keeps cached in the device's memory pool
    (work queue, init pass, stiffness probe queue) to the driver."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    _lib.check(lib.xsq_trim_memory(dev.index if dev.index is not None else 0))


def update_nfs(result):
    """Mirror the reference's global NFS (sum over lanes; forces a sync)."""
    NFS[()] = int(result.n_rejected.sum().item())
    return int(NFS)
