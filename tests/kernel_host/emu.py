"""TEST INFRASTRUCTURE -- builds and drives tests/kernel_host/emu.cpp: the CUDA
kernel sources compiled for the host (one thread walks all trajectories)."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_lib = None
_lib_ev = None
_lib_ev_nt = None
_bits = None


def build(events=False, no_terminal=False):
    src = [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "cuda_shim.h")]
    src += [os.path.join(ROOT, "extensisq_b200", "csrc", f) for f in
            ("xsq_rk_core.cuh", "xsq_rk_fast.cuh", "xsq_math.cuh", "xsq_intrin.cuh",
             "xsq_params.h", "xsq_rhs.cuh", "xsq_tableaux_gen.cuh", "xsq_math_tables_gen.cuh",
             "xsq_swag_core.cuh", "xsq_swag_fast.cuh")]
    out = os.path.join(HERE, "_build", ("xsq_emu_events_nt.so" if no_terminal else
                                        "xsq_emu_events.so") if events else "xsq_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if (not os.path.exists(out) or
            os.path.getmtime(out) < max(os.path.getmtime(f) for f in src)):
        subprocess.check_call(
            ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-w",
             "-I", os.path.join(ROOT, "extensisq_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
             "-I", "/usr/local/cuda/include", "-o", out, src[0]] +
            (["-DXSQ_EMU_EVENTS"] if events else []) +
            (["-DXSQ_EVENTS_NO_TERMINAL"] if no_terminal else []))
    return out


def load():
    global _lib, _bits
    if _lib is None:
        _lib = C.CDLL(build())
        raw = zlib.decompress(open(os.path.join(ROOT, "oracle", "rcp64h_delta.bin.z"), "rb").read())
        _bits = np.frombuffer(raw, dtype=np.uint8).copy()
        _lib.xsq_emu_set_rcp_table(_bits.ctypes.data_as(C.c_void_p))
        _lib.xsq_emu_detail.restype = C.c_char_p
    return _lib


def load_events(no_terminal=False):
    """The same sources compiled with the event machinery (XSQ_EVENTS_N = 3, the
    event functions of oracle/problems.py EVENT_SETS['lorenz_sections']);
    no_terminal: the variant NVRTC builds for event sets without a terminal event."""
    global _lib_ev, _lib_ev_nt
    load()
    if no_terminal:
        if _lib_ev_nt is None:
            _lib_ev_nt = C.CDLL(build(events=True, no_terminal=True))
            _lib_ev_nt.xsq_emu_set_rcp_table(_bits.ctypes.data_as(C.c_void_p))
            _lib_ev_nt.xsq_emu_detail.restype = C.c_char_p
        return _lib_ev_nt
    if _lib_ev is None:
        _lib_ev = C.CDLL(build(events=True))
        _lib_ev.xsq_emu_set_rcp_table(_bits.ctypes.data_as(C.c_void_p))
        _lib_ev.xsq_emu_detail.restype = C.c_char_p
    return _lib_ev


def solve(rhs, t_span, y0, method, params=None, rtol=1e-3, atol=1e-6, first_step=None,
          max_step=np.inf, sc_params=None, interpolant=None, t_eval=None, forced_steps=None,
          nfev_stiff_detect=5000, max_steps=None, fast=True, queue_records=-1, k_max=None,
          events=None, max_event_records=16, event_queue_records=-1, no_terminal_build=False):
    """Same arguments as extensisq_b200.solve_ivp_batched (built-in rhs names,
    built-in methods); returns numpy arrays.  `events=(terminal, direction)`
    selects the events build (three Lorenz section functions, lorenz63 only);
    `event_queue_records`: -1 every possible record, 0 no queue (roots in the lane)."""
    from extensisq_b200 import _lib as L
    from extensisq_b200.batched import _sc_tuple
    lib = load_events(no_terminal_build) if events is not None else load()
    rid = {"lorenz63": 0, "vanderpol": 1, "arenstorf": 2}[rhs]      # include/xsq.h XSQ_RHS_*
    y0 = np.atleast_2d(np.asarray(y0, dtype=float))
    N, n = y0.shape
    y0_soa = np.ascontiguousarray(y0.T)
    prm_soa = None
    p = 0
    if params is not None:
        prm = np.asarray(params, dtype=float).reshape(N, -1)
        p = prm.shape[1]
        prm_soa = np.ascontiguousarray(prm.T)
    te = np.ascontiguousarray(np.asarray(t_eval, dtype=float)) if t_eval is not None else None
    n_eval = te.size if te is not None else 0
    pitch = (n_eval + 3) // 4 * 4
    hf = np.ascontiguousarray(np.asarray(forced_steps, dtype=float)) if forced_steps is not None else None
    y_eval = np.full((N, n, pitch), np.nan) if n_eval else None
    t_final, h_next = np.empty(N), np.empty(N)
    y_final = np.empty((n, N))
    ints = {k: np.zeros(N, np.int32) for k in ("n_accepted", "n_rejected", "nfev", "status",
                                               "n_eval_done", "stiff_flags")}
    atol_np = np.atleast_1d(np.asarray(atol, dtype=float))
    a = L.XsqRkArgs()
    a.struct_size = C.sizeof(L.XsqRkArgs)
    is_swag = getattr(method, "__name__", "") == "SWAG"
    a.method, a.rhs = (0 if is_swag else method._xsq_method), rid
    a.n_state, a.n_param = n, p
    a.interpolant = L.INTERPOLANTS[interpolant]
    a.n_lanes = N

    def ptr(x):
        return x.ctypes.data if x is not None else None
    a.y0, a.params = ptr(y0_soa), ptr(prm_soa)
    a.t0, a.t_bound = float(t_span[0]), float(t_span[1])
    a.rtol = float(rtol)
    atol_c = (C.c_double * atol_np.size)(*atol_np.tolist())
    a.atol = C.cast(atol_c, C.POINTER(C.c_double))
    a.n_atol = atol_np.size
    sc = _sc_tuple(sc_params) if sc_params is not None else None
    a.use_sc_params = 1 if sc is not None else 0
    if sc is not None:
        for i in range(4):
            a.sc_params[i] = sc[i]
    a.first_step = float(first_step) if first_step is not None else 0.0
    a.max_step = float(max_step)
    a.t_eval, a.n_eval = ptr(te), n_eval
    a.max_steps = int(max_steps) if max_steps else 0
    a.y_eval = ptr(y_eval)
    a.h_forced, a.n_forced = ptr(hf), (hf.size if hf is not None else 0)
    a.t_final, a.y_final, a.h_next = ptr(t_final), ptr(y_final), ptr(h_next)
    a.n_accepted, a.n_rejected = ptr(ints["n_accepted"]), ptr(ints["n_rejected"])
    a.nfev, a.status = ptr(ints["nfev"]), ptr(ints["status"])
    a.n_eval_done = ptr(ints["n_eval_done"]) if n_eval else None
    a.nfev_stiff_detect = int(nfev_stiff_detect)
    a.stiff_flags = ptr(ints["stiff_flags"])
    if events is not None:
        term, direc = events
        ne, cap = 3, int(max_event_records)
        t_ev = np.full((N, ne, cap), np.nan)
        y_ev = np.full((N, ne, cap, n), np.nan)
        ev_cnt = np.zeros((N, ne), np.int32)
        term_c = (C.c_int32 * ne)(*term)
        direc_c = (C.c_int32 * ne)(*direc)
        a.events, a.n_event_fns = 1, ne
        a.ev_terminal = C.cast(term_c, C.POINTER(C.c_int32))
        a.ev_direction = C.cast(direc_c, C.POINTER(C.c_int32))
        a.ev_capacity = cap
        a.t_events, a.y_events, a.ev_count = ptr(t_ev), ptr(y_ev), ptr(ev_cnt)
        lib.xsq_emu_set_event_queue(C.c_longlong(event_queue_records))
    used = C.c_int(0)
    if is_swag:
        a.nfev_stiff_detect = 0
        rc = lib.xsq_emu_swag_solve(C.byref(a), int(k_max) if k_max else 12)
    else:
        rc = lib.xsq_emu_rk_solve(C.byref(a), 1 if fast else 0, C.c_longlong(queue_records),
                                  C.byref(used))
    if rc != 0:
        raise RuntimeError(f"emu rc={rc}: {lib.xsq_emu_detail().decode()}")
    out = dict(t_final=t_final, y_final=np.ascontiguousarray(y_final.T), h_next=h_next,
               y=(y_eval[:, :, :n_eval] if n_eval else None), used_fast=bool(used.value))
    out.update(ints)
    if events is not None:
        out.update(t_events=t_ev, y_events=y_ev, event_counts=ev_cnt)
    return out
