"""ctypes binding of ``libxsq.so`` (``include/xsq.h``).

There is no CPU fallback: if the library is missing, or a compute entry point
is called without a CUDA device, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XSQ_LIB: another build of the same library (kernel experiments, tools/variants.sh)
LIB_PATH = os.environ.get("XSQ_LIB") or os.path.join(_HERE, "libxsq.so")

XSQ_MAX_STAGES = 18
XSQ_MAX_POLY = 8
XSQ_MAX_LANE_STATE = 16
XSQ_MAX_WARP_STATE = 1024
XSQ_METHOD_USER = 100
XSQ_RHS_USER_BASE = 1000

METHOD_IDS = {"Ts5": 0, "BS5": 1, "CK5": 2, "Me4": 3, "Pr7": 4, "Pr8": 5,
              "Pr9": 6, "CFMR7osc": 7, "CKdisc": 8}
INTERPOLANTS = {None: 0, "free": 1, "low": 2, "best": 3}

# xsq_lane_status -> the reference's messages
LANE_MESSAGES = {
    0: "The solver successfully reached the end of the integration interval.",
    1: "A termination event occurred.",
    -1: "Required step size is less than spacing between numbers.",
    -2: "Overflow or underflow encountered.",
    -3: "tolerance too tight",
    -4: "spectral radius estimation did not converge",
    -5: "step budget (max_steps) exhausted",
    -6: "event queue exhausted",
}

# every symbol include/xsq.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "xsq_abi_version", "xsq_strerror", "xsq_last_error_detail",
    "xsq_device_info", "xsq_tableau_load", "xsq_tableau_get",
    "xsq_rhs_builtin", "xsq_rhs_register_source", "xsq_user_compile_check",
    "xsq_events_register_source", "xsq_events_compile_check",
    "xsq_rk_solve", "xsq_rk_solve_host", "xsq_swag_solve",
    "xsq_comm_unique_id", "xsq_comm_create", "xsq_comm_destroy",
    "xsq_pde_register_source", "xsq_pde_register_vector_source", "xsq_rkc_solve", "xsq_rkc_stage_bench", "xsq_rkc_stage_bench_tma",
    "xsq_launch_count", "xsq_trim_memory", "xsq_profile_enable", "xsq_profile_last", "xsq_profile_get",
    "xsq_fp64_peak",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class XsqTableau(C.Structure):
    _fields_ = [
        ("n_stages", C.c_int32), ("order", C.c_int32),
        ("order_secondary", C.c_int32), ("n_poly", C.c_int32),
        ("A", (C.c_double * XSQ_MAX_STAGES) * XSQ_MAX_STAGES),
        ("B", C.c_double * XSQ_MAX_STAGES),
        ("C", C.c_double * XSQ_MAX_STAGES),
        ("E", C.c_double * (XSQ_MAX_STAGES + 1)),
        ("P", (C.c_double * XSQ_MAX_POLY) * (XSQ_MAX_STAGES + 1)),
        ("sc_params", C.c_double * 4),
        ("stbrad", C.c_double), ("tanang", C.c_double),
    ]


class XsqRkArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("method", C.c_int32), ("rhs", C.c_int32),
        ("n_state", C.c_int32), ("n_param", C.c_int32),
        ("interpolant", C.c_int32),
        ("n_lanes", C.c_int64),
        ("y0", C.c_void_p), ("params", C.c_void_p),
        ("t0", C.c_double), ("t_bound", C.c_double),
        ("rtol", C.c_double),
        ("atol", _dp), ("n_atol", C.c_int32), ("use_sc_params", C.c_int32),
        ("sc_params", C.c_double * 4),
        ("first_step", C.c_double), ("max_step", C.c_double),
        ("t_eval", C.c_void_p), ("n_eval", C.c_int32),
        ("max_steps", C.c_int32),
        ("y_eval", C.c_void_p),
        ("h_forced", C.c_void_p), ("n_forced", C.c_int32),
        ("reserved0", C.c_int32),
        ("t_final", C.c_void_p), ("y_final", C.c_void_p),
        ("h_next", C.c_void_p),
        ("n_accepted", C.c_void_p), ("n_rejected", C.c_void_p),
        ("nfev", C.c_void_p), ("status", C.c_void_p),
        ("n_eval_done", C.c_void_p),
        ("nfev_stiff_detect", C.c_int32), ("reserved1", C.c_int32),
        ("stiff_flags", C.c_void_p),
        ("events", C.c_int32), ("n_event_fns", C.c_int32),
        ("ev_terminal", _ip), ("ev_direction", _ip),
        ("ev_capacity", C.c_int32), ("reserved2", C.c_int32),
        ("t_events", C.c_void_p), ("y_events", C.c_void_p),
        ("ev_count", C.c_void_p),
        ("first_step_lanes", C.c_void_p),
    ]


RHO_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


class XsqRkcResult(C.Structure):
    _fields_ = [
        ("t_final", C.c_double),
        ("n_accepted", C.c_int32), ("n_rejected", C.c_int32),
        ("nfev", C.c_int32), ("nfesig", C.c_int32), ("maxm", C.c_int32),
        ("status", C.c_int32), ("n_eval_done", C.c_int32),
        ("reserved", C.c_int32), ("kernel_launches", C.c_int64),
    ]


class XsqRkcArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("pde", C.c_int32), ("nx", C.c_int32),
        ("rows_global", C.c_int32), ("rows_local", C.c_int32),
        ("row0", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("u0", C.c_void_p),
        ("t0", C.c_double), ("t_bound", C.c_double),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("first_step", C.c_double), ("max_step", C.c_double),
        ("const_jac", C.c_int32), ("max_steps", C.c_int32),
        ("rho_const", C.c_double), ("rho_cb", RHO_FN), ("rho_user", C.c_void_p),
        ("t_eval", _dp), ("n_eval", C.c_int32), ("reserved", C.c_int32),
        ("u_eval", C.c_void_p), ("u_final", C.c_void_p),
        ("result", C.POINTER(XsqRkcResult)),
        ("pde_params", _dp), ("n_pde_params", C.c_int32),
        ("reserved2", C.c_int32),
    ]


class XsqError(RuntimeError):
    def __init__(self, code, what, detail):
        self.code = code
        super().__init__(f"libxsq: {what} ({code}): {detail}")


_lib = None


def load():
    """Load libxsq.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import "
            "__graft_entry__ as g; g.build()'` or `make -C "
            "extensisq_b200/csrc`.  extensisq_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.xsq_abi_version.restype = C.c_int
    lib.xsq_strerror.restype = C.c_char_p
    lib.xsq_strerror.argtypes = [C.c_int]
    lib.xsq_last_error_detail.restype = C.c_char_p
    lib.xsq_device_info.argtypes = [C.c_int, _ip, _ip, _ip]
    lib.xsq_tableau_load.argtypes = [C.POINTER(XsqTableau)]
    lib.xsq_tableau_get.argtypes = [C.c_int32, C.POINTER(XsqTableau)]
    lib.xsq_rhs_builtin.argtypes = [C.c_char_p, _ip, _ip, _ip]
    lib.xsq_rhs_register_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int32,
                                            C.c_int32, _ip]
    lib.xsq_user_compile_check.argtypes = [C.c_int32, C.c_int32]
    lib.xsq_events_register_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, _ip]
    lib.xsq_events_compile_check.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.xsq_rk_solve.argtypes = [C.POINTER(XsqRkArgs), C.c_void_p]
    lib.xsq_rk_solve_host.argtypes = [C.POINTER(XsqRkArgs), C.c_int]
    lib.xsq_swag_solve.argtypes = [C.POINTER(XsqRkArgs), C.c_int32, C.c_void_p]
    lib.xsq_comm_unique_id.argtypes = [C.c_char_p]
    lib.xsq_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_char_p,
                                    C.POINTER(C.c_void_p)]
    lib.xsq_comm_destroy.argtypes = [C.c_void_p]
    lib.xsq_pde_register_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int32,
                                            _ip]
    lib.xsq_pde_register_vector_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, _ip]
    lib.xsq_rkc_solve.argtypes = [C.POINTER(XsqRkcArgs), C.c_void_p,
                                  C.c_void_p]
    lib.xsq_rkc_stage_bench.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp,
                                        C.c_void_p]
    lib.xsq_rkc_stage_bench_tma.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _dp, C.c_void_p]
    lib.xsq_launch_count.restype = C.c_int64
    lib.xsq_trim_memory.argtypes = [C.c_int]
    lib.xsq_profile_enable.argtypes = [C.c_int]
    lib.xsq_profile_last.argtypes = [_dp, _dp, _dp]
    lib.xsq_profile_get.argtypes = [C.c_int, _dp, _dp, _dp]
    lib.xsq_launch_count.argtypes = [C.c_int]
    lib.xsq_fp64_peak.argtypes = [C.c_int, C.c_int32, _dp]
    if lib.xsq_abi_version() != 2:
        raise ImportError("libxsq.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    lib = load()
    what = lib.xsq_strerror(rc).decode()
    detail = lib.xsq_last_error_detail().decode()
    if rc == -1:
        # argument errors surface as the reference's exception type
        raise ValueError(detail or what)
    raise XsqError(rc, what, detail)
