"""CPU-side checks of the C-ABI boundary: libxsq.so loads, exports every symbol
include/xsq.h declares, the ctypes mirrors have the C struct sizes, NVRTC
accepts the specialised kernels, and compute entry points FAIL LOUDLY without a
device (the product has no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

from extensisq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xsq.h")
NO_GPU = not torch.cuda.is_available()


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(xsq_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"libxsq.so does not export {s}"
    assert set(_lib.EXPORTS) == set(syms)
    assert lib.xsq_abi_version() == 2


def test_struct_sizes_match_the_c_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "xsq.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu\\n", sizeof(xsq_rk_args_t), '
                   'sizeof(xsq_tableau_t), sizeof(xsq_rkc_args_t), '
                   'sizeof(xsq_rkc_result_t)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    a, t, ra, rr = map(int, subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(_lib.XsqRkArgs)
    assert t == C.sizeof(_lib.XsqTableau)
    assert ra == C.sizeof(_lib.XsqRkcArgs)
    assert rr == C.sizeof(_lib.XsqRkcResult)


def test_strerror_and_builtin_rhs_lookup():
    lib = _lib.load()
    assert lib.xsq_strerror(0) == b"ok"
    assert b"argument" in lib.xsq_strerror(-1)
    h, n, p = C.c_int32(), C.c_int32(), C.c_int32()
    for name, shape in (("lorenz63", (3, 3)), ("vanderpol", (2, 1)),
                        ("arenstorf", (4, 1)), ("nbody32", (192, 33))):
        assert lib.xsq_rhs_builtin(name.encode(), C.byref(h), C.byref(n),
                                   C.byref(p)) == 0
        assert (n.value, p.value) == shape
    assert lib.xsq_rhs_builtin(b"nope", C.byref(h), None, None) == -1


def test_nvrtc_accepts_user_rhs_and_user_tableau():
    """The specialised kernel for a user RHS / user tableau compiles for
    sm_100a (NVRTC needs no device)."""
    lib = _lib.load()
    h = C.c_int32()
    src = (b"__device__ void rhs(double t, const double* y, const double* p,"
           b" double* dy) { dy[0] = y[1]; dy[1] = -p[0] * y[0]; }")
    assert lib.xsq_rhs_register_source(src, b"rhs", 2, 1, C.byref(h)) == 0
    assert h.value >= _lib.XSQ_RHS_USER_BASE
    assert lib.xsq_user_compile_check(0, h.value) == 0, \
        lib.xsq_last_error_detail().decode()
    bad = C.c_int32()
    assert lib.xsq_rhs_register_source(b"__device__ void rhs(double t) { oops }",
                                       b"rhs", 2, 0, C.byref(bad)) == 0
    assert lib.xsq_user_compile_check(0, bad.value) == -3
    assert b"NVRTC" in lib.xsq_last_error_detail()
    # Heun (docs/Demo_own_RK.ipynb) as a user tableau with a built-in rhs
    t = _lib.XsqTableau()
    t.n_stages, t.order, t.order_secondary, t.n_poly = 2, 2, 1, 0
    t.A[1][0] = 1.0
    t.B[0] = t.B[1] = 0.5
    t.C[1] = 1.0
    t.E[0], t.E[1], t.E[2] = 0.5, -0.5, 0.0
    for i, v in enumerate((1.0, 0.0, 0.0, 0.9)):
        t.sc_params[i] = v
    assert lib.xsq_tableau_load(C.byref(t)) == 0
    assert lib.xsq_user_compile_check(_lib.XSQ_METHOD_USER, 0) == 0, \
        lib.xsq_last_error_detail().decode()


@pytest.mark.skipif(not NO_GPU, reason="only meaningful without a device")
def test_compute_fails_loudly_without_a_device():
    import extensisq_b200 as xb
    lib = _lib.load()
    assert lib.xsq_device_info(0, None, None, None) == -2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        xb.solve_ivp_batched("lorenz63", (0.0, 1.0), [[1.0, 1.0, 1.0]],
                             xb.Ts5, params=[[10.0, 28.0, 8 / 3]])
    tf = C.c_double()
    assert lib.xsq_fp64_peak(0, 10, C.byref(tf)) == -2


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libxsq.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "extensisq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt
                assert "libxsq_oracle" not in txt


def test_event_kernels_compile_for_sm100a():
    """NVRTC accepts the (method, rhs, events) translation units (no device
    needed): the event machinery of xsq_rk_core.cuh behind XSQ_EVENTS_N."""
    import ctypes as C
    from extensisq_b200 import _lib
    from oracle.problems import EVENT_SETS
    lib = _lib.load()
    py, src = EVENT_SETS["lorenz_sections"]
    h = C.c_int32()
    assert lib.xsq_events_register_source(src.encode(), b"event", len(py), C.byref(h)) == 0
    for method in (_lib.METHOD_IDS["Ts5"], _lib.METHOD_IDS["BS5"], _lib.METHOD_IDS["CKdisc"]):
        rc = lib.xsq_events_compile_check(method, 0, h.value)
        if rc == -4 and b"libnvrtc" in lib.xsq_last_error_detail():
            pytest.skip("libnvrtc not available")
        assert rc == 0, lib.xsq_last_error_detail().decode()
    assert lib.xsq_events_register_source(src.encode(), b"event", 9, C.byref(h)) == -1
