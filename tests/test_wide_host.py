"""Host-side checks of the warp-per-system path for user right-hand sides with
n_state > 16 (no GPU needed): the generated translation unit compiles with
NVRTC, the argument limits, and the C oracle's warp-strided summation order."""
import ctypes as C

import numpy as np
import pytest

from extensisq_b200 import _lib
from oracle import c_oracle as CO
from oracle import rk_oracle as O

SRC = """
__device__ double chain(int i, double t, const double* y, const double* p) {
    const double l = i > 0 ? y[i - 1] : 0.0, r = i < %d - 1 ? y[i + 1] : 0.0;
    return p[0] * (l - 2.0 * y[i] + r);
}
"""


def register(n, n_param=1):
    lib = _lib.load()
    h = C.c_int32()
    rc = lib.xsq_rhs_register_source((SRC % n).encode(), b"chain", n, n_param, C.byref(h))
    return rc, h.value


@pytest.mark.parametrize("n", [17, 100, 192])
def test_wide_user_rhs_compiles_into_the_persistent_kernels(n):
    lib = _lib.load()
    rc, h = register(n)
    assert rc == 0
    rc = lib.xsq_user_compile_check(0, h)                   # Ts5 pair
    assert rc == 0


def test_wide_user_rhs_limits():
    assert register(1024)[0] == 0
    assert register(1025)[0] != 0
    assert register(64, n_param=17)[0] != 0      # parameters live in registers
    assert register(16, n_param=17)[0] == 0      # lane-per-system: unchanged


def test_oracle_warp_strided_order_is_a_reordering_only():
    """Same trajectory up to the rounding of the norms: equal counts on a smooth
    problem, states within 1e-11; and the default order is restored."""
    tabs = O.load_tableaux()
    n = 100

    def f(t, y, p):
        l = np.concatenate(([0.0], y[:-1]))
        r = np.concatenate((y[1:], [0.0]))
        return p[0] * (l - 2.0 * y + r)

    y0 = np.sin(np.pi * (np.arange(n) + 1.0) / (n + 1.0))[None, :] * np.array([[1.0], [2.0]])
    prm = np.array([[30.0], [40.0]])
    kw = dict(rtol=1e-6, atol=1e-9, user_fn=f, user_fn_params=True)
    a = CO.rk_batch(tabs["Ts5"], None, (0.0, 1.0), y0, params=prm, **kw)
    with CO.device_math(warp_strided=True):
        b = CO.rk_batch(tabs["Ts5"], None, (0.0, 1.0), y0, params=prm, **kw)
    c = CO.rk_batch(tabs["Ts5"], None, (0.0, 1.0), y0, params=prm, **kw)
    assert np.array_equal(a["y_final"], c["y_final"])          # mode restored
    assert np.array_equal(a["n_accepted"], b["n_accepted"])
    assert np.abs(a["y_final"] - b["y_final"]).max() <= 1e-7       # rtol = 1e-6
