"""The NumPy restatement of RungeKuttaNystrom (oracle/rk_oracle.py RKNState,
reference common.py:1207-1486; Fi4N, Fi5N: fine.py, Mu5Nmb: murua.py, MR6NN:
mikkawy.py) against golden vectors produced by the unmodified reference
(tools/gen_golden_rkn.py): every accepted t and y, nfev, status and the
stiffness diagnosis (rectangular stability domain, common.py:1322-1486)."""
import json
import os

import numpy as np
import pytest

from oracle import rk_oracle as O
from oracle.problems_rkn import make_fun

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "rkn_golden.json")) as fh:
    CASES = json.load(fh)["cases"]
TABS = O.load_tableaux_rkn()


def unhex(a):
    if a and isinstance(a[0], list):
        return np.array([[float.fromhex(v) for v in r] for r in a])
    return np.array([float.fromhex(v) for v in a])


def test_tableaux_match_the_reference_shapes():
    assert set(TABS) == {"Fi4N", "Fi5N", "Mu5Nmb", "MR6NN"}
    for t in TABS.values():
        s = t.n_stages
        assert t.A.shape == t.Ap.shape == (s, s)
        assert t.B.shape == t.Bp.shape == t.C.shape == (s,)
        assert t.E.shape == t.Ep.shape == (s + 1,)
    assert not TABS["MR6NN"].velocity_dependent and not TABS["MR6NN"].Ap.any()


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_rkn_restatement_is_bit_identical_to_the_reference(c):
    if c["n_t"] > 5000 and c["method"] != "Fi5N":
        pytest.skip("long stiff run: one method is enough for the CPU suite")
    opt = dict(c["options"])
    r = O.rk_solve(TABS[c["method"]], make_fun(c["problem"], c["params"]), c["t_span"],
                   unhex(c["y0"]), **opt)
    assert r["status"] == c["status"]
    assert r["nfev"] == c["nfev"]
    assert r["t"].size == c["n_t"]
    t, y = unhex(c["t"]), unhex(c["y"])
    k = t.size
    assert np.array_equal(r["t"][-k:], t)
    assert np.array_equal(r["y"][:, -k:], y)
    assert r["stiff_flags"] == c["flags"]


def test_wrong_problems_are_refused():
    """reference tests/test_rkn.py::test_wrong_problem"""
    with pytest.raises(AssertionError):                # odd size
        O.rk_solve(TABS["Fi4N"], lambda t, y: np.array([y[1], -y[0], 0.0]), (0, 1), [0., 1., 2.])
    with pytest.raises(AssertionError):                # f[:n] is not the velocity
        O.rk_solve(TABS["Fi4N"], lambda t, y: np.array([-y[0], y[1]]), (0, 1), [0.5, 1.])
    with pytest.raises(AssertionError):                # velocity dependent with MR6NN
        O.rk_solve(TABS["MR6NN"], lambda t, y: np.array([y[1], -y[0] - y[1]]), (0, 1), [0., 1.])


def test_c_restatement_in_device_arithmetic_against_the_reference_goldens():
    """rkn_stage and the Nystrom solution / error block of oracle/xsq_oracle.c (the
    kernels' arithmetic; the GPU is bit-identical to it, tests/test_gpu_rkn.py)
    against the reference's golden runs: status identical, and on every case where
    the other arithmetic takes the same number of steps the final state agrees to
    a multiple of the tolerance.  The fraction is printed and asserted."""
    from oracle import c_oracle as CO
    same = total = 0
    for c in CASES:
        if c["n_t"] > 5000 or c["problem"] == "nbody32":
            continue                    # long stiff runs / the 192-state callback: CPU time
        opt = dict(c["options"])
        opt.pop("nfev_stiff_detect", None)
        fun = make_fun(c["problem"], c["params"])
        with CO.device_math():
            o = CO.rk_batch(TABS[c["method"]], None, c["t_span"], [unhex(c["y0"])], user_fn=fun,
                            nfev_stiff_detect=0, **opt)
        assert int(o["status"][0]) == c["status"], c["id"]
        total += 1
        # the golden nfev may include the (host-only) stiffness probes: compare steps
        t, y = unhex(c["t"]), unhex(c["y"])
        if int(o["n_accepted"][0]) == c["n_t"] - 1:
            same += 1
            tol = 100 * (opt.get("atol", 1e-6) + opt.get("rtol", 1e-3) * np.abs(y[:, -1]))
            assert (np.abs(o["y_final"][0] - y[:, -1]) <= tol).all(), c["id"]
    print(f"\nC restatement of the Nystrom methods: same number of steps as the reference on "
          f"{same} of {total} golden cases")
    assert same >= 0.8 * total
