"""CKdisc: the NumPy restatement of
cash.py:115-416 against golden vectors produced by the unmodified reference
(tools/gen_golden_ckdisc.py)."""
import json
import os

import numpy as np
import pytest

from oracle import rk_oracle as RO
from oracle.problems import make_fun

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ckdisc_golden.json")) as fh:
    CASES = json.load(fh)["cases"]
TAB = RO.load_ckdisc()


def unhex(a):
    return np.array([[float.fromhex(v) for v in row] for row in a]
                    if a and isinstance(a[0], list)
                    else [float.fromhex(v) for v in a])


def ck_options(c):
    o = dict(c["options"])
    if isinstance(o.get("atol"), list):
        o["atol"] = np.array(o["atol"])
    return o


def ck_t_eval(c):
    return np.linspace(*c["t_eval"][:2], int(c["t_eval"][2])) if c["t_eval"] else None


def test_coefficients_are_the_reference_class_attributes():
    # cash.py:199-236: E = B_all[5] - B_all[4] in floating point; fallback
    # solutions advance a fraction c of the step
    assert TAB.n_stages == 6 and TAB.order == 5 and TAB.order_secondary == 4
    assert TAB.max_factor == 5 and TAB.min_factor == 1 / 5 and TAB.safety == 0.9
    assert np.array_equal(TAB.C_fallback, TAB.C[[1, 3]])
    assert TAB.E[-1] == 0.0
    for B, c in zip(TAB.B_fallback, TAB.C_fallback):
        assert abs(B.sum() - c) < 1e-15                 # consistency
    assert abs(TAB.B.sum() - 1) < 1e-15
    for Bh in TAB.B_assess:
        assert abs(Bh.sum() - 1) < 1e-15
    # order conditions up to the nominal order of each embedded solution
    assert abs(TAB.B_assess[0] @ TAB.C - 1 / 2) < 1e-15            # order 2
    assert abs(TAB.B_assess[1] @ TAB.C ** 2 - 1 / 3) < 1e-15       # order 3
    assert abs(TAB.B @ TAB.C ** 4 - 1 / 5) < 1e-15                 # order 5


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_numpy_oracle_is_bit_identical_to_the_reference(c):
    r = RO.rk_solve(TAB, make_fun(c["problem"], c["params"]), c["t_span"], c["y0"],
                    t_eval=ck_t_eval(c), **ck_options(c))
    assert r["status"] == c["status"]
    assert r["nfev"] == c["nfev"]
    assert r["n_rejected"] == c["nfs"]
    assert np.array_equal(r["t"], unhex(c["t"]))
    assert np.array_equal(r["y"], unhex(c["y"]))


def test_both_interpolants_are_exercised():
    # the golden set must cover the Horner (order 5 accepted) and the cubic
    # (fallback accepted) dense output, cash.py:406-416
    c = next(c for c in CASES if c["id"] == "f2_teval")
    st_orders = set()
    fun = make_fun(c["problem"], c["params"])
    st = RO.RKState(TAB, fun, 0.0, c["y0"], 10.0, nfev_stiff_detect=0, **ck_options(c))
    while st.t < 10.0:
        ok, _ = RO.ckdisc_step(st)
        assert ok
        st_orders.add(st.order_accepted)
    assert {1, 2, 4} <= st_orders or {2, 4} <= st_orders


def test_c_restatement_against_the_reference_goldens():
    """ck_solve_one (oracle/xsq_oracle.c) in the reference's arithmetic (sqrt,
    pow, true division; sums in index order with fma instead of dgemv order):
    the same nfev / NFS / accepted steps as the unmodified reference on the
    golden cases that are not sensitive to last-digit changes, states close on
    every case with the same step sequence.  The fraction is printed."""
    from oracle import c_oracle as CO
    same = 0
    for c in CASES:
        fun = make_fun(c["problem"], c["params"])
        o = CO.rk_batch(TAB, None, c["t_span"], [c["y0"]], t_eval=ck_t_eval(c),
                        user_fn=fun, **ck_options(c))
        t_g, y_g = unhex(c["t"]), unhex(c["y"])
        assert int(o["status"][0]) == c["status"]
        eq = int(o["nfev"][0]) == c["nfev"] and int(o["n_rejected"][0]) == c["nfs"]
        same += eq
        if eq:
            # rounding differences grow along the trajectory like the method's
            # own error: compare at a multiple of the requested tolerance
            opt = ck_options(c)
            tol = dict(rtol=max(1e-9, 100 * opt.get("rtol", 1e-3)),
                       atol=max(1e-11, 100 * float(np.max(opt.get("atol", 1e-6)))))
            if ck_t_eval(c) is None:
                assert int(o["n_accepted"][0]) == len(t_g) - 1
                assert np.allclose(o["y_final"][0], y_g[:, -1], **tol), c["id"]
            else:
                assert np.allclose(o["y"][0], y_g, **tol), c["id"]
        else:
            assert abs(int(o["nfev"][0]) - c["nfev"]) <= 0.12 * c["nfev"] + 6
    print(f"\nC restatement of CKdisc: identical nfev/NFS on {same} of {len(CASES)} golden cases")
    assert same >= len(CASES) // 2
