"""Forward sensitivities: the restatement of
the reference's sens_forward (sensitivity.py:60-217) on the restated solvers,
against golden vectors of the unmodified reference."""
import json
import os

import numpy as np
import pytest

from oracle import rk_oracle as RO
from oracle import sens_oracle as SO

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "sens_golden.json")) as fh:
    CASES = json.load(fh)["cases"]
TABS = RO.load_tableaux()


def unhex(a):
    return np.array([float.fromhex(v) for v in a])


def case_args(c):
    atol = np.array(c["atol"]) if isinstance(c["atol"], list) else c["atol"]
    te = np.linspace(*c["t_eval"][:2], int(c["t_eval"][2])) if c["t_eval"] else None
    return atol, te


@pytest.mark.parametrize("c", CASES, ids=lambda c: c["id"])
def test_sens_oracle_bit_identical_to_reference(c):
    fun, jac, dfdp, _ = SO.PROBLEMS[c["problem"]]
    atol, te = case_args(c)
    sens, yf, sol = SO.sens_forward(TABS[c["method"]], fun, c["t_span"], c["y0"], jac, dfdp,
                                    np.array(c["dy0dp"]), c["p"], atol=atol, rtol=c["rtol"],
                                    t_eval=te)
    assert sol["nfev"] == c["nfev"]
    assert np.array_equal(sens.reshape(-1), unhex(c["sens"]))
    assert np.array_equal(yf, unhex(c["yf"]))
    assert np.array_equal(sol["y"].reshape(-1), unhex(c["y"]))


def test_reference_test_values_for_robertson():
    # reference tests/test_sens.py:58-62 (its own expected numbers, rtol 1e-3)
    c = next(c for c in CASES if c["id"] == "rob_BS5")
    sens = unhex(c["sens"]).reshape(3, 3)
    np.testing.assert_allclose(unhex(c["yf"]), [9.8517e-01, 3.3864e-05, 1.4794e-02], rtol=1e-3)
    np.testing.assert_allclose(sens, [[-3.5595e-01, 9.5428e-08, -1.5832e-11],
                                      [3.9026e-04, -2.1310e-10, -5.2900e-13],
                                      [3.5556e-01, -9.5215e-08, 1.6361e-11]], rtol=1e-3)
