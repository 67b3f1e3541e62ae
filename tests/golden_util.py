"""Shared helpers for the golden-vector tests (tests/golden/rk_golden.npz,
produced from the unmodified reference by tools/gen_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "rk_golden.npz")

BUILTIN_PROBLEMS = {"lorenz63", "vanderpol", "arenstorf"}


class Golden:
    def __init__(self):
        self.z = np.load(GOLDEN)
        meta = json.loads(str(self.z["__meta__"]))
        self.cases = meta["cases"]
        self.by_id = {c["id"]: c for c in self.cases}

    def arr(self, cid, name):
        return self.z[f"{cid}/{name}"]


def case_options(c):
    """solver options of a golden case; like the reference, stiffness
    diagnosis defaults to every 5000 evaluations when not given."""
    opts = dict(c.get("options", {}))
    opts.setdefault("nfev_stiff_detect", 5000)
    if "atol_vec" in c:
        opts["atol"] = np.array(c["atol_vec"])
    if "sc_params" in opts:
        opts["sc_params"] = tuple(opts["sc_params"])
    return opts


def case_span(c):
    return [float(x) for x in c["t_span"]]


def case_t_eval(c):
    return np.linspace(*c["t_eval"]) if c.get("t_eval") else None


def stability_limited(c):
    """Cases whose step size is limited by stability, not accuracy: many
    rejections, and accept/reject decisions that flip under 1-ulp changes of
    the error norm (hard part 2).  For these only
    approximate count parity can hold between different summation orders."""
    n_acc = max(c["n_t"] - 1, 1) if not c.get("t_eval") else None
    if c["problem"] == "vanderpol" and c["params"][0] >= 10.0:
        return True
    return False
