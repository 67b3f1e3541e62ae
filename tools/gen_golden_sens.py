#!/usr/bin/env python
"""Dev-time tool: golden vectors for forward sensitivities (SURVEY.md section
8f, rank 4): the UNMODIFIED ``extensisq.sens_forward`` (sensitivity.py:60-217)
with explicit methods of the reference, stored in tests/golden/sens_golden.json.

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_sens.py
"""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import extensisq as ref                      # noqa: E402
from oracle.sens_oracle import PROBLEMS      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "sens_golden.json")
CASES = [
    # id, problem, method, t_span, y0, p, dy0dp (None = zeros), rtol, atol, t_eval
    ("rob_BS5", "robertson", "BS5", (0., 0.4), [1., 0., 0.], [0.04, 1e4, 3e7], None, 1e-4, [1e-8, 1e-14, 1e-6], None),
    ("rob_Ts5", "robertson", "Ts5", (0., 0.4), [1., 0., 0.], [0.04, 1e4, 3e7], None, 1e-5, [1e-8, 1e-14, 1e-6], None),
    ("lor_Pr8", "lorenz", "Pr8", (0., 1.5), [1., 1., 1.], [10., 28., 8. / 3.], None, 1e-8, 1e-10, None),
    ("lor_CK5_dy0", "lorenz", "CK5", (0., 1.0), [1., 1., 1.], [10., 28., 0.0], [[1., 0., 0.], [0., .5, 0.], [0., 0., 0.]], 1e-6, 1e-8, None),
    ("vdp_BS5_teval", "vanderpol", "BS5", (0., 5.0), [2., 0.], [3.0], None, 1e-6, 1e-9, (0., 5., 11)),
]


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=float).reshape(-1)]


def main():
    out = []
    for cid, prob, m, span, y0, p, d0, rtol, atol, te in CASES:
        fun, jac, dfdp, _ = PROBLEMS[prob]
        ny, npar = len(y0), len(p)
        dy0dp = np.zeros((ny, npar)) if d0 is None else np.array(d0)
        a = np.array(atol) if isinstance(atol, list) else atol
        t_eval = np.linspace(*te) if te else None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sens, yf, sol = ref.sens_forward(fun, span, np.array(y0), jac, dfdp, dy0dp, p,
                                             atol=a, rtol=rtol, method=getattr(ref, m),
                                             t_eval=t_eval)
        out.append(dict(id=cid, problem=prob, method=m, t_span=span, y0=y0, p=p,
                        dy0dp=dy0dp.tolist(), rtol=rtol, atol=atol, t_eval=list(te) if te else None,
                        nfev=int(sol.nfev), n_t=int(sol.t.size), sens=hx(sens), yf=hx(yf),
                        y=hx(sol.y), y_shape=list(sol.y.shape)))
        print(cid, sol.nfev, sol.t.size, np.abs(sens).max())
    with open(OUT, "w") as fh:
        json.dump(dict(reference_version=ref.__version__, cases=out), fh)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
