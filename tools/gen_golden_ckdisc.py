#!/usr/bin/env python
"""Dev-time tool: golden vectors for CKdisc.

Runs the UNMODIFIED reference (``extensisq.CKdisc`` through scipy's
``solve_ivp``, imported from /root/reference, this container only) on
non-smooth and smooth problems and stores every accepted (t, y), the dense
output at t_eval, nfev, NFS and the status, losslessly (hex floats), in
``tests/golden/ckdisc_golden.json``.  The right-hand sides are the ones in
``oracle/problems.py`` (same expressions).

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_ckdisc.py
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
import extensisq.common as refcommon         # noqa: E402
from scipy.integrate import solve_ivp        # noqa: E402
from oracle.problems import make_fun         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ckdisc_golden.json")


def hx(a):
    a = np.asarray(a, dtype=float)
    if a.ndim == 2:
        return [[float(v).hex() for v in row] for row in a]
    return [float(v).hex() for v in a]


CASES = [
    # id, problem, params, y0, t_span, options, t_eval (start, stop, num)
    ("f2_default", "detest_f2", [], [110.], [0., 10.], {}, None),
    ("f2_atol1e-3", "detest_f2", [], [110.], [0., 10.], dict(rtol=1e-13, atol=1e-3), None),
    ("f2_atol1e-6", "detest_f2", [], [110.], [0., 10.], dict(rtol=1e-13, atol=1e-6), None),
    ("f2_atol1e-9", "detest_f2", [], [110.], [0., 10.], dict(rtol=1e-13, atol=1e-9), None),
    ("f2_teval", "detest_f2", [], [110.], [0., 10.], dict(rtol=1e-5, atol=1e-7), (0., 10., 57)),
    ("f2_first_step", "detest_f2", [], [110.], [0., 6.], dict(first_step=0.01, max_step=0.4), None),
    ("square_default", "square_forced", [], [1., 0.], [0., 12.], {}, None),
    ("square_tight", "square_forced", [], [1., 0.], [0., 12.], dict(rtol=1e-8, atol=1e-10), None),
    ("square_teval", "square_forced", [], [1., 0.], [0., 12.], dict(rtol=1e-6, atol=1e-8), (0., 12., 101)),
    ("square_atolvec", "square_forced", [], [1., 0.], [0., 8.], dict(rtol=1e-5, atol=[1e-7, 1e-5]), None),
    ("lorenz_default", "lorenz63", [10., 28., 8. / 3.], [1., 1., 1.], [0., 3.], {}, None),
    ("lorenz_tight", "lorenz63", [10., 28., 8. / 3.], [1., 1., 1.], [0., 3.], dict(rtol=1e-8, atol=1e-10), None),
    ("lorenz_back", "lorenz63", [10., 28., 8. / 3.], [-5., -7., 20.], [1., 0.], dict(rtol=1e-6, atol=1e-9), (1., 0., 21)),
    ("vdp_mu2", "vanderpol", [2.0], [2., 0.], [0., 10.], dict(rtol=1e-6, atol=1e-8), None),
    ("vdp_mu20_teval", "vanderpol", [20.0], [2., 0.], [0., 30.], dict(rtol=1e-4, atol=1e-6), (0., 30., 61)),
    ("arenstorf_short", "arenstorf", [0.012277471], [0.994, 0., 0., -2.00158510637908252240537862224],
     [0., 2.], dict(rtol=1e-7, atol=1e-9), None),
]


def main():
    out = []
    for cid, prob, prm, y0, span, opts, te in CASES:
        fun = make_fun(prob, prm)
        o = dict(opts)
        if "atol" in o and isinstance(o["atol"], list):
            o["atol"] = np.array(o["atol"])
        t_eval = np.linspace(*te) if te else None
        refcommon.NFS[()] = 0
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            r = solve_ivp(fun, span, y0, method=ref.CKdisc, t_eval=t_eval, **o)
        out.append(dict(id=cid, problem=prob, params=prm, y0=y0, t_span=span,
                        options=opts, t_eval=list(te) if te else None,
                        status=int(r.status), nfev=int(r.nfev),
                        nfs=int(refcommon.NFS), t=hx(r.t), y=hx(r.y)))
        print(cid, r.status, r.nfev, int(refcommon.NFS), r.t.size)
    with open(OUT, "w") as fh:
        json.dump(dict(reference_version=ref.__version__, cases=out), fh)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
