#!/usr/bin/env python
"""Dev-time tool: golden vectors for event detection (SURVEY.md section 8f,
rank 3): the UNMODIFIED reference classes driven by scipy's ``solve_ivp`` with
``events=`` (terminal / counted-terminal / directional / several events,
t_eval, backward integration, every interpolant kind).  Stored losslessly in
``tests/golden/events_golden.json``.

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_events.py
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
from oracle.problems import make_fun, EVENT_SETS   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "events_golden.json")


def hx(a):
    a = np.asarray(a, dtype=float)
    if a.ndim == 2:
        return [[float(v).hex() for v in row] for row in a]
    return [float(v).hex() for v in a]


L = [10., 28., 8. / 3.]
CASES = [
    # id, method, problem, params, y0, t_span, options, t_eval, event set, terminal[], direction[]
    ("ground_Ts5", "Ts5", "ballistic", [], [10., 5.], [0., 5.], {}, None, "ground", [1, 0], [-1, 0]),
    ("ground_BS5low", "BS5", "ballistic", [], [10., 5.], [0., 5.], dict(rtol=1e-6, atol=1e-9), None, "ground", [1, 0], [-1, 0]),
    ("ground_BS5best", "BS5", "ballistic", [], [10., 5.], [0., 5.], dict(interpolant="best"), (0., 5., 26), "ground", [1, 0], [-1, 0]),
    ("ground_CKdisc", "CKdisc", "ballistic", [], [10., 5.], [0., 5.], dict(rtol=1e-5, atol=1e-8), None, "ground", [1, 0], [0, 0]),
    ("ground_Me4_noterm", "Me4", "ballistic", [], [10., 5.], [0., 3.], {}, None, "ground", [0, 0], [0, 0]),
    ("lorenz_Ts5", "Ts5", "lorenz63", L, [1., 1., 1.], [0., 8.], dict(rtol=1e-6, atol=1e-9), None, "lorenz_sections", [0, 0, 0], [1, 0, -1]),
    ("lorenz_Pr8_count5", "Pr8", "lorenz63", L, [1., 1., 1.], [0., 20.], dict(rtol=1e-8, atol=1e-10), None, "lorenz_sections", [5, 0, 0], [0, 0, 0]),
    ("lorenz_CK5_teval", "CK5", "lorenz63", L, [1., 1., 1.], [0., 6.], dict(rtol=1e-6, atol=1e-9), (0., 6., 61), "lorenz_sections", [0, 3, 0], [1, 1, 0]),
    ("lorenz_Pr9", "Pr9", "lorenz63", L, [-5., -7., 20.], [0., 5.], dict(rtol=1e-7, atol=1e-9), None, "lorenz_sections", [0, 0, 0], [-1, -1, 1]),
    ("lorenz_back_CFMR", "CFMR7osc", "lorenz63", L, [-5., -7., 20.], [1., 0.], dict(rtol=1e-6, atol=1e-9), (1., 0., 11), "lorenz_sections", [0, 0, 0], [0, 0, 0]),
    ("vdp_Pr7", "Pr7", "vanderpol", [2.0], [2., 0.], [0., 12.], dict(rtol=1e-6, atol=1e-8), None, "vdp_cross", [0, 0, 1], [-1, 0, 0]),
    ("vdp_Ts5_teval_term", "Ts5", "vanderpol", [2.0], [2., 0.], [0., 12.], dict(rtol=1e-5, atol=1e-7), (0., 12., 49), "vdp_cross", [4, 0, 0], [0, 1, 0]),
    ("vdp_CKdisc", "CKdisc", "vanderpol", [5.0], [2., 0.], [0., 12.], dict(rtol=1e-5, atol=1e-7), None, "vdp_cross", [0, 0, 1], [0, 0, 0]),
    # SWAG (shampine.py): events on SwagDenseOutput / LinearDenseOutput
    ("swag_ground", "SWAG", "ballistic", [], [10., 5.], [0., 5.], dict(rtol=1e-6, atol=1e-9), None, "ground", [1, 0], [-1, 0]),
    ("swag_lorenz", "SWAG", "lorenz63", L, [1., 1., 1.], [0., 6.], dict(rtol=1e-6, atol=1e-9), None, "lorenz_sections", [0, 0, 0], [1, 0, -1]),
    ("swag_vdp_teval_term", "SWAG", "vanderpol", [2.0], [2., 0.], [0., 12.], dict(rtol=1e-5, atol=1e-7), (0., 12., 49), "vdp_cross", [4, 0, 0], [0, 1, 0]),
    ("swag_lorenz_back", "SWAG", "lorenz63", L, [-5., -7., 20.], [1., 0.], dict(rtol=1e-6, atol=1e-9), (1., 0., 11), "lorenz_sections", [0, 0, 0], [0, 0, 0]),
]


def main():
    out = []
    for cid, mname, prob, prm, y0, span, opts, te, evset, term, direc in CASES:
        fun = make_fun(prob, prm)
        pyev, _ = EVENT_SETS[evset]
        evs = []
        for g, tr, d in zip(pyev, term, direc):
            f = (lambda t, y, g=g: g(t, y))
            f.terminal = tr
            f.direction = d
            evs.append(f)
        t_eval = np.linspace(*te) if te else None
        refcommon.NFS[()] = 0
        if mname == "SWAG":
            import extensisq.shampine as sh
            sh.NFS[()] = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = solve_ivp(fun, span, y0, method=getattr(ref, mname), t_eval=t_eval,
                          events=evs, **opts)
        out.append(dict(id=cid, method=mname, problem=prob, params=prm, y0=y0, t_span=span,
                        options=opts, t_eval=list(te) if te else None, events=evset,
                        terminal=term, direction=direc, status=int(r.status),
                        nfev=int(r.nfev), nfs=int(sh.NFS) if mname == "SWAG" else int(refcommon.NFS), t=hx(r.t), y=hx(r.y),
                        t_events=[hx(a) for a in r.t_events],
                        y_events=[hx(np.asarray(a).reshape(-1, len(y0))) if len(a) else []
                                  for a in r.y_events]))
        print(cid, r.status, r.nfev, [len(a) for a in r.t_events], r.t.size)
    with open(OUT, "w") as fh:
        json.dump(dict(reference_version=ref.__version__, cases=out), fh)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
