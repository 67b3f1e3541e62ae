#!/usr/bin/env python
"""Dev-time tool: read the coefficient attributes of the reference's
Runge-Kutta-Nystrom classes (Fi4N, Fi5N: extensisq/fine.py; Mu5Nmb:
extensisq/murua.py; MR6NN: extensisq/mikkawy.py; base class
extensisq/common.py:1207-1320) and write them, hex-float, to
``extensisq_b200/data/tableaux_rkn.json``.  Only numbers travel.

``E`` / ``Ep`` of Mu5Nmb are stored as the class holds them; the reference
scales them by 0.75 at construction (`scale_embedded=True`, murua.py:224-227),
which the loaders repeat in floating point.

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_tableaux_rkn.py
Then: python tools/gen_header.py
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import extensisq as ref  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "extensisq_b200", "data", "tableaux_rkn.json")


def hexarr(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return [float(x).hex() for x in a]
    return [[float(x).hex() for x in row] for row in a]


def nystrom(cls, src):
    s = int(cls.n_stages)
    has_ap = cls.Ap is not NotImplemented
    d = dict(
        name=cls.__name__, source=src, n_stages=s, order=int(cls.order),
        order_secondary=int(cls.order_secondary), sc_params=cls.sc_params,
        stbre=(None if cls.stbre is NotImplemented else float(cls.stbre)),
        stbim=(None if cls.stbim is NotImplemented else float(cls.stbim)),
        tanang=(None if cls.tanang is NotImplemented else float(cls.tanang)),
        velocity_dependent=bool(has_ap),
        A=hexarr(cls.A), Ap=hexarr(cls.Ap if has_ap else np.zeros((s, s))),
        B=hexarr(cls.B), Bp=hexarr(cls.Bp), C=hexarr(cls.C),
        E=hexarr(cls.E), Ep=hexarr(cls.Ep),
        embedded_scale=(0.75 if cls.__name__ == "Mu5Nmb" else 1.0),
    )
    assert cls.A.shape == (s, s) and cls.B.shape == (s,) and cls.Bp.shape == (s,)
    assert cls.E.shape == (s + 1,) and cls.Ep.shape == (s + 1,)
    return d


def main():
    tabs = {
        "Fi4N": nystrom(ref.Fi4N, "extensisq/fine.py:89-113"),
        "Fi5N": nystrom(ref.Fi5N, "extensisq/fine.py:222-256"),
        "Mu5Nmb": nystrom(ref.Mu5Nmb, "extensisq/murua.py:105-172"),
        "MR6NN": nystrom(ref.MR6NN, "extensisq/mikkawy.py:87-121"),
    }
    with open(OUT, "w") as fh:
        json.dump({"reference_version": ref.__version__, "tableaux": tabs}, fh, indent=1)
    print("wrote", OUT, {k: (v["n_stages"], v["velocity_dependent"]) for k, v in tabs.items()})


if __name__ == "__main__":
    main()
