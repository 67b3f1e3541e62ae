#!/usr/bin/env python
"""Dev-time tool: read the tableau class attributes of the reference package
and write them (hex-float, lossless) to ``extensisq_b200/data/tableaux.json``.

Only numbers travel: the reference's classes are imported from
``/root/reference`` (read-only, this container only), their public data
attributes ``n_stages, order, order_secondary, A, B, C, E, P, stbrad, tanang,
sc_params`` (``extensisq/common.py:88-121``) are read and serialised.  SURVEY.md
§9: "read tableaux from the imported classes, don't retype them" (Ts5's first
column of A is derived at import, ``tsitouras.py:100``).

Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_tableaux.py
Then: python tools/gen_header.py       (json -> CUDA constexpr header)
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import extensisq as ref  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "extensisq_b200", "data", "tableaux.json")


def hexarr(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return [float(x).hex() for x in a]
    return [[float(x).hex() for x in row] for row in a]


def generic(cls, src):
    d = dict(
        name=cls.__name__, source=src,
        n_stages=int(cls.n_stages), order=int(cls.order),
        order_secondary=int(cls.order_secondary),
        sc_params=cls.sc_params,
        stbrad=float(cls.stbrad), tanang=float(cls.tanang),
        A=hexarr(cls.A), B=hexarr(cls.B), C=hexarr(cls.C), E=hexarr(cls.E),
        P=hexarr(cls.P),
    )
    assert cls.A.shape == (cls.n_stages, cls.n_stages)
    assert cls.B.shape == (cls.n_stages,) and cls.C.shape == (cls.n_stages,)
    assert cls.E.shape == (cls.n_stages + 1,)
    assert cls.P.shape[0] == cls.n_stages + 1
    return d


def main():
    tabs = {}
    tabs["Ts5"] = generic(ref.Ts5, "extensisq/tsitouras.py:83-115")
    tabs["CK5"] = generic(ref.CK5, "extensisq/cash.py:82-112")
    tabs["Me4"] = generic(ref.Me4, "extensisq/merson.py:82-122")
    tabs["Pr7"] = generic(ref.Pr7, "extensisq/prince.py:79-128")
    tabs["Pr8"] = generic(ref.Pr8, "extensisq/prince.py:205-372")
    tabs["Pr9"] = generic(ref.Pr9, "extensisq/prince.py:449-746")
    tabs["CFMR7osc"] = generic(ref.CFMR7osc, "extensisq/calvo.py:89-149")
    bs5 = generic(ref.BS5, "extensisq/bogacki.py:103-215")
    bs5.update(
        E_pre=hexarr(ref.BS5.E_pre), B_scale_pre=hexarr(ref.BS5.B_scale_pre),
        C_extra=hexarr(ref.BS5.C_extra), A_extra=hexarr(ref.BS5.A_extra),
        Plow=hexarr(ref.BS5.Plow), Pbest=hexarr(ref.BS5.Pbest),
        n_extra_stages=int(ref.BS5.n_extra_stages))
    tabs["BS5"] = bs5
    # CKdisc (cash.py:184-236): CK5's A and C, its own E (B_all[5] - B_all[4]
    # evaluated in floating point), assessment and fallback weights.  Kept out
    # of "tableaux": it is not a RungeKutta._step_impl method.
    ck = ref.CKdisc
    ckdisc = dict(
        name="CKdisc", source="extensisq/cash.py:184-236",
        n_stages=int(ck.n_stages), order=int(ck.order),
        order_secondary=int(ck.order_secondary), sc_params="standard",
        max_factor=float(ck.max_factor), min_factor=float(ck.min_factor),
        safety=0.9,                                   # cash.py:6
        A=hexarr(ck.A), B=hexarr(ck.B), C=hexarr(ck.C), E=hexarr(ck.E),
        P=hexarr(ck.P), B_assess=hexarr(ck.B_assess),
        E_assess=hexarr(ck.E_assess), C_fallback=hexarr(ck.C_fallback),
        B_fallback=hexarr(ck.B_fallback), E_fallback=hexarr(ck.E_fallback))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        json.dump(dict(reference_version=ref.__version__, tableaux=tabs,
                       ckdisc=ckdisc), fh, indent=0)
    print("wrote", OUT, {k: v["n_stages"] for k, v in tabs.items()})


if __name__ == "__main__":
    main()
