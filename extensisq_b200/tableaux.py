"""Tableau classes with the reference's names and class attributes.

Each class carries ``n_stages, order, order_secondary, A, B, C, E, P, stbrad,
tanang, sc_params`` exactly as the reference's classes do
(``extensisq/common.py:88-121``; data read from the reference at dev time by
``tools/gen_tableaux.py`` and stored lossless in ``data/tableaux.json``).  They
are passed as ``method=`` to :func:`extensisq_b200.solve_ivp_batched`, the way
the reference's classes are passed to ``scipy.integrate.solve_ivp``
(``README.md:26-35``).

User-defined methods subclass :class:`RungeKutta` and define the same
attributes (``docs/Demo_own_RK.ipynb``); they are uploaded to the device with
``xsq_tableau_load`` and compiled into a specialised kernel on first use.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_JSON = os.path.join(_HERE, "data", "tableaux.json")

MIN_FACTOR = 0.2        # common.py:18
MAX_FACTOR = 4.0        # common.py:19
MAX_FACTOR0 = 10        # common.py:20


def _unhex(a):
    if isinstance(a[0], list):
        return np.array([[float.fromhex(x) for x in r] for r in a])
    return np.array([float.fromhex(x) for x in a])


class RungeKutta:
    """Base of all explicit tableau methods (reference: common.py:69-121).

    Unlike the reference's class this is not a scipy ``OdeSolver``: the
    stepping loop lives in the persistent CUDA kernel, so the class is the
    method *description*; see :class:`extensisq_b200.batched.BatchedSolver`
    for the object that holds device state."""
    n_stages: int = NotImplemented
    order: int = NotImplemented
    order_secondary: int = NotImplemented
    A: np.ndarray = NotImplemented
    B: np.ndarray = NotImplemented
    C: np.ndarray = NotImplemented
    E: np.ndarray = NotImplemented
    P: np.ndarray = NotImplemented
    stbrad: float = NotImplemented
    tanang: float = NotImplemented
    sc_params = "standard"
    max_factor = MAX_FACTOR0
    min_factor = MIN_FACTOR
    _xsq_method = None          # built-in id, None for user tableaux

    @classmethod
    def validate(cls):
        """Shape checks a user tableau must satisfy (common.py:98-111)."""
        s = cls.n_stages
        if not isinstance(s, int) or s < 1:
            raise ValueError("n_stages must be a positive int")
        for name, shape in (("A", (s, s)), ("B", (s,)), ("C", (s,)),
                            ("E", (s + 1,))):
            a = getattr(cls, name)
            if not isinstance(a, np.ndarray) or a.shape != shape:
                raise ValueError(f"{cls.__name__}.{name} must have shape "
                                 f"{shape}")
        if isinstance(cls.P, np.ndarray) and cls.P.shape[0] != s + 1:
            raise ValueError(f"{cls.__name__}.P must have {s + 1} rows")
        if np.any(np.triu(cls.A) != 0):
            raise ValueError("explicit methods need a strictly lower "
                             "triangular A")


def _make(name, d):
    attrs = dict(
        __doc__=f"{name} tableau; reference {d['source']}.",
        n_stages=d["n_stages"], order=d["order"],
        order_secondary=d["order_secondary"], sc_params=d["sc_params"],
        stbrad=d["stbrad"], tanang=d["tanang"],
        A=_unhex(d["A"]), B=_unhex(d["B"]), C=_unhex(d["C"]),
        E=_unhex(d["E"]), P=_unhex(d["P"]),
    )
    for k in ("E_pre", "B_scale_pre", "C_extra", "A_extra", "Plow", "Pbest"):
        if k in d:
            attrs[k] = _unhex(d[k])
    if "n_extra_stages" in d:
        attrs["n_extra_stages"] = d["n_extra_stages"]
    for v in attrs.values():
        if isinstance(v, np.ndarray):
            v.setflags(write=False)
    return type(name, (RungeKutta,), attrs)


with open(_JSON) as _fh:
    _raw = json.load(_fh)
REFERENCE_VERSION = _raw["reference_version"]
_METHOD_IDS = {"Ts5": 0, "BS5": 1, "CK5": 2, "Me4": 3, "Pr7": 4, "Pr8": 5,
               "Pr9": 6, "CFMR7osc": 7, "CKdisc": 8}

Ts5 = _make("Ts5", _raw["tableaux"]["Ts5"])
BS5 = _make("BS5", _raw["tableaux"]["BS5"])
CK5 = _make("CK5", _raw["tableaux"]["CK5"])
Me4 = _make("Me4", _raw["tableaux"]["Me4"])
Pr7 = _make("Pr7", _raw["tableaux"]["Pr7"])
Pr8 = _make("Pr8", _raw["tableaux"]["Pr8"])
Pr9 = _make("Pr9", _raw["tableaux"]["Pr9"])
CFMR7osc = _make("CFMR7osc", _raw["tableaux"]["CFMR7osc"])
for _n, _c in (("Ts5", Ts5), ("BS5", BS5), ("CK5", CK5), ("Me4", Me4),
               ("Pr7", Pr7), ("Pr8", Pr8), ("Pr9", Pr9),
               ("CFMR7osc", CFMR7osc)):
    _c._xsq_method = _METHOD_IDS[_n]


def _make_ckdisc(d):
    """CKdisc (cash.py:115-416): CK5's stages with assessment and fallback
    weights; its own step rule (max_factor 5, min_factor 1/5), so `sc_params`
    has no effect and the stiffness diagnosis is off (cash.py:238-240)."""
    cls = _make("CKdisc", dict(d, stbrad=None, tanang=None))
    for k in ("B_assess", "E_assess", "C_fallback", "B_fallback", "E_fallback"):
        v = _unhex(d[k])
        v.setflags(write=False)
        setattr(cls, k, v)
    cls.max_factor = d["max_factor"]
    cls.min_factor = d["min_factor"]
    cls.__doc__ = ("Cash-Karp variable order (5, 3, 2) method for non-smooth "
                   "problems; reference extensisq/cash.py:115-416.")
    return cls


CKdisc = _make_ckdisc(_raw["ckdisc"])
CKdisc._xsq_method = _METHOD_IDS["CKdisc"]
del _raw, _fh, _n, _c

BUILTIN = {c.__name__: c for c in (Ts5, BS5, CK5, Me4, Pr7, Pr8, Pr9,
                                   CFMR7osc)}


class RungeKuttaNystrom(RungeKutta):
    """Base of the explicit Runge-Kutta-Nystrom methods for second order
    problems in first order form ``[v, a] = fun(t, [x, v])`` (reference:
    common.py:1207-1320).  Besides ``A, B, C, E`` (positions, with h**2) the
    classes carry ``Ap, Bp, Ep`` (velocities, with h); ``Ap`` is all zero for a
    method for velocity independent problems (``velocity_dependent = False``).
    On the device: final state and counters (no ``t_eval``, no events, no
    stiffness diagnosis)."""
    Ap: np.ndarray = NotImplemented
    Bp: np.ndarray = NotImplemented
    Ep: np.ndarray = NotImplemented
    stbre: float = NotImplemented
    stbim: float = NotImplemented
    velocity_dependent = True


def _make_rkn(name, d):
    scale = float(d.get("embedded_scale", 1.0))    # murua.py:224-227 (scale_embedded=True)
    attrs = dict(
        __doc__=f"{name} Runge-Kutta-Nystrom tableau; reference {d['source']}.",
        n_stages=d["n_stages"], order=d["order"],
        order_secondary=d["order_secondary"], sc_params=d["sc_params"],
        stbre=d["stbre"], stbim=d["stbim"], tanang=d["tanang"], stbrad=None,
        velocity_dependent=bool(d["velocity_dependent"]),
        A=_unhex(d["A"]), Ap=_unhex(d["Ap"]), B=_unhex(d["B"]), Bp=_unhex(d["Bp"]),
        C=_unhex(d["C"]), E=_unhex(d["E"]) * scale, Ep=_unhex(d["Ep"]) * scale,
        P=None,
    )
    for v in attrs.values():
        if isinstance(v, np.ndarray):
            v.setflags(write=False)
    return type(name, (RungeKuttaNystrom,), attrs)


with open(os.path.join(os.path.dirname(_JSON), "tableaux_rkn.json")) as _fh:
    _rkn = json.load(_fh)["tableaux"]
Fi4N = _make_rkn("Fi4N", _rkn["Fi4N"])
Fi5N = _make_rkn("Fi5N", _rkn["Fi5N"])
Mu5Nmb = _make_rkn("Mu5Nmb", _rkn["Mu5Nmb"])
MR6NN = _make_rkn("MR6NN", _rkn["MR6NN"])
for _k, _c in enumerate((Fi4N, Fi5N, Mu5Nmb, MR6NN)):
    _c._xsq_method = 9 + _k                       # XSQ_FI4N .. XSQ_MR6NN
del _rkn, _fh, _k, _c
BUILTIN_RKN = {c.__name__: c for c in (Fi4N, Fi5N, Mu5Nmb, MR6NN)}


class SWAG:
    """Variable-order (1..12) Adams-Bashforth-Moulton PECE of Shampine, Gordon
    and Watts; reference ``extensisq/shampine.py:10-495``.  Like the tableau
    classes this is the method *description* handed to ``solve_ivp_batched``
    (``method=SWAG``, option ``k_max``); the stepping runs in the
    ``swag_persistent`` CUDA kernel."""
    k_max = 12
    _xsq_method = 200
