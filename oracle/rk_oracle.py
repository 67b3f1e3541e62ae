"""TEST INFRASTRUCTURE — CPU oracle (NumPy restatement) of extensisq's explicit
adaptive Runge-Kutta path.  NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

It restates, as one tableau-driven function instead of the reference's
``OdeSolver`` class hierarchy, the algorithm of

* ``extensisq/common.py:30-66``    validate_tol / calculate_scale / norm
* ``extensisq/common.py:123-220``  RungeKutta.__init__ (min-step rule,
                                   controller presets, first step)
* ``extensisq/common.py:222-356``  _step_impl / _reassess_stepsize /
                                   _comp_sol_err / _rk_stage
* ``extensisq/bogacki.py:238-393`` BS5 step with pre-error, its interpolants
* ``extensisq/calvo.py:152-261``   CFMR7osc step with early rejection
* ``extensisq/common.py:519-763``  h_start (Watts / SLATEC dhstrt)
* ``extensisq/common.py:766-821``  Horner / cubic dense output
* ``scipy/integrate/_ivp/ivp.py:659-731`` solve_ivp's loop and t_eval slicing
* ``scipy/integrate/_ivp/base.py:179-210`` OdeSolver.step status handling

The NumPy expressions deliberately use the same operations (``K[:i].T @ a``,
``np.maximum``, ``x @ x``) as the reference so that, under the same
NumPy/OpenBLAS, the oracle is *bit-identical* to the live reference.  That
is how it is pinned: ``tools/gen_golden.py`` runs the unmodified reference in
the build container and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` requires exact equality of step counts and
states (to 1e-13 relative, to tolerate a different BLAS on the GPU host).

Stiffness diagnosis (``common.py:370-516``, ``stiff_a..d`` ``:824-1204``, a
port of RKSuite) only emits warnings and spends extra RHS evaluations; it
never changes t, y or h.  It is restated in ``_diagnose_stiffness`` /
``stiff_probe``; default ``nfev_stiff_detect=5000`` as in the reference.
"""
import json
import os
from math import copysign, sqrt

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLEAUX_JSON = os.path.join(os.path.dirname(_HERE), "extensisq_b200",
                              "data", "tableaux.json")

TOO_SMALL_STEP = "Required step size is less than spacing between numbers."
OVERFLOW = "Overflow or underflow encountered."

MIN_FACTOR = 0.2        # common.py:18
MAX_FACTOR = 4.0        # common.py:19
MAX_FACTOR0 = 10        # common.py:20

SC_PRESETS = {"G": (0.7, -0.4, 0, 0.9),      # common.py:167-169
              "S": (0.6, -0.2, 0, 0.9),
              "standard": (1, 0, 0, 0.9)}


def _unhex(a):
    if isinstance(a[0], list):
        return np.array([[float.fromhex(x) for x in r] for r in a])
    return np.array([float.fromhex(x) for x in a])


class Tableau:
    """Plain data holder (A, B, C, E, P, ...) for one method."""

    def __init__(self, d):
        self.name = d["name"]
        self.n_stages = d["n_stages"]
        self.order = d["order"]
        self.order_secondary = d["order_secondary"]
        self.sc_params = d["sc_params"]
        self.stbrad = d.get("stbrad")
        self.tanang = d.get("tanang")
        for k in ("max_factor", "min_factor", "safety"):      # CKdisc
            if k in d:
                setattr(self, k, float(d[k]))
        for k in ("Ap", "Bp", "Ep"):                          # Runge-Kutta-Nystrom
            if k in d:
                setattr(self, k, _unhex(d[k]))
        if "Bp" in d:
            self.stbre, self.stbim = d.get("stbre"), d.get("stbim")
            self.velocity_dependent = d["velocity_dependent"]
            self.embedded_scale = d.get("embedded_scale", 1.0)
        for k in ("A", "B", "C", "E", "P", "E_pre", "B_scale_pre", "C_extra",
                  "A_extra", "Plow", "Pbest", "B_assess", "E_assess",
                  "C_fallback", "B_fallback", "E_fallback"):
            if k in d:
                setattr(self, k, _unhex(d[k]))
        if hasattr(self, "A_extra"):
            # the reference builds A_extra as ``np.array(rows).T`` (strided
            # rows, bogacki.py:148-160); keep that memory order so BLAS takes
            # the same dgemv path and the oracle stays bit-identical
            self.A_extra = np.asfortranarray(self.A_extra)
        self.variant = {"BS5": "bs5", "CFMR7osc": "cfmr",
                        "CKdisc": "ckdisc"}.get(self.name, "generic")
        if "Bp" in d:
            self.variant = "nystrom"
            if self.embedded_scale != 1.0:         # murua.py:224-227, scale_embedded=True
                self.E = self.E * self.embedded_scale
                self.Ep = self.Ep * self.embedded_scale


def load_tableaux(path=_TABLEAUX_JSON):
    with open(path) as fh:
        raw = json.load(fh)["tableaux"]
    return {k: Tableau(v) for k, v in raw.items()}


def load_tableaux_rkn(path=None):
    """Fi4N, Fi5N, Mu5Nmb, MR6NN (extensisq/fine.py, murua.py, mikkawy.py)."""
    path = path or os.path.join(os.path.dirname(_TABLEAUX_JSON), "tableaux_rkn.json")
    return load_tableaux(path)


def load_ckdisc(path=_TABLEAUX_JSON):
    """CKdisc's coefficient set (cash.py:184-236); not a RungeKutta._step_impl
    method, so it is kept apart from load_tableaux()."""
    with open(path) as fh:
        return Tableau(json.load(fh)["ckdisc"])


# --------------------------------------------------------------------------
# helpers  (common.py:30-66)
# --------------------------------------------------------------------------
def validate_tol(rtol, atol, y):
    atol = np.asarray(atol)
    if atol.ndim > 0 and atol.shape != (y.size,):
        raise ValueError("`atol` has wrong shape.")
    if np.any(atol < 0):
        raise ValueError("`atol` must be positive.")
    if not isinstance(rtol, float):
        raise ValueError("`rtol` must be a float.")
    if rtol < 0:
        raise ValueError("`rtol` must be positive.")
    tiny = np.finfo(y.dtype).tiny
    atol = np.maximum(atol, sqrt(tiny))
    epsneg = np.finfo(y.dtype).epsneg
    rtol = np.minimum(np.maximum(rtol, 10 * epsneg), 0.1)
    return rtol, atol


def calculate_scale(atol, rtol, y, y_new, _mean=False):
    if _mean:
        return atol + rtol * 0.5 * (np.abs(y) + np.abs(y_new))
    return atol + rtol * np.maximum(np.abs(y), np.abs(y_new))


def norm(x):
    return (np.real(x @ x.conjugate()) / x.size) ** 0.5


# --------------------------------------------------------------------------
# starting step  (common.py:519-763; J/T options are not on the RK path)
# --------------------------------------------------------------------------
def h_start(df, a, b, y, yprime, morder, rtol, atol):
    if y.size == 0:
        return np.inf
    neq = y.size
    spy = np.empty_like(y)
    pv = np.empty_like(y)
    etol = atol + rtol * np.abs(y)
    big = sqrt(np.finfo(y.dtype).max)
    small = np.nextafter(np.finfo(y.dtype).epsneg, 1.0)

    dx = b - a
    absdx = abs(dx)
    relper = small ** 0.375
    da = copysign(max(min(relper * abs(a), absdx), 100. * small * abs(a)), dx)
    da = da or relper * dx
    sf = df(a + da, y)
    yp = sf - yprime
    delf = norm(yp)
    dfdxb = big
    if delf < big * abs(da):
        dfdxb = delf / abs(da)
    fbnd = norm(sf)

    dely = relper * norm(y)
    dely = dely or relper
    dely = copysign(dely, dx)
    delf = norm(yprime)
    fbnd = max(fbnd, delf)
    if delf:
        spy[:] = yprime
        yp[:] = yprime
    else:
        spy[:] = 0.0
        yp[:] = 1.0
        delf = norm(yp)

    dfdub = 0.0
    lk = min(neq + 1, 3)
    for k in range(1, lk + 1):
        pv[:] = y + dely / delf * yp
        if k == 2:
            yp[:] = df(a + da, pv)
            pv[:] = yp - sf
        else:
            yp[:] = df(a, pv)
            pv[:] = yp - yprime
        fbnd = max(fbnd, norm(yp))
        delf = norm(pv)
        if delf >= big * abs(dely):
            dfdub = big
            break
        dfdub = max(dfdub, delf / abs(dely))
        if k == lk:
            break
        delf = delf or 1.0
        if k == 2:
            dy = y.copy()
            dy[:] = np.where(dy, dy, dely / relper)
        else:
            dy = pv.copy()
            dy[:] = np.where(dy, dy, delf)
        spy[:] = np.where(spy, spy, yp)
        yp[:] = np.where(spy, np.copysign(dy, spy), dy)
        delf = norm(yp)

    ydpb = dfdxb + dfdub * fbnd
    tolexp = np.log10(etol)
    tolsum = tolexp.sum()
    tolmin = min(tolexp.min(), big)
    tolp = 10.0 ** (0.5 * (tolsum / neq + tolmin) / (morder + 1))
    h = absdx
    if ydpb == 0.0 and fbnd == 0.0:
        if tolp < 1.0:
            h = absdx * tolp
    elif ydpb == 0.0:
        if tolp < fbnd * absdx:
            h = tolp / fbnd
    else:
        srydpb = sqrt(0.5 * ydpb)
        if tolp < srydpb * absdx:
            h = tolp / srydpb
    if dfdub:
        h = min(h, 1.0 / dfdub)
    h = max(h, 100.0 * small * abs(a))
    h = h or small * abs(b)
    return copysign(h, dx)


# --------------------------------------------------------------------------
# dense output  (common.py:766-821)
# --------------------------------------------------------------------------
def horner(t_old, t, y_old, Q, ts):
    """Q is K.T @ P (n x p), not yet scaled by h."""
    h = t - t_old
    Qh = Q * h
    ts = np.asarray(ts, dtype=float)
    x = (ts - t_old) / h
    y = Qh.T[-1, :, np.newaxis] * x
    for q in reversed(Qh.T[:-1]):
        y += q[:, np.newaxis]
        y *= x
    y += y_old[:, np.newaxis]
    return y


def cubic(t_old, t, y_old, y, f_old, f, ts):
    h = t - t_old
    x = (np.asarray(ts, dtype=float) - t_old) / h
    h00 = (1.0 + 2.0 * x) * (1.0 - x) ** 2
    h10 = x * (1.0 - x) ** 2 * h
    h01 = x ** 2 * (3.0 - 2.0 * x)
    h11 = x ** 2 * (x - 1.0) * h
    return (h00 * y_old[:, np.newaxis] + h10 * f_old[:, np.newaxis]
            + h01 * y[:, np.newaxis] + h11 * f[:, np.newaxis])


# --------------------------------------------------------------------------
# stiffness diagnosis  (common.py:824-1204: stiff_a, stiff_b, stiff_c, stiff_d)
# --------------------------------------------------------------------------
def _wdot(a, b, wt):
    return (a / wt) @ (b / wt)


def _jac_times(v, havg, x, y, f, fxy, wt, scale, vdotv):      # stiff_d
    temp1 = scale / sqrt(vdotv)
    z = f(x, y + temp1 * v)
    z = havg / temp1 * (z - fxy)
    return z, (z / wt) @ (z / wt)


def _dominant_real(v1v1, v0v1, v0v0, rold):                   # stiff_b
    r = v0v1 / v0v0
    rho = abs(r)
    det = v0v0 * v1v1 - v0v1 ** 2
    res = abs(det / v0v0)
    rootre = det == 0.0 or (res <= 1e-6 * v1v1 and
                            abs(r - rold) <= 0.001 * rho)
    root1 = [r if rootre else 0.0, 0.0]
    return r, rho, root1, [0.0, 0.0], rootre


def _quadratic_roots(alpha, beta):                            # stiff_c
    r1, r2 = [0.0, 0.0], [0.0, 0.0]
    temp = alpha / 2
    disc = temp ** 2 - beta
    if disc == 0.0:
        r1[0] = r2[0] = -temp
        return r1, r2
    sqdisc = sqrt(abs(disc))
    if disc < 0.0:
        r1[0] = r2[0] = -temp
        r1[1] = sqdisc
        r2[1] = -sqdisc
    else:
        r1[0] = -temp - sqdisc if temp > 0.0 else -temp + sqdisc
        r2[0] = beta / r1[0]
    return r1, r2


def stiff_probe(f, x, y, hnow, havg, xend, maxfcn, wt, fxy, v0, cost):
    """stiff_a: returns (stif, rootre, roots) with stif in {True, False,
    None}; roots = (root1, root2, rho) or None."""
    epsneg = np.finfo(float).epsneg
    rootre = None
    if abs(hnow / havg) > 5 or abs(hnow / havg) < 0.2:
        return False, rootre, None
    xtrfcn = cost * abs((xend - x) / havg)
    if xtrfcn <= maxfcn:
        return False, rootre, None
    ynrm = sqrt((y / wt) @ (y / wt))
    sqrrmc = sqrt(epsneg)
    scale = ynrm * sqrrmc
    if scale == 0.0:
        ynrm = sqrt((v0 / wt) @ (v0 / wt))
        scale = ynrm * sqrrmc
        if scale == 0.0:
            return None, rootre, None
    v0v0 = (v0 / wt) @ (v0 / wt)
    if v0v0 == 0.0:
        v0[:] = 1.0
        v0v0 = (v0 / wt) @ (v0 / wt)
    v0nrm = sqrt(v0v0)
    v0 /= v0nrm
    v0v0 = 1.0
    for ntry in range(8):
        v1, v1v1 = _jac_times(v0, havg, x, y, f, fxy, wt, scale, v0v0)
        if sqrt(v1v1) > 1.0e10 * sqrt(v0v0):
            return None, None, None
        v0v1 = (v0 / wt) @ (v1 / wt)
        if ntry == 0:
            rold = v0v1 / v0v0
            if abs(rold) < epsneg ** (1 / 3):
                return False, None, None
        else:
            rold, rho, root1, root2, rootre = _dominant_real(v1v1, v0v1,
                                                             v0v0, rold)
            if rootre:
                break
        v2, v2v2 = _jac_times(v1, havg, x, y, f, fxy, wt, scale, v1v1)
        v0v2 = (v0 / wt) @ (v2 / wt)
        v1v2 = (v1 / wt) @ (v2 / wt)
        rold, rho, root1, root2, rootre = _dominant_real(v2v2, v1v2, v1v1,
                                                         rold)
        if rootre:
            break
        det1 = v0v0 * v1v1 - v0v1 ** 2
        alpha1 = (-v0v0 * v1v2 + v0v1 * v0v2) / det1
        beta1 = (v0v1 * v1v2 - v1v1 * v0v2) / det1
        v3, v3v3 = _jac_times(v2, havg, x, y, f, fxy, wt, scale, v2v2)
        v1v3 = (v1 / wt) @ (v3 / wt)
        v2v3 = (v2 / wt) @ (v3 / wt)
        rold, rho, root1, root2, rootre = _dominant_real(v3v3, v2v3, v2v2,
                                                         rold)
        if rootre:
            break
        det2 = v1v1 * v2v2 - v1v2 ** 2
        alpha2 = (-v1v1 * v2v3 + v1v2 * v1v3) / det2
        beta2 = (v1v2 * v2v3 - v2v2 * v1v3) / det2
        res2 = abs(v3v3 + v2v2 * alpha2 ** 2 + v1v1 * beta2 ** 2 +
                   2 * v2v3 * alpha2 + 2 * v1v3 * beta2 +
                   2 * v1v2 * alpha2 * beta2)
        if res2 <= 1e-6 * v3v3:
            r1, r2 = _quadratic_roots(alpha1, beta1)
            root1, root2 = _quadratic_roots(alpha2, beta2)
            rho = sqrt(root1[0] ** 2 + root1[1] ** 2)
            D1 = (root1[0] - r1[0]) ** 2 + (root1[1] - r1[1]) ** 2
            D2 = (root1[0] - r2[0]) ** 2 + (root1[1] - r2[1]) ** 2
            if sqrt(min(D1, D2)) <= 0.001 * rho:
                break
        v3nrm = sqrt(v3v3)
        v0 = v3 / v3nrm
        v0v0 = 1.0
    else:
        return None, None, None
    return None, rootre, (root1, root2, rho)


# diagnosis codes recorded per trajectory (the reference only warns/logs)
STIFF_REAL, STIFF_COMPLEX, OSCILLATORY = 1, 2, 4


def _diagnose_stiffness(st):                      # common.py:370-516
    if st.nfev_stiff_detect == 0:
        return
    tab = st.tab
    st.okstp += 1
    h = st.h_previous
    st.havg = 0.9 * st.havg + 0.1 * h
    if st.okstp == 20:
        st.havg = h
        st.jflstp = 0
    if st.okstp % 40 == 39:
        lotsfl = st.jflstp >= 10
        st.jflstp = 0
    else:
        lotsfl = False
    many_steps = st.nfev_stiff_detect // tab.n_stages
    toomch = st.okstp % many_steps == many_steps - 1
    if not (toomch or lotsfl):
        return
    s = tab.n_stages
    avgy = 0.5 * (np.abs(st.y) + np.abs(st.y_old))
    wt = np.maximum(avgy, sqrt(np.finfo(float).tiny))
    nystrom = tab.variant == "nystrom"
    if nystrom:                                   # common.py:1372-1386
        v0 = np.atleast_1d(_rkn_estimate_error(st, st.h_previous))
        stif, rootre, root = stiff_probe(
            st.fun_first_order, st.t, st.y, st.h_previous, st.havg, st.t_bound,
            st.nfev_stiff_detect, wt, np.concatenate((st.y[st.nh:], st.f)), v0, s)
    else:
        v0 = np.atleast_1d(st.h_previous * (st.K[:s + st.FSAL].T @
                                            tab.E[:s + st.FSAL]))
        stif, rootre, root = stiff_probe(
            st.fun, st.t, st.y, st.h_previous, st.havg, st.t_bound,
            st.nfev_stiff_detect, wt, st.f, v0, s)
    st.n_stiff_tests += 1
    if root is not None:
        root1, root2, rho = root
        rootre = root1[1] == 0.0
        if root1[0] > 0.0:
            stif = False
        else:
            rho2 = sqrt(root2[0] ** 2 + root2[1] ** 2)
            if rho2 >= 0.9 * rho and root2[0] > 0.0:
                stif = False
            elif abs(root1[1]) > abs(root1[0]) * tab.tanang:
                stif = None
            elif nystrom:                         # common.py:1412-1413
                stif = (abs(root1[0]) >= 0.85 * tab.stbre or
                        abs(root1[1]) >= 0.9 * tab.stbim)
            else:
                stif = rho >= 0.9 * tab.stbrad
    if stif is None:
        if rootre is not None and not rootre and lotsfl:
            st.stiff_flags |= OSCILLATORY
    elif stif and rootre is not None:
        st.stiff_flags |= STIFF_REAL if rootre else STIFF_COMPLEX


# --------------------------------------------------------------------------
# the solver state + one step
# --------------------------------------------------------------------------
class RKState:
    """Everything RungeKutta.__init__ sets up (common.py:187-220)."""

    def __init__(self, tab, fun, t0, y0, t_bound, max_step=np.inf, rtol=1e-3,
                 atol=1e-6, first_step=None, sc_params=None,
                 interpolant=None, nfev_stiff_detect=5000):
        self.tab = tab
        self.nfev = 0
        self._fun = fun
        self.t = t0
        self.t_old = None
        self.y = np.array(y0, dtype=float)
        self.n = self.y.size
        self.t_bound = t_bound
        self.direction = np.sign(t_bound - t0) if t_bound != t0 else 1
        if max_step <= 0:
            raise ValueError("`max_step` must be positive.")
        self.max_step = max_step
        self.rtol, self.atol = validate_tol(rtol, atol, self.y)
        self.f = self.fun(self.t, self.y)
        s = tab.n_stages
        self.error_exponent = -1 / (min(tab.order_secondary, tab.order) + 1)
        # min-step rule, common.py:123-148
        cdiff = 1.
        for c1 in tab.C:
            for c2 in tab.C:
                diff = abs(c1 - c2)
                if diff:
                    cdiff = min(cdiff, diff)
        cdiff = max(cdiff, 1e-3)
        self.h_min_a = 10 * np.finfo(float).epsneg / cdiff
        self.h_min_b = sqrt(np.finfo(float).tiny)
        self.tiny_err = self.h_min_b
        # controller, common.py:166-185
        scp = sc_params or tab.sc_params
        if isinstance(scp, str) and scp in SC_PRESETS:
            kb1, kb2, a, g = SC_PRESETS[scp]
        elif isinstance(scp, tuple) and len(scp) == 4:
            kb1, kb2, a, g = scp
        else:
            raise ValueError('sc_params should be a tuple of length 4 or one '
                             'of the strings "G", "S", "W" or "standard"')
        self.minbeta1 = kb1 * self.error_exponent
        self.minbeta2 = kb2 * self.error_exponent
        self.minalpha = -a
        self.safety = g
        self.safety_sc = g ** (kb1 + kb2)
        self.standard_sc = True
        self.max_factor = MAX_FACTOR0
        self.min_factor = MIN_FACTOR
        # first step, common.py:207-214 (+ scipy validate_first_step)
        if first_step is None:
            b = self.t + self.direction * min(abs(t_bound - self.t),
                                              self.max_step)
            self.h_abs = abs(h_start(self.fun, self.t, b, self.y, self.f,
                                     tab.order_secondary, self.rtol,
                                     self.atol))
        else:
            if first_step <= 0:
                raise ValueError("`first_step` must be positive.")
            if first_step > np.abs(t_bound - t0):
                raise ValueError("`first_step` exceeds bounds.")
            self.h_abs = first_step
        self.FSAL = 1 if tab.E[s] else 0
        # BS5 keeps extended storage for its interpolants, bogacki.py:217-236
        self.interpolant = interpolant
        nrow = s + 1
        if tab.variant == "bs5":
            self.interpolant = interpolant or "low"
            if self.interpolant not in ("best", "low", "free"):
                raise ValueError(
                    "interpolant should be one of: 'best', 'low', 'free'")
            nrow = {"best": s + 4, "low": s + 2, "free": s + 1}[
                self.interpolant]
        self.K_ext = np.zeros((nrow, self.n))
        self.K = self.K_ext[:s + 1]
        self.h_previous = None
        self.y_old = None
        self.f_old = None
        self.error_norm_old = None
        # stiffness diagnosis state, common.py:150-164
        if not (isinstance(nfev_stiff_detect, int) and nfev_stiff_detect >= 0):
            raise ValueError(
                "`nfev_stiff_detect` must be a non-negative integer.")
        self.nfev_stiff_detect = nfev_stiff_detect
        if tab.stbrad is None or tab.tanang is None:   # common.py:155-160
            self.nfev_stiff_detect = 0
        self.jflstp = 0
        self.okstp = 0
        self.havg = 0.0
        self.n_stiff_tests = 0
        self.stiff_flags = 0
        self.n_rejected = 0            # the reference's global NFS
        self.n_accepted = 0
        self.status = "running"

    def fun(self, t, y):
        self.nfev += 1
        return np.asarray(self._fun(t, y), dtype=float)


class RKNState(RKState):
    """RungeKuttaNystrom.__init__ (common.py:1240-1277) on top of RungeKutta's:
    `fun` is the first order form [v, a] = fun(t, [x, v]); K holds accelerations."""

    def __init__(self, tab, fun, t0, y0, t_bound, **kw):
        stiff = kw.pop("nfev_stiff_detect", 5000)
        super().__init__(tab, fun, t0, y0, t_bound, nfev_stiff_detect=0, **kw)
        if not (isinstance(stiff, int) and stiff >= 0):
            raise ValueError("`nfev_stiff_detect` must be a non-negative integer.")
        self.nfev_stiff_detect = stiff            # common.py:1225-1238
        if tab.stbre is None or tab.stbim is None or tab.tanang is None:
            self.nfev_stiff_detect = 0
        n = self.nh = self.y.size // 2
        msg = ('This method is for second order problems'
               ' and `fun` should have signature: [v, a] = fun(t, [x, v]).')
        if (self.y.size % 2) or not np.all(self.y[n:] == self.f[:n]):
            raise AssertionError(msg)
        elif np.all(self.y[n:] == self.y[:n]):
            y_test = self.y.copy()
            y_test[n:] *= 1 + 1e-8
            y_test[n:] += 1e-8
            if not np.all(np.asarray(self._fun(t0, y_test))[:n] == y_test[n:]):
                raise AssertionError(msg)
        if not tab.velocity_dependent:
            y_test = self.y.copy()
            y_test[n:] *= 1 + 1e-8
            y_test[n:] += 1e-8
            if not np.all(np.asarray(self._fun(t0, y_test))[n:] == self.f[n:]):
                raise AssertionError("This method is for velocity independent ODEs, "
                                     "but `fun` seems velocity dependent.")
        self.Ap = tab.Ap                          # zeros for a velocity independent method
        s = tab.n_stages
        if tab.Ep[s] != 0.:
            self.FSAL = 1
        self.K_ext = np.empty((s + 1, n))
        self.K = self.K_ext
        self.f = self.f[n:]
        self._first = self._fun

    def fun_first_order(self, t, y):
        # the reference keeps the caller's raw function here (common.py:1271), so
        # the evaluations of the stiffness probe are NOT counted in nfev
        return np.asarray(self._first(t, y), dtype=float)

    def fun(self, t, y):
        self.nfev += 1
        out = np.asarray(self._fun(t, y), dtype=float)
        return out[self.nh:] if hasattr(self, "_first") else out


def _reassess_stepsize(st):                       # common.py:310-331
    h_abs = st.h_abs
    min_step = max(st.h_min_a * (abs(st.t) + h_abs), st.h_min_b)
    if h_abs < min_step or h_abs > st.max_step:
        h_abs = min(st.max_step, max(min_step, h_abs))
        st.standard_sc = True
    d = abs(st.t_bound - st.t)
    if d < 2 * h_abs:
        if d > h_abs:
            h_abs = max(0.5 * d, min_step)
            st.standard_sc = True
        else:
            h_abs = d
    return h_abs, min_step


def _rk_stage(st, h, i):                          # common.py:353-356
    if st.tab.variant == "nystrom":               # common.py:1279-1285
        n = st.nh
        dt = st.tab.C[i] * h
        du = (st.K[:i, :].T @ st.tab.A[i, :i]) * h**2 + dt * st.y[n:]
        dv = (st.K[:i, :].T @ st.Ap[i, :i]) * h
        st.K[i] = st.fun(st.t + dt, st.y + np.concatenate((du, dv)))
        return
    dy = h * (st.K[:i, :].T @ st.tab.A[i, :i])
    st.K[i] = st.fun(st.t + st.tab.C[i] * h, st.y + dy)


def _rkn_estimate_error(st, h):                   # common.py:1303-1309
    s = st.tab.n_stages
    eu = (st.K[:s + st.FSAL, :].T @ st.tab.E[:s + st.FSAL]) * h**2
    ev = (st.K[:s + st.FSAL, :].T @ st.tab.Ep[:s + st.FSAL]) * h
    return np.concatenate((eu, ev))


def _comp_sol_err(st, y, h):                      # common.py:333-351
    s = st.tab.n_stages
    if st.tab.variant == "nystrom":               # common.py:1287-1301
        n = st.nh
        du = (st.K[:s, :].T @ st.tab.B) * h**2 + h * st.y[n:]
        dv = (st.K[:s, :].T @ st.tab.Bp) * h
        y_new = y + np.concatenate((du, dv))
        scale = calculate_scale(st.atol, st.rtol, y, y_new)
        if st.FSAL:
            st.K[s, :] = st.fun(st.t + h, y_new)
        return y_new, norm(_rkn_estimate_error(st, h) / scale)
    y_new = y + h * (st.K[:s].T @ st.tab.B)
    scale = calculate_scale(st.atol, st.rtol, y, y_new)
    if st.FSAL:
        st.K[s, :] = st.fun(st.t + h, y_new)
    err = h * (st.K[:s + st.FSAL].T @ st.tab.E[:s + st.FSAL])
    return y_new, norm(err / scale)


def _pre_error(st, y, h):
    tab = st.tab
    if tab.variant == "bs5":                      # bogacki.py:340-346
        y_pre = y + h * (st.K[:6].T @ tab.B_scale_pre)
        scale = calculate_scale(st.atol, st.rtol, y, y_pre)
        err = h * (st.K[:6, :].T @ tab.E_pre)
    else:                                         # calvo.py:255-261
        y_pre = y + h * (st.K[:8].T @ tab.A[8, :8])
        scale = calculate_scale(st.atol, st.rtol, y, y_pre)
        err = h * (st.K[:8, :].T @ tab.E[:8])
    return norm(err / scale)


def rk_step(st, forced_h=None):
    """One call of _step_impl.  Returns (success, message).

    ``forced_h``: take exactly this |h| and accept whatever the error is
    (the "forced fixed step sequence" mode of BASELINE.json's north_star)."""
    tab = st.tab
    s = tab.n_stages
    t, y = st.t, st.y
    if forced_h is not None:
        h = forced_h * st.direction
        st.K[0] = st.f
        for i in range(1, s):
            _rk_stage(st, h, i)
        y_new, error_norm = _comp_sol_err(st, y, h)
        h_abs = forced_h
    else:
        h_abs, min_step = _reassess_stepsize(st)
        step_accepted = False
        step_rejected = False
        early = tab.variant in ("bs5", "cfmr")
        while not step_accepted:
            if h_abs < min_step:
                return False, TOO_SMALL_STEP
            h = h_abs * st.direction
            st.K[0] = st.f
            n_first = s - 1 if early else s
            for i in range(1, n_first):
                _rk_stage(st, h, i)
            if early:
                # bogacki.py:262-275, calvo.py:174-187
                error_norm_pre = _pre_error(st, y, h)
                if error_norm_pre > 1:
                    step_rejected = True
                    h_abs *= max(st.min_factor, st.safety *
                                 error_norm_pre ** st.error_exponent)
                    st.n_rejected += 1
                    if st.nfev_stiff_detect:          # bogacki.py:272-273
                        st.jflstp += 1
                    continue
                _rk_stage(st, h, s - 1)
            y_new, error_norm = _comp_sol_err(st, y, h)

            if error_norm < 1:                    # common.py:249-276
                step_accepted = True
                if error_norm < st.tiny_err:
                    factor = st.max_factor
                    st.standard_sc = True
                elif st.standard_sc:
                    factor = st.safety * error_norm ** st.error_exponent
                    st.standard_sc = False
                else:
                    h_ratio = h / st.h_previous
                    factor = st.safety_sc * (
                        error_norm ** st.minbeta1 *
                        st.error_norm_old ** st.minbeta2 *
                        h_ratio ** st.minalpha)
                    factor = min(st.max_factor, max(st.min_factor, factor))
                if step_rejected:
                    factor = min(1, factor)
                h_abs *= factor
                if factor < MAX_FACTOR:
                    st.max_factor = MAX_FACTOR
            else:
                bad = np.isnan(error_norm) or np.isinf(error_norm)
                if tab.variant == "bs5" and bad:  # bogacki.py:314-315
                    return False, OVERFLOW
                step_rejected = True
                h_abs *= max(st.min_factor,
                             st.safety * error_norm ** st.error_exponent)
                st.n_rejected += 1
                st.jflstp += 1                    # common.py:284
                if bad:                           # common.py:286-287
                    return False, OVERFLOW

    if not st.FSAL:                               # common.py:289-291
        st.K[s] = st.fun(t + h, y_new)
    st.h_previous = h
    st.y_old = y
    st.h_abs = h_abs
    st.f_old = st.f
    st.f = st.K[s].copy()
    st.error_norm_old = error_norm
    st.t_old = t
    st.t = t + h
    st.y = y_new
    st.n_accepted += 1
    if forced_h is None:
        _diagnose_stiffness(st)                   # common.py:306
    return True, None


def _ck_sol_err_tol(st, h, B, E, i=6):             # cash.py:394-404
    sol = h * (st.K[:i, :].T @ B[:i]) + st.y
    err = h * (st.K[:i, :].T @ E[:i])
    tol = calculate_scale(st.atol, st.rtol, st.y, sol)
    return sol, err, tol


def ckdisc_step(st):
    """CKdisc._step_impl, cash.py:245-388: the Cash-Karp variable order
    (5, 3, 2) step.  Convergence is assessed after stages 1 and 3; when the
    fifth order result is not accepted, embedded third / second order
    solutions over a FRACTION of the step (c = 3/5, 1/5) are tried before the
    step is repeated."""
    tab = st.tab
    t, y = st.t, st.y
    if not hasattr(st, "twiddle"):                 # cash.py:241-243
        st.twiddle = [1.5, 1.1]
        st.quit = [100., 100.]
    twiddle, quit = st.twiddle, st.quit
    h_abs, min_step = _reassess_stepsize(st)
    order_accepted = 0
    step_rejected = False
    while not order_accepted:
        if h_abs < min_step:
            return False, TOO_SMALL_STEP
        h = h_abs * st.direction
        st.K[0] = st.f
        _rk_stage(st, h, 1)
        _, err_a, tol = _ck_sol_err_tol(st, h, tab.B_assess[0],
                                        tab.E_assess[0], 2)
        E1 = norm(err_a / tol) ** (1 / 2)
        esttol = E1 / quit[0]
        if E1 < twiddle[0] * quit[0]:
            _rk_stage(st, h, 2)
            _rk_stage(st, h, 3)
            _, err_a, tol = _ck_sol_err_tol(st, h, tab.B_assess[1],
                                            tab.E_assess[1], 4)
            E2 = norm(err_a / tol) ** (1 / 3)
            esttol = E2 / quit[1]
            if E2 < twiddle[1] * quit[1]:
                _rk_stage(st, h, 4)
                _rk_stage(st, h, 5)
                y_new, err, tol = _ck_sol_err_tol(st, h, tab.B, tab.E)
                E4 = norm(err / tol) ** (1 / 5)
                E4 = E4 or 1e-160
                esttol = E4
                if E4 < 1:
                    order_accepted = 4
                    factor = min(tab.max_factor, tab.safety / E4)
                    if step_rejected:
                        factor = min(1.0, factor)
                    h_abs *= factor
                    q = [E1 / E4, E2 / E4]
                    for j in (0, 1):
                        if q[j] > quit[j]:
                            q[j] = min(q[j], 10 * quit[j])
                        else:
                            q[j] = max(q[j], 2 / 3 * quit[j])
                        quit[j] = max(1., min(10000., q[j]))
                    break
                if np.isnan(E4) or np.isinf(E4):
                    return False, OVERFLOW
                e = [E1, E2]
                for i in (0, 1):
                    EQ = e[i] / quit[i]
                    if EQ < twiddle[i]:
                        twiddle[i] = max(1.1, EQ)
                if E2 < 1:
                    y_new, err, tol = _ck_sol_err_tol(
                        st, h, tab.B_fallback[1], tab.E_fallback[1], 4)
                    if norm(err / tol) < 1:
                        order_accepted = 2
                        h_abs *= tab.C_fallback[1]
                        h = h_abs * st.direction
                        break
            if E1 < 1:
                y_new, err, tol = _ck_sol_err_tol(
                    st, h, tab.B_fallback[0], tab.E_fallback[0], 2)
                if norm(err / tol) < 1:
                    order_accepted = 1
                    h_abs *= tab.C_fallback[0]
                    h = h_abs * st.direction
                    break
                else:
                    step_rejected = True
                    h_abs *= tab.C_fallback[0]
                    st.n_rejected += 1
                    continue
        step_rejected = True
        h_abs *= max(tab.min_factor, tab.safety / esttol)
        st.n_rejected += 1
        continue
    t_new = t + h
    f_new = st.fun(t_new, y_new)
    st.K[-1, :] = f_new
    st.order_accepted = order_accepted
    st.h_previous = h
    st.y_old = y
    st.h_abs = h_abs
    st.f = f_new
    st.t_old = t
    st.t = t_new
    st.y = y_new
    st.n_accepted += 1
    return True, None


def make_dense(st):
    """``solver.dense_output()`` for the last accepted step: a callable sol(ts)
    (ts scalar or array).  common.py:358-368; bogacki.py:348-393 (the extra
    stages of BS5's 'low'/'best' interpolants are evaluated HERE, once per
    call, as in the reference); cash.py:406-416."""
    tab = st.tab
    t_old, t, y_old, y = st.t_old, st.t, st.y_old, st.y

    def wrap(fn):
        def sol(ts):
            ts = np.asarray(ts, dtype=float)
            out = fn(np.atleast_1d(ts))
            return out if ts.shape else out[:, 0]
        return sol
    if t == t_old:                 # scipy base.py:224-226 ConstantDenseOutput
        return wrap(lambda ts: np.repeat(y[:, None], np.size(ts), axis=1))
    if tab.variant == "ckdisc" and st.order_accepted != 4:
        K0, K1 = st.K[0].copy(), st.K[-1].copy()
        return wrap(lambda ts: cubic(t_old, t, y_old, y, K0, K1, ts))
    if tab.variant != "bs5":
        Q = st.K.T @ tab.P
        return wrap(lambda ts: horner(t_old, t, y_old, Q, ts))
    h = st.h_previous
    K = st.K_ext
    s = tab.n_stages
    if st.interpolant == "free":
        Q = K.T @ tab.P
        return wrap(lambda ts: horner(t_old, t, y_old, Q, ts))
    if st.interpolant == "low":
        r = s + 1
        dy = K[:r, :].T @ tab.A_extra[0, :r] * h
        K[r] = st.fun(t_old + tab.C_extra[0] * h, y_old + dy)
        Q = K.T @ tab.Plow
        return wrap(lambda ts: horner(t_old, t, y_old, Q, ts))
    for r, (a, c) in enumerate(zip(tab.A_extra, tab.C_extra), start=s + 1):
        dy = K[:r, :].T @ a[:r] * h
        K[r] = st.fun(t_old + c * h, y_old + dy)
    Pb = tab.Pbest
    Q = np.empty((K.shape[1], Pb.shape[1]))
    Q[:, 0] = st.K[7]
    KP = K * Pb[:, 1, np.newaxis]
    Q[:, 1] = (KP[4] + ((KP[5] + KP[7]) + KP[0]) + ((KP[2] + KP[8]) +
               KP[9]) + ((KP[3] + KP[10]) + KP[6]))
    KP = K * Pb[:, 2, np.newaxis]
    Q[:, 2] = (KP[4] + KP[5] + ((KP[2] + KP[8]) + (KP[9] + KP[7]) +
               KP[0]) + ((KP[3] + KP[10]) + KP[6]))
    KP = K * Pb[:, 3, np.newaxis]
    Q[:, 3] = (((KP[3] + KP[7]) + (KP[6] + KP[5]) + KP[4]) + ((KP[9] +
               KP[8]) + (KP[2] + KP[10]) + KP[0]))
    KP = K * Pb[:, 4, np.newaxis]
    Q[:, 4] = ((KP[9] + KP[8]) + ((KP[6] + KP[5]) + KP[4]) + ((KP[3] +
               KP[7]) + (KP[2] + KP[10]) + KP[0]))
    KP = K * Pb[:, 5, np.newaxis]
    Q[:, 5] = (KP[4] + ((KP[9] + KP[7]) + (KP[6] + KP[5])) + ((KP[3] +
               KP[8]) + (KP[2] + KP[10]) + KP[0]))
    # anchored at the END of the step (bogacki.py:389-393)
    return wrap(lambda ts: horner(t, t + h, y, Q, ts))


def dense_eval(st, ts):
    """dense_output()(ts) over the last accepted step."""
    return make_dense(st)(ts)


# --------------------------------------------------------------------------
# events: scipy's solve_ivp machinery (third party, scipy/integrate/_ivp/ivp.py
# find_active_events / handle_events / solve_event_equation, and
# scipy/optimize/Zeros/brentq.c), restated; pinned in tests against
# scipy.optimize.brentq and against solve_ivp runs of the reference.
# --------------------------------------------------------------------------
_EPS = float(np.finfo(float).eps)


def brentq(f, xa, xb, xtol=4 * _EPS, rtol=4 * _EPS, maxiter=100):
    """Brent's method as coded in scipy's brentq.c.  Returns (root, calls)."""
    def neg(v):
        return copysign(1.0, v) < 0
    xpre, xcur = xa, xb
    xblk = fblk = spre = scur = 0.0
    fpre, fcur = f(xpre), f(xcur)
    calls = 2
    if fpre == 0:
        return xpre, calls
    if fcur == 0:
        return xcur, calls
    if neg(fpre) == neg(fcur):
        raise ValueError("f(a) and f(b) must have different signs")
    for _ in range(maxiter):
        if fpre != 0 and fcur != 0 and neg(fpre) != neg(fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, calls
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)
            else:
                dpre = (fpre - fcur) / (xpre - xcur)
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = (-fcur * (fblk * dblk - fpre * dpre) /
                        (dblk * dpre * (fblk - fpre)))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = f(xcur)
        calls += 1
    return xcur, calls


def find_active_events(g, g_new, direction):       # ivp.py
    g, g_new = np.asarray(g), np.asarray(g_new)
    up = (g <= 0) & (g_new >= 0)
    down = (g >= 0) & (g_new <= 0)
    either = up | down
    mask = (up & (direction > 0) | down & (direction < 0) |
            either & (direction == 0))
    return np.nonzero(mask)[0]


def handle_events(sol, events, active_events, event_count, max_events,
                  t_old, t):                          # ivp.py
    roots = np.asarray([brentq(lambda tt, ev=events[i]: ev(tt, sol(tt)),
                               t_old, t)[0] for i in active_events])
    if np.any(event_count[active_events] >= max_events[active_events]):
        order = np.argsort(roots) if t > t_old else np.argsort(-roots)
        active_events = active_events[order]
        roots = roots[order]
        k = np.nonzero(event_count[active_events] >=
                       max_events[active_events])[0][0]
        return active_events[:k + 1], roots[:k + 1], True
    return active_events, roots, False


def rk_solve(tab, fun, t_span, y0, rtol=1e-3, atol=1e-6, max_step=np.inf,
             first_step=None, t_eval=None, sc_params=None, interpolant=None,
             forced_h=None, record=False, max_steps=None,
             nfev_stiff_detect=5000, events=None):
    """solve_ivp(fun, t_span, y0, method=<tab>, ...) restated.

    Returns a dict with t, y (n x n_t), n_accepted, n_rejected, nfev, status,
    message, t_final, y_final and (``record=True``) the accepted |h| sequence.
    With ``forced_h`` (sequence of |h|) exactly len(forced_h) steps are taken
    and t_span[1] only gives the direction."""
    t0, tf = map(float, t_span)
    step_fn = ckdisc_step if tab.variant == "ckdisc" else rk_step
    if tab.variant == "ckdisc":
        if forced_h is not None or sc_params is not None:
            raise ValueError("CKdisc has no forced steps / sc_params")
        nfev_stiff_detect = 0                      # cash.py:238-240
    RKState_ = RKNState if tab.variant == "nystrom" else RKState
    if forced_h is not None:
        st = RKState_(tab, fun, t0, y0, copysign(np.inf, tf - t0),
                     rtol=rtol, atol=atol, first_step=forced_h[0],
                     sc_params=sc_params, interpolant=interpolant,
                     nfev_stiff_detect=0)
    else:
        st = RKState_(tab, fun, t0, y0, tf, max_step=max_step, rtol=rtol,
                     atol=atol, first_step=first_step, sc_params=sc_params,
                     interpolant=interpolant,
                     nfev_stiff_detect=nfev_stiff_detect)
    if t_eval is not None:
        t_eval = np.asarray(t_eval, dtype=float)
        if st.direction > 0:
            t_eval_i = 0
        else:
            t_eval = t_eval[::-1]
            t_eval_i = t_eval.shape[0]
        ts, ys = [], []
    else:
        ts, ys = [t0], [st.y]
    hs = []
    status = None
    message = None
    k = 0
    # events: list of (fn(t, y), terminal, direction); terminal False/0 = never,
    # True/1 = first occurrence, n = n-th occurrence (ivp.py prepare_events)
    t_events = y_events = None
    if events is not None:
        ev_fns = [e[0] for e in events]
        max_events = np.array([np.inf if not e[1] else float(int(e[1]))
                               for e in events])
        ev_dir = np.array([float(e[2]) for e in events])
        event_count = np.zeros(len(events))
        g = [ev(t0, st.y) for ev in ev_fns]
        t_events = [[] for _ in events]
        y_events = [[] for _ in events]
    while status is None:
        # OdeSolver.step, scipy base.py:179-210
        if st.n == 0 or st.t == st.t_bound:
            st.t_old = st.t
            st.t = st.t_bound
            st.status = "finished"
        else:
            if forced_h is not None:
                ok, message = rk_step(st, forced_h=forced_h[k])
            else:
                ok, message = step_fn(st)
            if not ok:
                st.status = "failed"
            elif st.direction * (st.t - st.t_bound) >= 0:
                st.status = "finished"
        k += 1
        if st.status == "finished":
            status = 0
        elif st.status == "failed":
            status = -1
            break
        if record:
            hs.append(abs(st.h_previous))
        t = st.t
        y_out = st.y
        sol = None
        if events is not None:                    # ivp.py, event block
            g_new = [ev(t, st.y) for ev in ev_fns]
            active = find_active_events(g, g_new, ev_dir)
            if active.size > 0:
                sol = make_dense(st)
                event_count[active] += 1
                idx, roots, terminate = handle_events(
                    sol, ev_fns, active, event_count, max_events, st.t_old, t)
                for e, te in zip(idx, roots):
                    t_events[e].append(te)
                    y_events[e].append(sol(te))
                if terminate:
                    status = 1
                    t = roots[-1]
                    y_out = sol(t)
            g = g_new
        if t_eval is None:
            ts.append(t)
            ys.append(y_out)
        else:                                     # ivp.py:711-728
            if st.direction > 0:
                i_new = np.searchsorted(t_eval, t, side="right")
                step_pts = t_eval[t_eval_i:i_new]
            else:
                i_new = np.searchsorted(t_eval, t, side="left")
                step_pts = t_eval[i_new:t_eval_i][::-1]
            if step_pts.size > 0:
                if sol is None:
                    sol = make_dense(st)
                ts.append(step_pts)
                ys.append(sol(step_pts))
                t_eval_i = i_new
        if forced_h is not None and k >= len(forced_h):
            status = 0
        if max_steps is not None and k >= max_steps and status is None:
            status = 0
    if t_eval is None:
        ts = np.array(ts)
        ys = np.vstack(ys).T
    elif ts:
        ts = np.hstack(ts)
        ys = np.hstack(ys)
    else:
        ts = np.zeros(0)
        ys = np.zeros((st.n, 0))
    out = dict(t=ts, y=ys, n_accepted=st.n_accepted, n_rejected=st.n_rejected,
               nfev=st.nfev, status=status, message=message, t_final=st.t,
               y_final=st.y.copy(), h_next=st.h_abs,
               n_stiff_tests=st.n_stiff_tests, stiff_flags=st.stiff_flags)
    if events is not None:
        out["t_events"] = [np.asarray(te) for te in t_events]
        out["y_events"] = [np.asarray(ye) for ye in y_events]
        if status == 1:                  # the solver itself stands at st.t
            out["t_final"], out["y_final"] = t, np.array(y_out)
    if record:
        out["h"] = np.array(hs)
    return out


# --------------------------------------------------------------------------
# built-in right-hand sides (SURVEY.md §8d synthetic inputs)
# --------------------------------------------------------------------------
def lorenz63(sigma, rho, beta):
    def f(t, y):
        return np.array([sigma * (y[1] - y[0]),
                         y[0] * (rho - y[2]) - y[1],
                         y[0] * y[1] - beta * y[2]])
    return f


def vanderpol(mu):
    def f(t, y):
        return np.array([y[1], mu * (1.0 - y[0] * y[0]) * y[1] - y[0]])
    return f


def arenstorf(mu):
    def f(t, y):
        x, yy, vx, vy = y
        mup = 1.0 - mu
        d1 = ((x + mu) * (x + mu) + yy * yy)
        d1 = d1 * sqrt(d1)
        d2 = ((x - mup) * (x - mup) + yy * yy)
        d2 = d2 * sqrt(d2)
        return np.array([
            vx, vy,
            x + 2.0 * vy - mup * (x + mu) / d1 - mu * (x - mup) / d2,
            yy - 2.0 * vx - mup * yy / d1 - mu * yy / d2])
    return f


def nbody(masses, eps2):
    """3-D softened gravity, G=1; y = [pos(3*nb), vel(3*nb)] body-major."""
    m = np.asarray(masses, dtype=float)
    nb = m.size

    def f(t, y):
        pos = y[:3 * nb].reshape(nb, 3)
        acc = np.zeros((nb, 3))
        for i in range(nb):
            a = np.zeros(3)
            for j in range(nb):
                if j == i:
                    continue
                d = pos[j] - pos[i]
                r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + eps2
                inv = 1.0 / (r2 * sqrt(r2))
                a += m[j] * inv * d
            acc[i] = a
        return np.concatenate([y[3 * nb:], acc.ravel()])
    return f
