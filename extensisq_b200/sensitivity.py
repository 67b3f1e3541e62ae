"""Forward sensitivity analysis as an ensemble client of the device solvers.

Mirrors ``extensisq.sens_forward`` (reference sensitivity.py:60-217): the IVP
``y' = f(t, y, p)`` is integrated together with ``S = dy/dp``,
``S' = (df/dy) S + df/dp``, as ONE system of ``ny (1 + np)`` equations per lane,
so a parameter sweep with sensitivities is still one persistent-kernel launch.

What differs from the reference, and why:
* ``fun``, ``jac``, ``dfdp`` are device code (one CUDA source string defining
  the three ``__device__`` functions); the combined right-hand side is
  generated from them and compiled with NVRTC.
* The reference gives parameter j the absolute tolerance ``atol / |p_j|``
  (sensitivity.py:157-162).  Here every lane may have its own ``p`` while the
  tolerances are shared by the ensemble, so the kernel integrates the scaled
  sensitivities ``|p_j| S_j`` under the plain ``atol`` -- the same error test,
  term by term, in exact arithmetic -- and the result is scaled back.
"""
from collections import namedtuple

import numpy as np
import torch

from .batched import DeviceRHS, solve_ivp_batched
from .tableaux import BS5
from . import _lib

SensitivityOutput = namedtuple("ForwardSensitivityOutput", "sensf yf sol")
_cache = {}


def _combined_source(src, names, ny, npar):
    f, j, d = names
    return src + f"""
__device__ void xsq_sens_rhs(double t, const double* Y, const double* p, double* dY) {{
    // Y = [y, |p_0| S[:,0], |p_1| S[:,1], ...]   (columns: order='F' as in
    // sensitivity.py:165-170)
    double J[{ny * ny}], D[{ny * npar}];
    {f}(t, Y, p, dY);
    {j}(t, Y, p, J);
    {d}(t, Y, p, D);
    for (int q = 0; q < {npar}; ++q) {{
        const double fac = p[q] != 0.0 ? fabs(p[q]) : 1.0;
        for (int i = 0; i < {ny}; ++i) {{
            double acc = 0.0;
            for (int k = 0; k < {ny}; ++k) acc = fma(J[i * {ny} + k], Y[{ny} + k + {ny} * q], acc);
            dY[{ny} + i + {ny} * q] = fma(fac, D[i * {npar} + q], acc);
        }}
    }}
}}
"""


def _combined_source_wide(src, names, ny, npar):
    """The same combined system, one component per call (xsq_rhs.cuh WideSystem:
    ny (1 + np) > 16 states, a warp per system).  Every component evaluates the
    user's functions in full; a system this wide has few states per thread."""
    f, j, d = names
    return src + f"""
__device__ double xsq_sens_rhs(int i, double t, const double* Y, const double* p) {{
    if (i < {ny}) {{
        double dy[{ny}];
        {f}(t, Y, p, dy);
        return dy[i];
    }}
    const int q = (i - {ny}) / {ny}, r = (i - {ny}) % {ny};
    double J[{ny * ny}], D[{ny * npar}];
    {j}(t, Y, p, J);
    {d}(t, Y, p, D);
    const double fac = p[q] != 0.0 ? fabs(p[q]) : 1.0;
    double acc = 0.0;
    for (int k = 0; k < {ny}; ++k) acc = fma(J[r * {ny} + k], Y[{ny} + k + {ny} * q], acc);
    return fma(fac, D[r * {npar} + q], acc);
}}
"""


def sens_forward(src, t_span, y0, dy0dp, p, atol=1e-6, rtol=1e-3, method=BS5,
                 t_eval=None, names=("fun", "jac", "dfdp"), device=None, **options):
    """``src``: CUDA source defining
    ``__device__ void fun (double t, const double* y, const double* p, double* dy)``,
    ``__device__ void jac (double t, const double* y, const double* p, double* J)``  (ny x ny, row major),
    ``__device__ void dfdp(double t, const double* y, const double* p, double* D)``  (ny x np, row major).
    ``y0`` [N, ny] (or [ny]), ``p`` [N, np] (or [np]), ``dy0dp`` [N, ny, np] (or
    [ny, np]).  Returns (sensf [N, ny, np], yf [N, ny], sol) like the reference
    (sensitivity.py:214-217); ``sol`` is the BatchedOdeResult of the combined
    system (its sensitivity block is scaled by |p_j|, see the module docstring;
    ``unscale(sol.y, p)`` undoes that for t_eval output)."""
    y0 = np.atleast_2d(np.asarray(y0, dtype=np.float64))
    N, ny = y0.shape
    p = np.asarray(p, dtype=np.float64)
    p = np.broadcast_to(p, (N, p.shape[-1])).copy()
    npar = p.shape[1]
    dy0dp = np.asarray(dy0dp, dtype=np.float64)
    if dy0dp.shape[-2:] != (ny, npar):                 # sensitivity.py:139-140
        raise AssertionError("`dy0dp` should be a array of size (ny, np)")
    dy0dp = np.broadcast_to(dy0dp, (N, ny, npar))
    wide = ny * (1 + npar) > _lib.XSQ_MAX_LANE_STATE
    if ny * (1 + npar) > _lib.XSQ_MAX_WARP_STATE or (wide and npar > 16):
        raise ValueError(f"ny * (np + 1) = {ny * (1 + npar)} exceeds the "
                         f"{_lib.XSQ_MAX_WARP_STATE} states (16 parameters) of a "
                         "warp-per-system kernel")
    if t_eval is not None and float(t_eval[-1]) != float(t_span[1]):
        raise AssertionError("if `t_eval` is used, the last point should be "
                             "t_span[-1]")                 # sensitivity.py:143-145
    if not isinstance(rtol, float):
        raise AssertionError("rtol should be a float")
    atol_y = np.broadcast_to(np.asarray(atol, dtype=np.float64), (ny,)) \
        if np.ndim(atol) == 0 or len(np.atleast_1d(atol)) == ny else None
    if atol_y is None:
        raise AssertionError("`atol` should be a float or a sequence of floats "
                             "of length Ny")
    key = (src, tuple(names), ny, npar)
    if key not in _cache:
        gen = _combined_source_wide if wide else _combined_source
        _cache[key] = DeviceRHS.from_source(gen(src, names, ny, npar),
                                            "xsq_sens_rhs", ny * (1 + npar), npar)
    fac = np.where(p != 0.0, np.abs(p), 1.0)                       # [N, np]
    total_y0 = np.concatenate(
        [y0, (dy0dp * fac[:, None, :]).transpose(0, 2, 1).reshape(N, ny * npar)], axis=1)
    total_atol = np.tile(atol_y, 1 + npar)
    sol = solve_ivp_batched(_cache[key], t_span, total_y0, method, params=p, rtol=rtol,
                            atol=total_atol, t_eval=t_eval, device=device, **options)
    yf = sol.y_final[:, :ny]
    fac_t = torch.as_tensor(fac, device=yf.device)
    sensf = sol.y_final[:, ny:].reshape(N, npar, ny).transpose(1, 2) / fac_t[:, None, :]
    return SensitivityOutput(sensf, yf, sol)


def unscale(y_eval, p, ny):
    """[N, ny (1 + np), n_eval] combined dense output -> (y [N, ny, n_eval],
    S [N, ny, np, n_eval])."""
    N = y_eval.shape[0]
    p = torch.as_tensor(np.broadcast_to(np.asarray(p, dtype=np.float64), (N, np.shape(p)[-1])).copy(),
                        device=y_eval.device)
    fac = torch.where(p != 0, p.abs(), torch.ones_like(p))
    npar = p.shape[1]
    S = y_eval[:, ny:, :].reshape(N, npar, ny, -1).transpose(1, 2) / fac[:, None, :, None]
    return y_eval[:, :ny, :], S
