"""TEST INFRASTRUCTURE -- second order problems in first order form,
[v, a] = fun(t, [x, v]), for the Runge-Kutta-Nystrom methods (reference
common.py:1207-1320).  Only tests/ and tools/ import this."""
from math import sqrt

import numpy as np

from . import rk_oracle as RO


def oscillator(w):                       # reference tests/test_rkn.py:14-15 (w = 1)
    def f(t, y):
        return np.array([y[1], -(w * w) * y[0]])
    return f


def kepler():                            # planar two-body problem, velocity independent
    def f(t, y):
        r2 = y[0] * y[0] + y[1] * y[1]
        r3 = r2 * sqrt(r2)
        return np.array([y[2], y[3], -y[0] / r3, -y[1] / r3])
    return f


def damped(k, c):                        # x'' = -k x - c x': velocity dependent, stiff for large c
    def f(t, y):
        return np.array([y[1], -k * y[0] - c * y[1]])
    return f


def nbody32_setup(seed=7):
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.5, 1.5, 32) / 32.0
    pos = rng.uniform(-1.0, 1.0, (32, 3))
    vel = rng.uniform(-0.3, 0.3, (32, 3))
    return m, 0.01, np.concatenate([pos.ravel(), vel.ravel()])


def make_fun(problem, params):
    if problem == "oscillator":
        return oscillator(params[0])
    if problem == "kepler":
        return kepler()
    if problem == "damped":
        return damped(params[0], params[1])
    if problem == "vanderpol":
        return RO.vanderpol(params[0])
    if problem == "arenstorf":
        return RO.arenstorf(params[0])
    if problem == "nbody32":
        m, eps2, _ = nbody32_setup()
        return RO.nbody(m, eps2)
    raise KeyError(problem)


PROBLEMS = {
    "oscillator": dict(y0=lambda p: np.array([0.0, 1.0])),
    "kepler": dict(y0=lambda p: np.array([1.0 - p[0], 0.0, 0.0, sqrt((1.0 + p[0]) / (1.0 - p[0]))])),
    "damped": dict(y0=lambda p: np.array([1.0, 0.0])),
    "vanderpol": dict(y0=lambda p: np.array([2.0, 0.0])),
    "arenstorf": dict(y0=lambda p: np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224])),
    "nbody32": dict(y0=lambda p: nbody32_setup()[2]),
}
