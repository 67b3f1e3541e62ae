"""GPU parity tests of the SSV2stab path (xsq_rkc_solve through ctypes):
against the reference's golden vectors, the C oracle at larger grids, and --
when two devices are visible -- the 2-rank slab decomposition against the
1-rank run."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import extensisq_b200 as xb
from oracle import c_oracle as CO
from oracle.problems import heat2d_reaction
from test_rkc_oracle_golden import CASES, Z

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def solve(c_or_nx, t_span, use_rho=True, t_eval=None, **opts):
    nx = c_or_nx
    _, y0, rho = heat2d_reaction(nx)
    r = xb.solve_pde_rkc("heat2d_reaction", t_span, y0.reshape(nx, nx),
                         rho_jac=float(rho) if use_rho else None,
                         t_eval=t_eval, max_steps=100000, **opts)
    torch.cuda.synchronize()
    return r


@pytest.mark.parametrize("c", [c for c in CASES if not c.get("notebook")],
                         ids=lambda c: c["id"])
def test_rkc_vs_reference_golden(c):
    te = np.linspace(*c["t_eval"]) if c.get("t_eval") else None
    r = solve(c["nx"], c["t_span"], c["use_rho"], te, **c["options"])
    assert r.status == 0
    assert (r.nfev, r.n_rejected, r.nfesig, r.maxm) == \
        (c["nfev"], c["nfs"], c["nfesig"], c["maxm"])
    yg = Z[c["id"] + "/y"]
    if te is None:
        assert r.n_accepted == c["n_t"] - 1
        got = r.y_final.cpu().numpy().reshape(-1)
    else:
        got = r.y.cpu().numpy().reshape(te.size, -1).T
    assert np.abs(got - yg).max() <= 1e-11
    assert xb.maxm == c["maxm"] and xb.nfesig == c["nfesig"]


@pytest.mark.parametrize("nx,tol", [(256, 1e-4), (512, 1e-4), (256, 1e-6)])
def test_rkc_vs_c_oracle_larger_grids(nx, tol):
    """BASELINE.md section 2, C5 proxy: N = 256 / 512 -> 10 accepted steps,
    nfev 607 / 1205, s-max 76 / 151 at tol 1e-4."""
    _, y0, rho = heat2d_reaction(nx)
    r = solve(nx, (0.0, 0.05), rtol=tol, atol=tol)
    ref = CO.rkc_solve(y0, (0.0, 0.05), rtol=tol, atol=tol, rho=rho)
    assert (r.n_accepted, r.n_rejected, r.nfev, r.maxm) == \
        (ref["n_accepted"], ref["n_rejected"], ref["nfev"], ref["maxm"])
    if tol == 1e-4:
        assert (r.n_accepted, r.nfev, r.maxm) == \
            {256: (10, 607, 76), 512: (10, 1205, 151)}[nx]
    got = r.y_final.cpu().numpy().reshape(-1)
    assert np.abs(got - ref["y_final"]).max() <= 1e-11


def test_rkc_power_iteration_and_rho_callback():
    nx = 64
    _, y0, rho = heat2d_reaction(nx)
    r = solve(nx, (0.0, 0.05), use_rho=False, rtol=1e-4, atol=1e-4)
    ref = CO.rkc_solve(y0, (0.0, 0.05), rtol=1e-4, atol=1e-4, rho=None)
    assert (r.nfev, r.nfesig, r.maxm, r.n_accepted) == \
        (ref["nfev"], ref["nfesig"], ref["maxm"], ref["n_accepted"])
    assert r.nfesig > 0
    got = r.y_final.cpu().numpy().reshape(-1)
    assert np.abs(got - ref["y_final"]).max() <= 1e-10
    calls = []
    r2 = xb.solve_pde_rkc("heat2d_reaction", (0.0, 0.05), y0.reshape(nx, nx),
                          rho_jac=lambda t: calls.append(t) or float(rho),
                          rtol=1e-4, atol=1e-4)
    r3 = solve(nx, (0.0, 0.05), rtol=1e-4, atol=1e-4)
    assert len(calls) >= r2.n_accepted
    assert torch.equal(r2.y_final, r3.y_final)


def test_rkc_argument_errors_and_edge_cases():
    nx = 32
    _, y0, rho = heat2d_reaction(nx)
    u0 = y0.reshape(nx, nx)
    with pytest.raises(TypeError, match="const_jac"):
        xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), u0, const_jac=1)
    with pytest.raises(TypeError, match="rho_jac"):
        xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), u0, rho_jac=3)
    with pytest.raises(ValueError, match="positive float"):
        xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), u0, rho_jac=-1.0)
    with pytest.raises(ValueError, match="first_step"):
        xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), u0, first_step=1.0)
    with pytest.raises(ValueError, match="multiple of 4"):
        xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), np.zeros((5, 5)))
    r = xb.solve_pde_rkc("heat2d_reaction", (0.3, 0.3), u0, rho_jac=float(rho),
                         t_eval=[0.3])
    assert r.status == 0 and r.n_accepted == 0
    assert np.array_equal(r.y_final.cpu().numpy(), u0)
    assert np.array_equal(r.y[0].cpu().numpy(), u0)
    # step budget: the safety net reports -5 instead of running on
    r = xb.solve_pde_rkc("heat2d_reaction", (0, 0.05), u0, rho_jac=float(rho),
                         rtol=1e-6, atol=1e-6, max_steps=3)
    assert r.status == -5 and r.n_accepted == 3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_rkc_two_rank_slab_decomposition_matches_one_rank(tmp_path):
    """Domain decomposition: 2 ranks, NCCL halo rows
    per stage + scalar all-gather per step; same step counts as 1 rank and the
    same state to 1e-12 (the error-norm partial sums are added in a different
    order)."""
    out = tmp_path / "mp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests",
                                                   "mp_rkc_worker.py"),
           str(out)]
    subprocess.check_call(cmd, cwd=ROOT, timeout=600)
    import json
    res = json.loads(out.read_text())
    assert res["ok"], res


USER_PDE_SRC = r"""
// u_t = D*Lap(u) - b*(u_x) + u*(1 - u) + A*sin(w t)*x*(1-x)*y*(1-y)
// p = (D, b, A, w); upwind-free central difference for u_x
__device__ double adr(double t, double x, double y, double inv_h2, double uc,
                      double un, double us, double uw, double ue, const double* p) {
    const double lap = (((un + us) + (uw + ue)) - 4.0 * uc) * inv_h2;
    const double ux = (ue - uw) * (0.5 * sqrt(inv_h2));
    const double src = p[2] * sin(p[3] * t) * (x * (1.0 - x)) * (y * (1.0 - y));
    return ((p[0] * lap - p[1] * ux) + uc * (1.0 - uc)) + src;
}
"""


def test_user_pde_rhs_vs_c_oracle():
    """A user right-hand side (advection-diffusion-reaction with a
    time-dependent source) registered as CUDA source and compiled into the
    fused stage kernel; compared with the C oracle driving the same function
    written in NumPy.  Exercises the time and coordinate arguments."""
    nx = 48
    prm = [0.5, 3.0, 40.0, 25.0]
    h = 1.0 / (nx + 1)
    xs = np.arange(1, nx + 1) * h
    X, Y = np.meshgrid(xs, xs)            # X varies along a row, Y down rows
    u0 = np.sin(np.pi * X) * np.sin(2 * np.pi * Y) * 0.3
    inv_h2 = (nx + 1.0) ** 2
    work = np.zeros((nx + 2, nx + 2))

    def fun(t, yv):
        work[1:-1, 1:-1] = yv.reshape(nx, nx)
        u = work[1:-1, 1:-1]
        un, us = work[:-2, 1:-1], work[2:, 1:-1]
        uw, ue = work[1:-1, :-2], work[1:-1, 2:]
        lap = (((un + us) + (uw + ue)) - 4.0 * u) * inv_h2
        ux = (ue - uw) * (0.5 * np.sqrt(inv_h2))
        src = prm[2] * np.sin(prm[3] * t) * (X * (1.0 - X)) * (Y * (1.0 - Y))
        return (((prm[0] * lap - prm[1] * ux) + u * (1.0 - u)) + src).reshape(-1)

    pde = xb.PdeRHS.from_source(USER_PDE_SRC, "adr", n_param=4)
    rho = 8.0 * inv_h2 * prm[0] + 4.0
    te = np.linspace(0.0, 0.2, 6)
    r = xb.solve_pde_rkc(pde, (0.0, 0.2), u0, rho_jac=float(rho), rtol=1e-5,
                         atol=1e-5, t_eval=te, pde_params=prm, max_steps=5000)
    ref = CO.rkc_solve(u0.reshape(-1), (0.0, 0.2), rtol=1e-5, atol=1e-5,
                       rho=rho, t_eval=te, fun=fun)
    assert r.status == 0
    assert (r.n_accepted, r.n_rejected, r.nfev, r.maxm) == \
        (ref["n_accepted"], ref["n_rejected"], ref["nfev"], ref["maxm"])
    got = r.y.cpu().numpy().reshape(te.size, -1).T
    assert np.abs(got - ref["y"]).max() <= 1e-10
    # power iteration instead of rho_jac, same user kernel
    r2 = xb.solve_pde_rkc(pde, (0.0, 0.05), u0, rtol=1e-4, atol=1e-4,
                          pde_params=prm, max_steps=5000)
    ref2 = CO.rkc_solve(u0.reshape(-1), (0.0, 0.05), rtol=1e-4, atol=1e-4,
                        rho=None, fun=fun)
    # The estimated spectral radius enters m = 1 + int(sqrt(1.54 h sprad + 1)):
    # for this non-symmetric Jacobian the block-wise reduction order of the
    # norms moves sprad in the last digits and int() flips once, after which
    # the step sequences differ (both valid).  Same number of power-iteration
    # evaluations, similar work, same solution to the tolerance.
    assert r2.status == 0 and r2.nfesig == ref2["nfesig"]
    assert abs(r2.nfev - ref2["nfev"]) <= 0.1 * ref2["nfev"]
    assert np.abs(r2.y_final.cpu().numpy().reshape(-1) -
                  ref2["y_final"]).max() <= 10 * 1e-4
    with pytest.raises(ValueError, match="pde_params"):
        xb.solve_pde_rkc(pde, (0.0, 0.05), u0, rho_jac=float(rho))


# ---- general systems: the reference's published 3-D table ---------------------
def test_ssv2stab_general_system_reproduces_the_notebook_table():
    """docs/Demo_SSV2stab.ipynb:350-356 (3-D heat problem on 39^3 points, linear
    with a time dependent source and boundary): steps (rejected) / f-evals / s-max
    for tol = 1e-1 .. 1e-4, as published, through PdeRHS.from_vector_source --
    one device function per component instead of the built-in 2-D stencil."""
    from oracle.problems import HEAT3D_VECTOR_SRC, heat3d_notebook
    N = 39
    _, y0, rho = heat3d_notebook(N)
    rhs = xb.PdeRHS.from_vector_source(HEAT3D_VECTOR_SRC % N, "heat3d", N ** 3)
    table = {1e-1: (6, 1, 402, 132), 1e-2: (15, 4, 729, 85),
             1e-3: (27, 2, 786, 40), 1e-4: (57, 0, 1087, 26)}
    meta = {c["id"]: c for c in CASES}
    for tol, (steps, rej, nfev, smax) in table.items():
        r = xb.solve_pde_rkc(rhs, (0.0, 0.7), y0, rtol=tol, atol=tol, const_jac=True,
                             rho_jac=float(rho))
        assert r.status == 0
        assert (r.n_accepted + r.n_rejected, r.n_rejected, r.nfev, r.maxm) == (steps, rej, nfev, smax)
        cid = f"heat3d_tol{tol:g}"
        assert cid in meta
        y = r.y_final.cpu().numpy()
        assert y.shape == (N ** 3,)
        assert np.abs(y[::97] - Z[cid + "/y_sample"]).max() <= 1e-9


def test_ssv2stab_two_species_combustion_table():
    """docs/Demo_SSV2stab.ipynb cell 9 (printed table): the 3-D two-species
    combustion problem, 2 x 40^3 equations, nonlinear, spectral radius by the
    power iteration: steps (rejected) / f-evals / f-sigma / s-max as published."""
    from oracle.problems import COMBUSTION_VECTOR_SRC
    N = 40
    L, alpha, delta, R = 0.9, 1.0, 20.0, 5.0
    D = R * np.exp(delta) / (alpha * delta)
    rhs = xb.PdeRHS.from_vector_source(COMBUSTION_VECTOR_SRC % N, "combustion", 2 * N ** 3, 4)
    y0 = np.ones(2 * N ** 3)
    table = {1e-4: (51, 1, 525, 21, 36), 1e-5: (124, 0, 781, 27, 29), 1e-6: (270, 0, 1270, 39, 20)}
    for tol, (steps, rej, nfev, nsig, smax) in table.items():
        r = xb.solve_pde_rkc(rhs, (0.0, 0.30), y0, rtol=tol, atol=tol, pde_params=[L, alpha, delta, D])
        assert r.status == 0
        got = (r.n_accepted + r.n_rejected, r.n_rejected, r.nfev, r.nfesig, r.maxm)
        print(f"combustion tol {tol:g}: steps {got[0]} ({got[1]}) f-evals {got[2]} f-sigma {got[3]} s-max {got[4]}")
        assert got == (steps, rej, nfev, nsig, smax)
        T = r.y_final.cpu().numpy()[N ** 3:]
        assert 1.0 <= T.min() and T.max() < 2.2          # burnt gas behind the front
