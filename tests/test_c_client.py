"""The C ABI used from plain C (tests/c_client/lorenz_client.c, gcc, no Python
or torch in the client): links against libxsq.so, fails loudly without a
device, and on a B200 gives the same per-lane results as the Python host layer."""
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "extensisq_b200")


def build_client(tmp_path):
    exe = str(tmp_path / "lorenz_client")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_client", "lorenz_client.c"), "-o", exe,
                           "-L", LIBDIR, "-lxsq", "-Wl,-rpath," + LIBDIR])
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_c_client_links_and_fails_loudly_without_a_device(tmp_path):
    exe = build_client(tmp_path)
    p = subprocess.run([exe, "4", "1.0"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout == ""
    assert "CUDA error" in p.stderr


@pytest.mark.gpu
def test_c_client_matches_the_python_host_layer(tmp_path):
    import extensisq_b200 as xb
    exe = build_client(tmp_path)
    N, T = 256, 2.0
    out = subprocess.run([exe, str(N), str(T)], capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in out.strip().split("\n")]
    assert len(rows) == N
    i = np.arange(N)
    y0 = np.stack([1.0 + 0.01 * i, np.ones(N), 20.0 - 0.02 * i], 1)
    prm = np.tile([10.0, 28.0, 8.0 / 3.0], (N, 1))
    r = xb.solve_ivp_batched("lorenz63", (0.0, T), y0, xb.Ts5, params=prm, rtol=1e-6, atol=1e-9,
                             max_steps=1000000)
    torch.cuda.synchronize()
    acc = np.array([int(x[1]) for x in rows]); rej = np.array([int(x[2]) for x in rows])
    nfev = np.array([int(x[3]) for x in rows]); st = np.array([int(x[4]) for x in rows])
    y = np.array([[float.fromhex(v) for v in x[5:8]] for x in rows])
    assert (st == 0).all()
    assert np.array_equal(acc, r.n_accepted.cpu().numpy())
    assert np.array_equal(rej, r.n_rejected.cpu().numpy())
    assert np.array_equal(nfev, r.nfev.cpu().numpy())
    assert np.array_equal(y, r.y_final.cpu().numpy())
