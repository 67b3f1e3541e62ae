"""GPU box: throughput of the other BASELINE.json configs (C3, C4, C5) --
device-resident timing with CUDA events; one JSON line per config."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["c3", "c4a", "c4b", "c5", "c2ck5"]
dev = torch.device("cuda")
if "c2ck5" in which:
    N = 1_250_000
    rng = np.random.default_rng(12345)
    y0 = torch.tensor(np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 40, N)], 1), device=dev)
    prm = torch.tensor(np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1), device=dev)
    for m in (xb.CK5, xb.BS5, xb.Pr8):
        r, ms = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., 100.), y0, m, params=prm, rtol=1e-8, atol=1e-10), 1)
        acc = int(r.n_accepted.sum())
        print(json.dumps(dict(config="C2 " + m.__name__, lanes=N, ms=ms, accepted=acc, rejected=int(r.n_rejected.sum()), steps_per_s=acc / ms * 1e3, ok=bool((r.status == 0).all()))))
if "c3" in which:
    N = int(os.environ.get("C3_LANES", 1_000_000))
    mu = 10.0 ** (-1 + 3 * np.arange(N) / (N - 1))
    y0 = torch.tensor(np.tile([2.0, 0.0], (N, 1)), device=dev)
    prm = torch.tensor(mu[:, None], device=dev)
    te = torch.linspace(0, 20, 1000, dtype=torch.float64, device=dev)
    for m in (xb.Pr8, xb.Pr9):
        for with_eval in (False, True):
            r, ms = timed(lambda: xb.solve_ivp_batched("vanderpol", (0., 20.), y0, m, params=prm, rtol=1e-8, atol=1e-10, t_eval=te if with_eval else None), 1)
            acc = int(r.n_accepted.sum())
            print(json.dumps(dict(config="C3 " + m.__name__, t_eval=with_eval, lanes=N, ms=ms, accepted=acc, rejected=int(r.n_rejected.sum()), steps_per_s=acc / ms * 1e3, out_GB=(N * 2 * 1000 * 8 / 1e9 if with_eval else 0), ok=bool((r.status == 0).all()))))
            del r
if "c4a" in which:
    N = 1_000_000
    rng = np.random.default_rng(2024)
    y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) + rng.uniform(-1e-3, 1e-3, (N, 4))
    y0 = torch.tensor(y0, device=dev); prm = torch.full((N, 1), 0.012277471, dtype=torch.float64, device=dev)
    T = 17.0652165601579625588917206249
    for m in (xb.SWAG, xb.Pr8):
        r, ms = timed(lambda: xb.solve_ivp_batched("arenstorf", (0., T), y0, m, params=prm, rtol=1e-8, atol=1e-10, max_steps=200000), 1)
        acc = int(r.n_accepted.sum())
        print(json.dumps(dict(config="C4i arenstorf " + m.__name__, lanes=N, ms=ms, accepted=acc, rejected=int(r.n_rejected.sum()), steps_per_s=acc / ms * 1e3, ok_frac=float((r.status == 0).double().mean()))))
if "c4b" in which:
    N, nb = 65536, 32
    rng = np.random.default_rng(2025)
    m_ = rng.uniform(0.5, 1.5, (N, nb)); pos = rng.normal(0, 1, (N, nb, 3)); vel = rng.normal(0, 0.3, (N, nb, 3))
    vel -= (m_[:, :, None] * vel).sum(1, keepdims=True) / m_.sum(1)[:, None, None]
    y0 = torch.tensor(np.concatenate([pos.reshape(N, -1), vel.reshape(N, -1)], 1), device=dev)
    prm = torch.tensor(np.concatenate([np.full((N, 1), 0.05 ** 2), m_], 1), device=dev)
    for m in (xb.SWAG, xb.Pr8):
        r, ms = timed(lambda: xb.solve_ivp_batched("nbody32", (0., 1.), y0, m, params=prm, rtol=1e-8, atol=1e-10, max_steps=200000), 1)
        acc = int(r.n_accepted.sum())
        print(json.dumps(dict(config="C4ii nbody32 " + m.__name__, systems=N, ms=ms, accepted=acc, rejected=int(r.n_rejected.sum()), steps_per_s=acc / ms * 1e3, nfev=int(r.nfev.sum()), pair_interactions_per_s=int(r.nfev.sum()) * 32 * 32 / ms * 1e3, ok_frac=float((r.status == 0).double().mean()))))
if "c5" in which:
    for nx in (4096, 16384):
        h = 1.0 / (nx + 1)
        x = torch.arange(1, nx + 1, dtype=torch.float64, device=dev) * h
        u0 = torch.outer(torch.sin(np.pi * x), torch.sin(np.pi * x))
        rho = 8.0 * (nx + 1.0) ** 2 + 2.0
        T = 0.05 * (512 / nx) ** 2
        t0 = time.perf_counter()
        r = xb.solve_pde_rkc("heat2d_reaction", (0.0, T), u0, rho_jac=float(rho), rtol=1e-4, atol=1e-4, max_steps=1000)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        pts = nx * nx
        print(json.dumps(dict(config=f"C5 ssv2stab {nx}^2 one GPU", T=T, s=dt, accepted=r.n_accepted, rejected=r.n_rejected, nfev=r.nfev, maxm=r.maxm, status=r.status, point_stages_per_s=pts * r.nfev / dt, algorithmic_GBps=pts * r.nfev * 40 / dt / 1e9, launches=r.kernel_launches)))
        del u0, r
