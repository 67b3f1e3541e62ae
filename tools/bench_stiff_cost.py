"""GPU box: cost of the stiffness diagnosis (nfev_stiff_detect 0 vs 5000) per config."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb


def timed(fn, reps=1):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


dev = torch.device("cuda")
N = int(os.environ.get("LANES", 1_000_000))
ONLY = os.environ.get("ONLY")
NSD = [int(v) for v in os.environ.get("NSD", "0,5000").split(",")]
rng = np.random.default_rng(2024)
y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) + rng.uniform(-1e-3, 1e-3, (N, 4))
y0 = torch.tensor(y0, device=dev)
prm = torch.full((N, 1), 0.012277471, dtype=torch.float64, device=dev)
T = 17.0652165601579625588917206249
mu = 10.0 ** (-1 + 3 * np.arange(N) / (N - 1))
y0v = torch.tensor(np.tile([2.0, 0.0], (N, 1)), device=dev)
prmv = torch.tensor(mu[:, None], device=dev)
for name, rhs, span, a, b, m in (("arenstorf Pr8", "arenstorf", (0., T), y0, prm, xb.Pr8),
                                 ("arenstorf Ts5", "arenstorf", (0., T), y0, prm, xb.Ts5),
                                 ("vanderpol Pr8", "vanderpol", (0., 20.), y0v, prmv, xb.Pr8)):
    if ONLY and ONLY != name:
        continue
    for nsd in NSD:
        r, ms = timed(lambda: xb.solve_ivp_batched(rhs, span, a, m, params=b, rtol=1e-8, atol=1e-10,
                                                   max_steps=int(os.environ.get("MAXSTEPS", 200000)), nfev_stiff_detect=nsd))
        print(json.dumps(dict(config=name, nfev_stiff_detect=nsd, ms=ms, accepted=int(r.n_accepted.sum()),
                              rejected=int(r.n_rejected.sum()), nfev=int(r.nfev.sum()),
                              flagged=int((r.stiff_flags != 0).sum()), failed=int((r.status != 0).sum()),
                              max_attempts=int((r.n_accepted + r.n_rejected).max()))))
