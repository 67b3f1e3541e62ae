"""GPU box: one SWAG solve of N perturbed Arenstorf orbits over a full period
(the C4 workload at reduced lane count) -- the target of an ncu capture."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb
N = int(os.environ.get("N", 150000))
rng = np.random.default_rng(2024)
y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) + rng.uniform(-1e-3, 1e-3, (N, 4))
y0 = torch.tensor(y0, device="cuda"); prm = torch.full((N, 1), 0.012277471, dtype=torch.float64, device="cuda")
T = 17.0652165601579625588917206249
for _ in range(2):
    r = xb.solve_ivp_batched("arenstorf", (0., T), y0, xb.SWAG, params=prm, rtol=1e-8, atol=1e-10, max_steps=200000)
    torch.cuda.synchronize()
print(int(r.n_accepted.sum()), int(r.n_rejected.sum()))
