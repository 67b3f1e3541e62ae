"""GPU box: throughput of the §8(f) additions on the C2 ensemble shape
(1.25 M Lorenz lanes, t in [0, T]): CKdisc, events (Poincare section), forward
sensitivities.  Device-resident timing with CUDA events; one JSON line each."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb

EVENT_SRC = r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[2] - 27.0;          // Poincare section z = 27
    if (k == 1) return y[0];                 // x = 0
    return y[0] * y[1] - 30.0;
}"""
SENS_SRC = r"""
__device__ void fun(double t, const double* y, const double* p, double* dy) {
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = y[0] * (p[1] - y[2]) - y[1];
    dy[2] = y[0] * y[1] - p[2] * y[2];
}
__device__ void jac(double t, const double* y, const double* p, double* J) {
    J[0] = -p[0];        J[1] = p[0];  J[2] = 0.;
    J[3] = p[1] - y[2];  J[4] = -1.;   J[5] = -y[0];
    J[6] = y[1];         J[7] = y[0];  J[8] = -p[2];
}
__device__ void dfdp(double t, const double* y, const double* p, double* D) {
    D[0] = y[1] - y[0];  D[1] = 0.;    D[2] = 0.;
    D[3] = 0.;           D[4] = y[0];  D[5] = 0.;
    D[6] = 0.;           D[7] = 0.;    D[8] = -y[2];
}"""


def timed(fn, reps=1):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


N = int(os.environ.get("LANES", 1_250_000))
T = float(os.environ.get("TEND", 20.0))
rng = np.random.default_rng(12345)
y0 = torch.tensor(np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 40, N)], 1), device="cuda")
prm_np = np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1)
prm = torch.tensor(prm_np, device="cuda")
kw = dict(rtol=1e-8, atol=1e-10)

r, ms = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., T), y0, xb.Ts5, params=prm, **kw))
base = int(r.n_accepted.sum())
print(json.dumps(dict(config="Ts5 plain", lanes=N, T=T, ms=ms, steps_per_s=base / ms * 1e3)))
r, ms = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., T), y0, xb.CKdisc, params=prm, **kw))
print(json.dumps(dict(config="CKdisc", lanes=N, T=T, ms=ms, steps_per_s=int(r.n_accepted.sum()) / ms * 1e3)))
ev = xb.DeviceEvents.from_source(EVENT_SRC, "event", 3, terminal=[0, 0, 0], direction=[1, 0, 0])
r, ms = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., T), y0, xb.Ts5, params=prm, events=ev,
                                           max_event_records=48, **kw))
print(json.dumps(dict(config="Ts5 + 3 event functions (NVRTC kernel)", lanes=N, T=T, ms=ms,
                      steps_per_s=int(r.n_accepted.sum()) / ms * 1e3,
                      events_found=int(r.event_counts.sum()),
                      events_per_s=int(r.event_counts.sum()) / ms * 1e3)))
M = N // 4
s, yf, sol = None, None, None
def sens():
    return xb.sens_forward(SENS_SRC, (0., T), y0[:M].cpu().numpy(), np.zeros((3, 3)), prm_np[:M],
                           method=xb.Ts5, **kw)
(s, yf, sol), ms = timed(sens)
print(json.dumps(dict(config="sens_forward Ts5, 12 states per lane (y + dy/d(sigma,rho,beta))", lanes=M, T=T,
                      ms=ms, steps_per_s=int(sol.n_accepted.sum()) / ms * 1e3,
                      note="includes host->device staging of y0/p in sens_forward")))
