"""GPU box: tiny runs of the newer code paths for compute-sanitizer
(`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb

QUICK = os.environ.get("SANITIZE_QUICK") == "1"      # the event-queue paths only, small
N = 96 if QUICK else 300
rng = np.random.default_rng(3)
y0 = np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(10, 35, N)], 1)
prm = np.tile([10.0, 28.0, 8.0 / 3.0], (N, 1))
EV = r"""
__device__ double event(int k, double t, const double* y, const double* p) {
    return k == 0 ? y[2] - 27.0 : y[0];
}"""
ev = xb.DeviceEvents.from_source(EV, "event", 2, terminal=[0, 3], direction=[1, 0])
te = np.linspace(0, 2, 9)
for m in ((xb.Ts5,) if QUICK else (xb.Ts5, xb.BS5, xb.CKdisc)):
    r = xb.solve_ivp_batched("lorenz63", (0., 2.), y0, m, params=prm, events=ev, t_eval=te,
                             max_event_records=4, rtol=1e-6, atol=1e-9)
    torch.cuda.synchronize()
    print(m.__name__, "events", int(r.event_counts.sum()), "status1", int((r.status == 1).sum()))
# event queue: general kernel (t_eval above), rk_fast with event hooks (no t_eval; terminal and
# not), a queue that overflows into the in-lane path, and no queue at all
for term in ([0, 0], [0, 3]):
    evq = ev.with_attributes(terminal=term, direction=[1, 0])
    for q in ((None, "200") if QUICK else (None, "700", "0")):
        if q is None:
            os.environ.pop("XSQ_EVENT_QUEUE_RECORDS", None)
        else:
            os.environ["XSQ_EVENT_QUEUE_RECORDS"] = q
        for m in ((xb.Ts5, xb.CKdisc) if QUICK else (xb.Ts5, xb.Pr8, xb.CKdisc)):
            r = xb.solve_ivp_batched("lorenz63", (0., 4.), y0, m, params=prm, events=evq,
                                     max_event_records=6, rtol=1e-6, atol=1e-9)
            torch.cuda.synchronize()
            print("event queue", q or "default", term, m.__name__, "events", int(r.event_counts.sum()),
                  "status1", int((r.status == 1).sum()))
os.environ.pop("XSQ_EVENT_QUEUE_RECORDS", None)
if QUICK:
    print("done (quick)")
    sys.exit(0)
# stiffness probe queue and slots (queue forced small -> both paths)
mu = 10.0 ** (-1 + 3 * np.arange(N) / (N - 1))
for q in ("", "0", "50"):
    if q:
        os.environ["XSQ_STIFF_QUEUE_RECORDS"] = q
    r = xb.solve_ivp_batched("vanderpol", (0., 20.), np.tile([2.0, 0.0], (N, 1)), xb.Ts5,
                             params=mu[:, None], rtol=1e-6, atol=1e-8, nfev_stiff_detect=300,
                             max_steps=200000)
    torch.cuda.synchronize()
    print("stiff queue", q or "default", "flagged", int((r.stiff_flags != 0).sum()), "nfev", int(r.nfev.sum()))
print("done")
