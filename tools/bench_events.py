"""GPU box: cost of event location on the C2 ensemble shape (Lorenz lanes,
three event functions, Poincare sections): plain solve, roots located inside
the lane (XSQ_EVENT_QUEUE_RECORDS=0, the round-1 path), roots through the event
queue, and a terminal-count variant.  One JSON line each."""
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


import ctypes as C
from extensisq_b200 import _lib
LIB = _lib.load()
LIB.xsq_profile_enable(1)


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


def kernels():
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    if LIB.xsq_profile_last(C.byref(a), C.byref(b), C.byref(c)) != 0:
        return None
    return dict(init_ms=a.value, persistent_ms=b.value, queues_ms=c.value)


N = int(os.environ.get("LANES", 1_250_000))
T = float(os.environ.get("TEND", 20.0))
CAP = int(os.environ.get("EVCAP", 48))
rng = np.random.default_rng(12345)
y0 = torch.tensor(np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 40, N)], 1), device="cuda")
prm = torch.tensor(np.stack([rng.uniform(9, 11, N), rng.uniform(24, 32, N), rng.uniform(2.4, 2.9, N)], 1), device="cuda")
kw = dict(rtol=1e-8, atol=1e-10)
r, ms0 = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., T), y0, xb.Ts5, params=prm, **kw))
print(json.dumps(dict(config="Ts5 plain", lanes=N, T=T, ms=ms0, steps_per_s=int(r.n_accepted.sum()) / ms0 * 1e3)), flush=True)
ref = None
MINB = os.environ.get("MINB_SWEEP", "")
runs = [("in-lane roots", [0, 0, 0], "0", ""), ("event queue", [0, 0, 0], None, ""),
        ("event queue, 2nd event terminal at its 40th occurrence", [0, 40, 0], None, "")]
runs += [(f"event queue, {b} CTAs/SM", [0, 0, 0], None, b) for b in MINB.split(",") if b]
runs += [(f"event queue, queue kernel compiled for {b} CTAs/SM", [0, 0, 0], None, "q" + b)
         for b in os.environ.get("EVQ_MINB_SWEEP", "").split(",") if b]
for name, term, q, minb in runs:
    os.environ.pop("XSQ_USER_MINB", None)
    os.environ.pop("XSQ_EVQ_MINB", None)
    if minb.startswith("q"):
        os.environ["XSQ_EVQ_MINB"] = minb[1:]
    elif minb:
        os.environ["XSQ_USER_MINB"] = minb
    if q is None:
        os.environ.pop("XSQ_EVENT_QUEUE_RECORDS", None)
    else:
        os.environ["XSQ_EVENT_QUEUE_RECORDS"] = q
    ev = xb.DeviceEvents.from_source(EVENT_SRC, "event", 3, terminal=term, direction=[1, 0, 0])
    r, ms = timed(lambda: xb.solve_ivp_batched("lorenz63", (0., T), y0, xb.Ts5, params=prm, events=ev,
                                               max_event_records=CAP, **kw))
    out = dict(config="Ts5 + 3 event functions: " + name, lanes=N, T=T, ms=ms, vs_plain=ms / ms0,
               steps_per_s=int(r.n_accepted.sum()) / ms * 1e3, events_found=int(r.event_counts.sum()),
               events_per_s=int(r.event_counts.sum()) / ms * 1e3, kernels=kernels())
    if term == [0, 0, 0]:
        cur = (r.t_events.clone(), r.y_events.clone(), r.event_counts.clone())
        if ref is None:
            ref = cur
        else:
            out["bit_identical_to_in_lane"] = all(
                bool(((a == b) | ((a != a) & (b != b))).all()) for a, b in zip(ref, cur))
    print(json.dumps(out), flush=True)
