"""TEST INFRASTRUCTURE (dev-time checker: compares the device path with oracle/;
not part of the product, not used by bench.py).  GPU box: find where the device's CKdisc step sequence leaves the oracle's
(one lane; the device is stopped after k attempts with max_steps=k)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb
from oracle import rk_oracle as RO
from oracle.problems import CUDA_SOURCES, make_fun

lane = int(sys.argv[1]) if len(sys.argv) > 1 else 35
rng = np.random.default_rng(7)
y0 = rng.uniform(60.0, 140.0, (48, 1))[lane]
kw = dict(rtol=1e-6, atol=1e-8)
TAB = RO.load_ckdisc()
fun = make_fun("detest_f2", [])
st = RO.RKState(TAB, fun, 0.0, y0, 6.0, nfev_stiff_detect=0, **kw)
seq = []
while st.t < 6.0:
    ok, _ = RO.ckdisc_step(st)
    seq.append((st.n_accepted, st.n_rejected, st.t, st.y[0], st.h_abs, st.nfev, st.order_accepted,
                tuple(st.twiddle), tuple(st.quit)))
n, p, src = CUDA_SOURCES["detest_f2"]
rhs = xb.DeviceRHS.from_source(src, "rhs", n, p)
by_attempts = {a + r: s for s in seq for a, r in [(s[0], s[1])]}
prev_ok = None
for k in range(1, seq[-1][0] + seq[-1][1] + 1):
    r = xb.solve_ivp_batched(rhs, (0.0, 6.0), [y0], xb.CKdisc, max_steps=k, **kw)
    torch.cuda.synchronize()
    g = (int(r.n_accepted[0]), int(r.n_rejected[0]), float(r.t_final[0]), float(r.y_final[0, 0]),
         float(r.h_next[0]), int(r.nfev[0]))
    if k in by_attempts:
        o = by_attempts[k]
        match = g[0] == o[0] and g[1] == o[1] and abs(g[2] - o[2]) <= 1e-9 * abs(o[2]) and g[5] == o[5]
        if not match:
            print("first mismatch after", k, "attempts")
            print("  device:", g)
            print("  oracle:", o)
            if prev_ok:
                print("  last agreeing state:", prev_ok)
            break
        prev_ok = (k, g, o)
else:
    print("sequences agree")
